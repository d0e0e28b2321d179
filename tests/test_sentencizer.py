"""CPU: sentence boundaries (auralis_b200/sentencizer.py) — the restatement of spaCy's `sentencizer` + blank-language
tokenizer rules the reference splits text with (`config/tokenizer.py:25-48,177-183`).  spaCy is not installed here, so the
expectations below are derived by hand from spaCy's published rules (UNPINNED); `tests/golden/make_sentence_golden.py` writes
`tests/golden/sentences.json` from spaCy itself where it is available, and the last test then checks against it."""
import json
import os

import pytest

from auralis_b200.sentencizer import sentencize
from auralis_b200.text import split_sentence

CASES = [
    ("en", "Mr. Smith went to Washington. He arrived at 3 p.m. and left.", ["Mr. Smith went to Washington.", "He arrived at 3 p.m. and left."]),
    ("en", 'He left. "Why?" she said.', ['He left. "', 'Why?"', "she said."]),        # quotes stay with the sentence that ended
    ("en", "Wait... what? Really!", ["Wait... what?", "Really!"]),                    # an ellipsis token is not a stop
    ("en", "The U.S. is big. J. K. Rowling wrote it.", ["The U.S. is big.", "J. K. Rowling wrote it."]),
    ("en", "It cost 3.14 dollars. OK. USA. Done", ["It cost 3.14 dollars.", "OK.", "USA.", "Done"]),
    ("en", "end.Next sentence here", ["end.", "Next sentence here"]),                  # infix lower.Upper
    ("en", "No stop at all", ["No stop at all"]),
    ("fr", "Bonjour M. Dupont. Ça va? Oui!", ["Bonjour M. Dupont.", "Ça va?", "Oui!"]),   # English() rules for fr
    ("es", "El Sr. Pérez llegó. ¿Cómo está Ud. hoy? Bien.", ["El Sr. Pérez llegó. ¿", "Cómo está Ud. hoy?", "Bien."]),
    ("zh", "今天天气很好。我们去公园散步吧！你觉得怎么样？好的「走吧」。", ["今天天气很好。", "我们去公园散步吧！", "你觉得怎么样？", "好的「走吧」。"]),
    ("ja", "今日はいい天気です。散歩に行きましょう！どうですか？", ["今日はいい天気です。", "散歩に行きましょう！", "どうですか？"]),
    ("ar", "مرحبا بكم. كيف حالك؟ بخير!", ["مرحبا بكم.", "كيف حالك؟", "بخير!"]),
]


@pytest.mark.parametrize("lang,text,expected", CASES)
def test_sentences(lang, text, expected):
    assert sentencize(text, lang) == expected


def test_unspaced_cjk_is_split_at_stops_not_mid_word():
    """a 138-character zh string with limit 82 must be packed from whole sentences, never cut inside a word."""
    zh = "今天天气很好，我们去公园散步吧。" * 9
    chunks = split_sentence(zh, "zh", 82)
    assert all(len(c) <= 82 for c in chunks) and len(chunks) >= 2
    assert all(c.rstrip().endswith("。") for c in chunks)                 # every chunk ends at a sentence stop
    assert "".join(c.replace(" ", "") for c in chunks) == zh


def test_english_packing_unchanged_on_plain_prose():
    text = " ".join(f"This is sentence number {i} of a fairly ordinary paragraph." for i in range(12))
    chunks = split_sentence(text, "en", 250)
    assert all(len(c) <= 250 for c in chunks) and len(chunks) == 3
    assert " ".join(c.strip() + ("." if c.endswith(" ") else "") for c in chunks).replace("  ", " ").count("sentence number") == 12


def test_against_spacy_golden():
    path = os.path.join(os.path.dirname(__file__), "golden", "sentences.json")
    if not os.path.exists(path):
        pytest.skip("no spaCy golden in this checkout (spaCy is not installed here): run tests/golden/make_sentence_golden.py")
    for rec in json.load(open(path)):
        assert sentencize(rec["text"], rec["lang"]) == rec["sentences"], rec
