"""Generates tests/golden/sentences.json from the REFERENCE's own sentence path — spaCy blank language + `sentencizer`, exactly
as `auralis/models/xttsv2/config/tokenizer.py:25-48,177-183` builds it.  spaCy is not installed in the build image, so the file
is not committed; run this wherever `import spacy` works (`pip install spacy sudachipy sudachidict_core`) and
tests/test_sentencizer.py::test_against_spacy_golden picks it up.

    python tests/golden/make_sentence_golden.py"""
import json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_sentencizer import CASES       # noqa: E402  (the same inputs the hand-derived expectations use)


def spacy_lang(lang):
    from spacy.lang.ar import Arabic
    from spacy.lang.en import English
    from spacy.lang.es import Spanish
    from spacy.lang.ja import Japanese
    from spacy.lang.zh import Chinese
    return {"zh": Chinese, "ja": Japanese, "ar": Arabic, "es": Spanish}.get(lang, English)()


out = []
for lang, text, _ in CASES:
    nlp = spacy_lang(lang)
    nlp.add_pipe("sentencizer")
    out.append({"lang": lang, "text": text, "sentences": [str(s).strip() for s in nlp(text).sents if str(s).strip()]})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sentences.json")
json.dump(out, open(path, "w"), ensure_ascii=False, indent=1)
print(path, len(out))
