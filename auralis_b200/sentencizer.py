"""Sentence boundaries the way the reference gets them: spaCy's rule-based `sentencizer` pipe on top of a blank language
tokenizer (`config/tokenizer.py:25-48,177-183`: `Chinese()` / `Japanese()` / `Arabic()` / `Spanish()`, `English()` for every
other language, then `nlp.add_pipe("sentencizer")`).  spaCy is a third-party package that is neither vendored in the
reference nor installed in this image, so this is a restatement of its published rules — UNPINNED: no golden can be produced
here; `tests/golden/make_sentence_golden.py` regenerates one wherever spaCy is importable and `tests/test_sentencizer.py`
then checks it.

What matters for chunk boundaries (and therefore for token ids):

* `Sentencizer.predict`: a token whose text is one of `punct_chars` arms a flag; the next token that is neither punctuation
  nor a `punct_chars` token starts a new sentence.  Closing quotes / brackets after the stop, and OPENING ones of the next
  sentence, therefore stay with the sentence that just ended.
* Whether a '.' is a token of its own is the tokenizer's business.  English suffix rules split a trailing '.' when it follows
  a digit, a lower-case letter, a closing punctuation mark or quote, or two upper-case letters; it is NOT split after a single
  upper-case letter ("J.", "U.S."), in the tokenizer exceptions ("Mr.", "e.g.", "a.m.", …, and "a." … "z."), and runs of
  dots ("...", "…") are one ellipsis token, which is not in `punct_chars`.  An infix rule splits `lower.Upper` inside a word.
* `Chinese()` segments by character (its default), so every character is a token: a stop ends a sentence wherever it stands —
  unspaced text included — and closing brackets / quotes stay attached.  `Japanese()` needs SudachiPy in the
  reference; its stops are handled the same way.
"""
from __future__ import annotations

import unicodedata
from typing import List

# spacy.pipeline.sentencizer.Sentencizer.default_punct_chars (the BMP part that can occur in the 17 supported languages)
PUNCT_CHARS = frozenset("!.?։؟۔܀܁܂߹।॥၊။።፧፨᙮᜵᜶᠃᠉᥄᥅‼‽⁇⁈⁉⸮⸼꓿꘎꘏꛳꛷꡶꡷꣎꣏꤯꧈꧉꩝꩞꩟꫰꫱꯫﹒﹖﹗！．？｡。")

# spacy.lang.en.tokenizer_exceptions (entries ending in '.') + spacy.lang.tokenizer_exceptions.BASE_EXCEPTIONS
_EXC_EN = frozenset("""a.m. Adm. Bros. co. Co. Corp. D.C. Dr. e.g. E.g. E.G. Gen. Gov. i.e. I.e. I.E. Inc. Jr. Ltd. Md. Messrs.
Mo. Mont. Mr. Mrs. Ms. p.m. Ph.D. Prof. Rep. Rev. Sen. St. vs. v.s. Ala. Ariz. Ark. Aug. Calif. Colo. Conn. Dec. Del. Feb. Fla.
Ga. Ia. Id. Ill. Ind. Jan. Jul. Jun. Kan. Kans. Ky. La. Mar. Mass. Mich. Minn. Miss. N.C. N.D. N.H. N.J. N.M. N.Y. Neb. Nebr.
Nev. Nov. Oct. Okla. Ore. Pa. S.C. Sep. Sept. Tenn. Va. Wash. Wis.""".split())
# spacy.lang.es.tokenizer_exceptions
_EXC_ES = frozenset("""a.C. a.J.C. d.C. d.J.C. apdo. Av. Avda. Cía. Dr. Dra. EE.UU. etc. fig. Gob. Gral. Ing. J.C. km/h Lic. m.n.
núm. P.D. Prof. Profa. q.e.p.d. Q.E.P.D. S.A. S.L. S.R.L. s.s.s. Sr. Sra. Srta. Ud. Uds. Vd. Vds. pág. p.ej. vol. 12m.""".split())
_CLOSERS = frozenset("\"'”’»›)]}）】』」》〉")


def _is_punct(ch: str) -> bool:
    return unicodedata.category(ch).startswith("P")


def _period_is_token(word: str, exc) -> bool:
    """`word` (no whitespace, closers already stripped) ends in a single '.': is that dot split off as its own token?"""
    if word in exc:
        return False
    body = word[:-1]
    if not body:
        return True
    if len(body) == 1 and body.islower():                 # BASE_EXCEPTIONS "a." .. "z."
        return False
    prev = body[-1]
    if prev.isdigit() or prev.islower() or _is_punct(prev) or prev in _CLOSERS:
        return True
    if prev.isupper():
        return len(body) >= 2 and body[-2].isupper()       # suffix rule (?<=[A-Z][A-Z])\.
    return True


def _stops_words(text: str, lang: str) -> List[int]:
    """offsets just behind every sentence-ending token, for whitespace-tokenised scripts"""
    exc = _EXC_ES if lang == "es" else (frozenset() if lang == "ar" else _EXC_EN)
    stops, i, n = [], 0, len(text)
    while i < n:
        if text[i].isspace():
            i += 1
            continue
        j = i
        while j < n and not text[j].isspace():
            j += 1
        word = text[i:j]
        # infix rule: lower '.' Upper inside a word
        for k in range(1, len(word) - 1):
            if word[k] == "." and word[k - 1].islower() and word[k + 1].isupper():
                stops.append(i + k + 1)
        core = word
        while core and core[-1] in _CLOSERS:
            core = core[:-1]
        if core:
            last = core[-1]
            if last == ".":
                dots = len(core) - len(core.rstrip("."))
                if dots == 1 and _period_is_token(core, exc):
                    stops.append(i + len(core))
            elif last in PUNCT_CHARS:
                stops.append(i + len(core))
        i = j
    return stops


def _stops_chars(text: str) -> List[int]:
    return [i + 1 for i, ch in enumerate(text) if ch in PUNCT_CHARS]


def sentencize(text: str, lang: str = "en") -> List[str]:
    stops = _stops_chars(text) if lang in ("zh", "ja", "zh-cn") else _stops_words(text, lang)
    out, start, n = [], 0, len(text)
    for s in sorted(set(stops)):
        if s <= start:
            continue
        # the new sentence starts at the first character behind the stop that is neither space, punctuation nor a stop
        k = s
        while k < n and (text[k].isspace() or _is_punct(text[k]) or text[k] in PUNCT_CHARS):
            k += 1
        if k >= n:
            break
        if k > start:
            out.append(text[start:k])
            start = k
    if start < n:
        out.append(text[start:])
    return [x for x in (y.strip() for y in out) if x]
