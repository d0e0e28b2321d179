"""TEST INFRASTRUCTURE — container-only loader for the reference's text front-end
(`/root/reference/src/auralis/models/xttsv2/config/tokenizer.py`), used by `tests/golden/make_text_golden.py`
and by the CPU tests that pin `auralis_b200/textnorm.py` + `text.py` when /root/reference is mounted.

The reference module imports five third-party packages that are not installed here (spacy, num2words, pypinyin,
hangul_romanize, cutlet).  They are replaced by stubs in `sys.modules` *for the import of that one module*:

  * `num2words` -> `marker_num2words`: returns a digit-free marker that encodes every argument of the call, so the
    reference's own orchestration (which substrings are verbalised, in which order, with which arguments, how
    integer currency amounts are trimmed) is pinned exactly even though num2words' word lists are not available;
  * `spacy.lang.*` -> a pipeline whose `sentencizer` is the punctuation rule of `auralis_b200.text.sentencize`
    (spaCy's rule-based sentencizer is third-party; what the reference adds — packing sentences into chunks, splitting
    over-long sentences, the trailing-dot rule — is its own code and is what this pins);
  * pypinyin / hangul_romanize / cutlet -> placeholders that raise if a test reaches transliteration.

Nothing under `auralis_b200/` imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from . import ref_import

_DIGITS = "abcdefghij"


def encode_number(x) -> str:
    """Digit-free spelling of a number literal: 12.5 -> 'bcpf' (digits a..j, 'p' = point, 'm' = minus)."""
    s = repr(x) if isinstance(x, float) else str(x)
    out = []
    for ch in s:
        if ch.isdigit():
            out.append(_DIGITS[int(ch)])
        elif ch == ".":
            out.append("p")
        elif ch == "-":
            out.append("m")
        elif ch in "e+":
            out.append("x")
        else:
            raise ValueError(f"unexpected character {ch!r} in number {s!r}")
    return "".join(out)


AND_EQUIVALENTS = {"en": ", ", "es": " con ", "fr": " et ", "de": " und ", "pt": " e ", "it": " e ", "pl": ", ",
                   "cs": ", ", "ru": ", ", "nl": ", ", "ar": ", ", "tr": ", ", "hu": ", ", "ko": ", "}


def marker_num2words(number, ordinal=False, lang="en", to="cardinal", **kwargs) -> str:
    """Stand-in with num2words' call signature.  Currency markers contain the language's unit separator once, like a
    real 'five dollars, fifty cents', so the reference's trimming of integer amounts acts on them."""
    lang_tag = "cz" if lang == "cz" else lang
    if to == "currency":
        cur = kwargs.get("currency", "EUR").lower()
        units = int(number)
        cents = int(round((float(number) - units) * 100))
        sep = AND_EQUIVALENTS.get("cs" if lang == "cz" else lang, ", ")
        return f"qcur{lang_tag}q{cur}q{encode_number(units)}{sep}qsub{encode_number(cents)}"
    kind = "ord" if (ordinal or to == "ordinal") else ("dec" if isinstance(number, float) else "card")
    return f"q{kind}{lang_tag}q{encode_number(number)}q"


class _Span:
    def __init__(self, text):
        self.text = text

    def __str__(self):
        return self.text


class _Doc:
    def __init__(self, sents):
        self.sents = [_Span(s) for s in sents]


class _Lang:
    """Minimal `spacy.lang.xx.Xx()` look-alike: add_pipe("sentencizer") + __call__ -> doc.sents."""
    lang = "en"

    def __init__(self):
        self.pipe_names = []

    def add_pipe(self, name):
        if name != "sentencizer":
            raise ValueError(name)
        self.pipe_names.append(name)

    def __call__(self, text):
        from auralis_b200.text import sentencize          # the restated spaCy rules (see module docstring)
        return _Doc(sentencize(text, self.lang))


def _raising(name):
    def f(*a, **k):
        raise RuntimeError(f"{name} is not installed in this container (stub)")
    return f


def load():
    """The reference tokenizer module, imported unmodified with the stubs above."""
    if not ref_import.available():
        raise RuntimeError("reference tree not mounted")
    ref_import.load()                                     # stub parent packages (auralis, auralis.models, ...)
    base = os.path.join(ref_import.REF_SRC, "auralis", "models", "xttsv2")
    ref_import._stub("auralis.models.xttsv2.config", os.path.join(base, "config"))
    saved = {}

    def put(name, mod):
        saved[name] = sys.modules.get(name)
        sys.modules[name] = mod

    n2w = types.ModuleType("num2words"); n2w.num2words = marker_num2words
    pyp = types.ModuleType("pypinyin"); pyp.pinyin = _raising("pypinyin"); pyp.Style = types.SimpleNamespace(TONE3=None)
    hr = types.ModuleType("hangul_romanize"); hr.Transliter = lambda rule: types.SimpleNamespace(translit=_raising("hangul_romanize"))
    hrr = types.ModuleType("hangul_romanize.rule"); hrr.academic = object()
    cut = types.ModuleType("cutlet"); cut.Cutlet = _raising("cutlet")
    put("num2words", n2w); put("pypinyin", pyp); put("hangul_romanize", hr); put("hangul_romanize.rule", hrr); put("cutlet", cut)
    have_spacy = "spacy" in sys.modules
    if not have_spacy:
        sp = types.ModuleType("spacy"); sp.__path__ = []
        spl = types.ModuleType("spacy.lang"); spl.__path__ = []
        put("spacy", sp); put("spacy.lang", spl)
        for code, cls in (("ar", "Arabic"), ("en", "English"), ("es", "Spanish"), ("ja", "Japanese"), ("zh", "Chinese")):
            m = types.ModuleType(f"spacy.lang.{code}")
            setattr(m, cls, type(cls, (_Lang,), {"lang": code}))       # the language object decides the tokenizer rules
            put(f"spacy.lang.{code}", m)
    try:
        mod = importlib.import_module("auralis.models.xttsv2.config.tokenizer")
    finally:
        for name, old in saved.items():                   # do not leave stubs behind for unrelated imports
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return mod
