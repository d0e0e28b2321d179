"""Per-language text cleaning in front of the BPE (the first "next" row of the hot path).

Restates the pipeline of the reference's `multilingual_cleaners` / `preprocess_text`
(`/root/reference/src/auralis/models/xttsv2/config/tokenizer.py:681-719,805-819`):

    drop '"'  ->  lowercase  ->  numbers  ->  abbreviations  ->  symbols  ->  collapse whitespace

The reference delegates number verbalisation to the third-party `num2words` package (not installed here, not
vendored by the reference).  English is restated below following num2words' published English conventions
("one thousand, two hundred and thirty-four", "three point one four", "twenty-first",
"five dollars, fifty cents"); for the other languages digits are kept and only the language-independent steps
run — that part of the row is still open and says so in DESIGN.md §7.
Transliteration (zh pinyin, ja romaji, ko) needs pypinyin / cutlet / hangul_romanize: not available, not restated.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Tuple

_ONES = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
         "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]
_SCALES = [(10 ** 18, "quintillion"), (10 ** 15, "quadrillion"), (10 ** 12, "trillion"), (10 ** 9, "billion"),
           (10 ** 6, "million"), (10 ** 3, "thousand")]
_ORD_IRREGULAR = {"one": "first", "two": "second", "three": "third", "five": "fifth", "eight": "eighth",
                  "nine": "ninth", "twelve": "twelfth"}


def _below_1000(n: int) -> str:
    parts = []
    if n >= 100:
        parts.append(_ONES[n // 100] + " hundred")
        n %= 100
        if n:
            parts.append("and")
    if n >= 20:
        parts.append(_TENS[n // 10] + ("-" + _ONES[n % 10] if n % 10 else ""))
    elif n or not parts:
        parts.append(_ONES[n])
    return " ".join(parts)


def cardinal_en(n: int) -> str:
    """1234 -> 'one thousand, two hundred and thirty-four'; 1001 -> 'one thousand and one'."""
    if n < 0:
        return "minus " + cardinal_en(-n)
    if n < 1000:
        return _below_1000(n)
    groups: List[Tuple[int, str]] = []          # (value of the group, text), most significant first
    rest = n
    for scale, name in _SCALES:
        if rest >= scale:
            q, rest = divmod(rest, scale)
            groups.append((q * scale, f"{cardinal_en(q)} {name}"))
    text = groups[0][1]
    for _, t in groups[1:]:
        text += ", " + t
    if rest:
        # a trailing group below one hundred is joined with "and", anything else with a comma
        text += (" and " if rest < 100 else ", ") + _below_1000(rest)
    return text


def ordinal_en(n: int) -> str:
    words = cardinal_en(n)
    head, sep, last = words.rpartition(" ")
    pre, hy, tail = last.rpartition("-")
    if tail in _ORD_IRREGULAR:
        tail = _ORD_IRREGULAR[tail]
    elif tail.endswith("y"):
        tail = tail[:-1] + "ieth"
    else:
        tail += "th"
    return head + sep + pre + hy + tail


def decimal_en(s: str) -> str:
    """'3.05' -> 'three point zero five' (digits after the point are read one by one)."""
    whole, _, frac = s.replace(",", ".").partition(".")
    return cardinal_en(int(whole or "0")) + " point " + " ".join(_ONES[int(d)] for d in frac)


_CURRENCY = {"USD": ("dollar", "dollars", "cent", "cents"), "GBP": ("pound", "pounds", "penny", "pence"),
             "EUR": ("euro", "euros", "cent", "cents")}


def currency_en(amount: float, code: str) -> str:
    """5.5 USD -> 'five dollars, fifty cents'; whole amounts drop the cents part (tokenizer.py:669-673)."""
    one, many, c_one, c_many = _CURRENCY[code]
    units = int(amount)
    cents = int(round((amount - units) * 100))
    text = f"{cardinal_en(units)} {one if units == 1 else many}"
    if float(amount).is_integer():
        return text
    return f"{text}, {cardinal_en(cents)} {c_one if cents == 1 else c_many}"


# English abbreviation / symbol behaviour of the reference (tokenizer.py:241-263 and :407-419): "\bmr\." -> "mister" ...
_ABBREV_EN = {"mrs": "misess", "mr": "mister", "dr": "doctor", "st": "saint", "co": "company", "jr": "junior",
              "maj": "major", "gen": "general", "drs": "doctors", "rev": "reverend", "lt": "lieutenant",
              "hon": "honorable", "sgt": "sergeant", "capt": "captain", "esq": "esquire", "ltd": "limited",
              "col": "colonel", "ft": "fort"}
_ABBREV_EN_RE = [(re.compile(r"\b%s\." % k, re.IGNORECASE), v) for k, v in _ABBREV_EN.items()]
_SYMBOLS_EN = [("&", " and "), ("@", " at "), ("%", " percent "), ("#", " hash "), ("$", " dollar "), ("£", " pound "),
               ("°", " degree ")]

_COMMA_NUMBER = re.compile(r"\b\d{1,3}(,\d{3})*(\.\d+)?\b")
_DOT_NUMBER = re.compile(r"\b\d{1,3}(\.\d{3})*(\,\d+)?\b")
_DECIMAL = re.compile(r"([0-9]+[.,][0-9]+)")
_ORDINAL_EN = re.compile(r"([0-9]+)(st|nd|rd|th)")
_NUMBER = re.compile(r"[0-9]+")
_CURRENCY_RE = {"GBP": re.compile(r"((£[0-9\.\,]*[0-9]+)|([0-9\.\,]*[0-9]+£))"),
                "USD": re.compile(r"((\$[0-9\.\,]*[0-9]+)|([0-9\.\,]*[0-9]+\$))"),
                "EUR": re.compile(r"(([0-9\.\,]*[0-9]+€)|((€[0-9\.\,]*[0-9]+)))")}
_WS = re.compile(r"\s+")

_CLEANED_LANGS = {"ar", "cs", "de", "en", "es", "fr", "hu", "it", "nl", "pl", "pt", "ru", "tr", "zh", "ko"}


def expand_numbers_en(text: str) -> str:
    """Order of tokenizer.py:681-700: thousands separators, currencies (GBP, USD, EUR), decimals, ordinals, integers."""
    text = _COMMA_NUMBER.sub(lambda m: m.group(0).replace(",", ""), text)
    for code in ("GBP", "USD", "EUR"):
        def cur(m, code=code):
            try:
                return currency_en(float(re.sub(r"[^\d.]", "", m.group(0).replace(",", "."))), code)
            except ValueError:
                return m.group(0)
        text = _CURRENCY_RE[code].sub(cur, text)
    text = _DECIMAL.sub(lambda m: decimal_en(m.group(1)), text)
    text = _ORDINAL_EN.sub(lambda m: ordinal_en(int(m.group(1))), text)
    return _NUMBER.sub(lambda m: cardinal_en(int(m.group(0))), text)


def multilingual_cleaners(text: str, lang: str) -> str:
    text = text.replace('"', "")
    if lang == "tr":
        text = text.replace("İ", "i").replace("Ö", "ö").replace("Ü", "ü")
    text = text.lower()
    if lang == "en":
        text = expand_numbers_en(text)
        for rx, rep in _ABBREV_EN_RE:
            text = rx.sub(rep, text)
        for sym, rep in _SYMBOLS_EN:
            text = text.replace(sym, rep).replace("  ", " ")
        text = text.strip()
    elif lang == "ru":
        text = _COMMA_NUMBER.sub(lambda m: m.group(0).replace(",", ""), text)
    elif lang != "zh":
        text = _DOT_NUMBER.sub(lambda m: m.group(0).replace(".", ""), text)
    return _WS.sub(" ", text)


def basic_cleaners(text: str) -> str:
    return _WS.sub(" ", text.lower())


def preprocess_text(text: str, lang: str) -> str:
    """tokenizer.py:805-819 (without the zh/ko/ja transliteration step)."""
    base = lang.split("-")[0]
    if base in _CLEANED_LANGS:
        return multilingual_cleaners(text, base)
    return basic_cleaners(text)


def format_for_bpe(text: str, lang: str) -> str:
    """tokenizer.py:913-917: language tag in front, spaces as the [SPACE] token."""
    base = lang.split("-")[0]
    code = "zh-cn" if base == "zh" else base
    return f"[{code}]{preprocess_text(text, lang)}".replace(" ", "[SPACE]")
