"""Per-language text cleaning in front of the BPE (the first "next" row of the hot path).

Restates `multilingual_cleaners` / `preprocess_text` of the reference
(`/root/reference/src/auralis/models/xttsv2/config/tokenizer.py:238-719,805-819`):

    drop '"'  ->  lowercase  ->  numbers  ->  abbreviations  ->  symbols  ->  collapse whitespace

What is pinned and how (tests/test_text_golden.py, vectors in tests/golden/text_cleaners.json):
  * the orchestration — separator stripping, currency / decimal / ordinal / integer passes and their order, the
    trimming of integer currency amounts, the abbreviation and symbol tables of all 15 cleaned languages — is compared
    with the reference's own functions, imported unmodified, with `num2words` replaced by a marker function on both sides;
  * the number WORDS come from the third-party `num2words` package (not installed here, not vendored by the
    reference).  `verbalise()` restates its conventions for en / es / fr / de / it / pt (the languages of
    BASELINE.json's multilingual config that use digits) and, in `numwords_more.py`, for nl / tr / hu / ru / pl / cs (cardinals,
    fractions, amounts of money, Dutch / Turkish / Hungarian / Russian ordinals); there is no copy of num2words to check against, so the word lists are "parity
    unpinned" (DESIGN.md §7).  For ar / ko digits are kept.
  * zh number normalisation is the reference's own `zh_num2words.TextNorm`: restated in `zh_textnorm.py` and pinned against
    that module (golden records + live fuzz, tests/test_zh_textnorm.py).  The zh / ja / ko romanisation is
    third-party (pypinyin, cutlet, hangul_romanize): it is called exactly as the reference calls it when the package is
    installed, otherwise the text passes through unromanised with a one-time warning.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Optional, Tuple

from . import numwords_more as _MORE

# ---------------------------------------------------------------------------------------------------------------
# tables (tokenizer.py:241-398 abbreviations, :407-594 symbols, :603-618 ordinal suffixes, :650-665 unit separators)
# ---------------------------------------------------------------------------------------------------------------
_ABBREVIATIONS: Dict[str, List[Tuple[str, str]]] = {
    "en": [("mrs", "misess"), ("mr", "mister"), ("dr", "doctor"), ("st", "saint"), ("co", "company"), ("jr", "junior"),
           ("maj", "major"), ("gen", "general"), ("drs", "doctors"), ("rev", "reverend"), ("lt", "lieutenant"),
           ("hon", "honorable"), ("sgt", "sergeant"), ("capt", "captain"), ("esq", "esquire"), ("ltd", "limited"),
           ("col", "colonel"), ("ft", "fort")],
    "es": [("sra", "señora"), ("sr", "señor"), ("dr", "doctor"), ("dra", "doctora"), ("st", "santo"), ("co", "compañía"),
           ("jr", "junior"), ("ltd", "limitada")],
    "fr": [("mme", "madame"), ("mr", "monsieur"), ("dr", "docteur"), ("st", "saint"), ("co", "compagnie"), ("jr", "junior"),
           ("ltd", "limitée")],
    "de": [("fr", "frau"), ("dr", "doktor"), ("st", "sankt"), ("co", "firma"), ("jr", "junior")],
    "pt": [("sra", "senhora"), ("sr", "senhor"), ("dr", "doutor"), ("dra", "doutora"), ("st", "santo"), ("co", "companhia"),
           ("jr", "júnior"), ("ltd", "limitada")],
    "it": [("sig", "signore"), ("dr", "dottore"), ("st", "santo"), ("co", "compagnia"), ("jr", "junior"), ("ltd", "limitata")],
    "pl": [("p", "pani"), ("m", "pan"), ("dr", "doktor"), ("sw", "święty"), ("jr", "junior")],
    "ar": [], "zh": [], "ko": [],
    "cs": [("dr", "doktor"), ("ing", "inženýr"), ("p", "pan")],
    "ru": [("г-жа", "госпожа"), ("г-н", "господин"), ("д-р", "доктор")],
    "nl": [("dhr", "de heer"), ("mevr", "mevrouw"), ("dr", "dokter"), ("jhr", "jonkheer")],
    "tr": [("b", "bay"), ("byk", "büyük"), ("dr", "doktor")],
    "hu": [("dr", "doktor"), ("b", "bácsi"), ("nőv", "nővér")],
}
# every language matches "\bKEY\." except Russian, whose keys end at a word boundary (tokenizer.py:358-365)
_ABBREV_RE = {lang: [(re.compile((r"\b%s\b" if lang == "ru" else r"\b%s\.") % k, re.IGNORECASE), v) for k, v in rows]
              for lang, rows in _ABBREVIATIONS.items()}

_SYMBOL_ORDER = ["&", "@", "%", "#", "$", "£", "°"]
_SYMBOL_WORDS: Dict[str, List[str]] = {
    "en": ["and", "at", "percent", "hash", "dollar", "pound", "degree"],
    "es": ["y", "arroba", "por ciento", "numeral", "dolar", "libra", "grados"],
    "fr": ["et", "arobase", "pour cent", "dièse", "dollar", "livre", "degrés"],
    "de": ["und", "at", "prozent", "raute", "dollar", "pfund", "grad"],
    "pt": ["e", "arroba", "por cento", "cardinal", "dólar", "libra", "graus"],
    "it": ["e", "chiocciola", "per cento", "cancelletto", "dollaro", "sterlina", "gradi"],
    "pl": ["i", "małpa", "procent", "krzyżyk", "dolar", "funt", "stopnie"],
    "ar": ["و", "على", "في المئة", "رقم", "دولار", "جنيه", "درجة"],
    "zh": ["和", "在", "百分之", "号", "美元", "英镑", "度"],
    "cs": ["a", "na", "procento", "křížek", "dolar", "libra", "stupně"],
    "ru": ["и", "собака", "процентов", "номер", "доллар", "фунт", "градус"],
    "nl": ["en", "bij", "procent", "hekje", "dollar", "pond", "graden"],
    "tr": ["ve", "at", "yüzde", "diyez", "dolar", "sterlin", "derece"],
    "hu": ["és", "kukac", "százalék", "kettőskereszt", "dollár", "font", "fok"],
    "ko": ["그리고", "에", "퍼센트", "번호", "달러", "파운드", "도"],
}

_ORDINAL_RE = {
    "en": re.compile(r"([0-9]+)(st|nd|rd|th)"),
    "es": re.compile(r"([0-9]+)(º|ª|er|o|a|os|as)"),
    "fr": re.compile(r"([0-9]+)(º|ª|er|re|e|ème)"),
    "de": re.compile(r"([0-9]+)(st|nd|rd|th|º|ª|\.(?=\s|$))"),
    "pt": re.compile(r"([0-9]+)(º|ª|o|a|os|as)"),
    "it": re.compile(r"([0-9]+)(º|°|ª|o|a|i|e)"),
    "pl": re.compile(r"([0-9]+)(º|ª|st|nd|rd|th)"),
    "ar": re.compile(r"([0-9]+)(ون|ين|ث|ر|ى)"),
    "cs": re.compile(r"([0-9]+)\.(?=\s|$)"),
    "ru": re.compile(r"([0-9]+)(-й|-я|-е|-ое|-ье|-го)"),
    "nl": re.compile(r"([0-9]+)(de|ste|e)"),
    "tr": re.compile(r"([0-9]+)(\.|inci|nci|uncu|üncü|\.)"),
    "hu": re.compile(r"([0-9]+)(\.|adik|edik|odik|edik|ödik|ödike|ik)"),
    "ko": re.compile(r"([0-9]+)(번째|번|차|째)"),
}
_NUMBER_RE = re.compile(r"[0-9]+")
_CURRENCY_RE = {"USD": re.compile(r"((\$[0-9\.\,]*[0-9]+)|([0-9\.\,]*[0-9]+\$))"),
                "GBP": re.compile(r"((£[0-9\.\,]*[0-9]+)|([0-9\.\,]*[0-9]+£))"),
                "EUR": re.compile(r"(([0-9\.\,]*[0-9]+€)|((€[0-9\.\,]*[0-9]+)))")}
_COMMA_NUMBER_RE = re.compile(r"\b\d{1,3}(,\d{3})*(\.\d+)?\b")
_DOT_NUMBER_RE = re.compile(r"\b\d{1,3}(\.\d{3})*(\,\d+)?\b")
_DECIMAL_RE = re.compile(r"([0-9]+[.,][0-9]+)")
_WS = re.compile(r"\s+")

# separator between the unit and sub-unit part of a verbalised amount; an integer amount is cut at its LAST occurrence
_AND_EQUIVALENTS = {"en": ", ", "es": " con ", "fr": " et ", "de": " und ", "pt": " e ", "it": " e ", "pl": ", ",
                    "cs": ", ", "ru": ", ", "nl": ", ", "ar": ", ", "tr": ", ", "hu": ", ", "ko": ", "}

_CLEANED_LANGS = {"ar", "cs", "de", "en", "es", "fr", "hu", "it", "nl", "pl", "pt", "ru", "tr", "zh", "ko"}

# `num2words(number, ordinal=False, lang="en", to="cardinal", currency=...)`-shaped callable
Verbaliser = Callable[..., str]


# ---------------------------------------------------------------------------------------------------------------
# number words (num2words conventions restated; see the module docstring for what is and is not pinned)
# ---------------------------------------------------------------------------------------------------------------
_EN_ONES = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
            "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_EN_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]
_EN_SCALES = [(10 ** 18, "quintillion"), (10 ** 15, "quadrillion"), (10 ** 12, "trillion"), (10 ** 9, "billion"),
              (10 ** 6, "million"), (10 ** 3, "thousand")]
_EN_ORD_IRREGULAR = {"one": "first", "two": "second", "three": "third", "five": "fifth", "eight": "eighth",
                     "nine": "ninth", "twelve": "twelfth"}


def _en_below_1000(n: int) -> str:
    parts = []
    if n >= 100:
        parts.append(_EN_ONES[n // 100] + " hundred")
        n %= 100
        if n:
            parts.append("and")
    if n >= 20:
        parts.append(_EN_TENS[n // 10] + ("-" + _EN_ONES[n % 10] if n % 10 else ""))
    elif n or not parts:
        parts.append(_EN_ONES[n])
    return " ".join(parts)


def cardinal_en(n: int) -> str:
    """1234 -> 'one thousand, two hundred and thirty-four'; 1001 -> 'one thousand and one'."""
    if n < 0:
        return "minus " + cardinal_en(-n)
    if n < 1000:
        return _en_below_1000(n)
    groups: List[str] = []
    rest = n
    for scale, name in _EN_SCALES:
        if rest >= scale:
            q, rest = divmod(rest, scale)
            groups.append(f"{cardinal_en(q)} {name}")
    text = ", ".join(groups)
    if rest:
        # a trailing group below one hundred is joined with "and", anything else with a comma
        text += (" and " if rest < 100 else ", ") + _en_below_1000(rest)
    return text


def ordinal_en(n: int) -> str:
    words = cardinal_en(n)
    head, sep, last = words.rpartition(" ")
    pre, hy, tail = last.rpartition("-")
    if tail in _EN_ORD_IRREGULAR:
        tail = _EN_ORD_IRREGULAR[tail]
    elif tail.endswith("y"):
        tail = tail[:-1] + "ieth"
    else:
        tail += "th"
    return head + sep + pre + hy + tail


# --- Spanish
_ES_0_29 = ["cero", "uno", "dos", "tres", "cuatro", "cinco", "seis", "siete", "ocho", "nueve", "diez", "once", "doce", "trece",
            "catorce", "quince", "dieciséis", "diecisiete", "dieciocho", "diecinueve", "veinte", "veintiuno", "veintidós",
            "veintitrés", "veinticuatro", "veinticinco", "veintiséis", "veintisiete", "veintiocho", "veintinueve"]
_ES_TENS = ["", "", "", "treinta", "cuarenta", "cincuenta", "sesenta", "setenta", "ochenta", "noventa"]
_ES_HUNDREDS = ["", "ciento", "doscientos", "trescientos", "cuatrocientos", "quinientos", "seiscientos", "setecientos",
                "ochocientos", "novecientos"]


def _es_below_1000(n: int) -> str:
    if n == 100:
        return "cien"
    parts = []
    if n >= 100:
        parts.append(_ES_HUNDREDS[n // 100])
        n %= 100
    if n >= 30:
        parts.append(_ES_TENS[n // 10] + (" y " + _ES_0_29[n % 10] if n % 10 else ""))
    elif n or not parts:
        parts.append(_ES_0_29[n])
    return " ".join(parts)


def _es_apocope(words: str) -> str:
    """'uno' in front of a noun or a scale word loses its o: veintiún mil, un millón, treinta y un euros."""
    if words.endswith("veintiuno"):
        return words[:-len("veintiuno")] + "veintiún"
    if words == "uno" or words.endswith(" uno"):
        return words[:-1]
    return words


def cardinal_es(n: int) -> str:
    if n < 0:
        return "menos " + cardinal_es(-n)
    if n < 1000:
        return _es_below_1000(n)
    if n < 10 ** 6:
        q, r = divmod(n, 1000)
        head = "mil" if q == 1 else _es_apocope(_es_below_1000(q)) + " mil"
        return head + (" " + _es_below_1000(r) if r else "")
    if n < 10 ** 12:
        q, r = divmod(n, 10 ** 6)
        head = "un millón" if q == 1 else _es_apocope(cardinal_es(q)) + " millones"
        return head + (" " + cardinal_es(r) if r else "")
    q, r = divmod(n, 10 ** 12)
    head = "un billón" if q == 1 else _es_apocope(cardinal_es(q)) + " billones"
    return head + (" " + cardinal_es(r) if r else "")


_ES_ORD_UNITS = ["", "primero", "segundo", "tercero", "cuarto", "quinto", "sexto", "séptimo", "octavo", "noveno"]
_ES_ORD_TENS = ["", "décimo", "vigésimo", "trigésimo", "cuadragésimo", "quincuagésimo", "sexagésimo", "septuagésimo",
                "octogésimo", "nonagésimo"]
_ES_ORD_HUNDREDS = ["", "centésimo", "ducentésimo", "tricentésimo", "cuadringentésimo", "quingentésimo", "sexcentésimo",
                    "septingentésimo", "octingentésimo", "noningentésimo"]


def ordinal_es(n: int) -> str:
    if n <= 0 or n >= 1000:
        return cardinal_es(n)
    if n == 11:
        return "undécimo"
    if n == 12:
        return "duodécimo"
    if 13 <= n <= 19:
        return "decimo" + _ES_ORD_UNITS[n - 10]
    parts = [_ES_ORD_HUNDREDS[n // 100], _ES_ORD_TENS[(n // 10) % 10], _ES_ORD_UNITS[n % 10]]
    return " ".join(p for p in parts if p)


# --- French
_FR_0_16 = ["zéro", "un", "deux", "trois", "quatre", "cinq", "six", "sept", "huit", "neuf", "dix", "onze", "douze", "treize",
            "quatorze", "quinze", "seize"]
_FR_TENS = ["", "dix", "vingt", "trente", "quarante", "cinquante", "soixante"]


def _fr_below_100(n: int) -> str:
    if n <= 16:
        return _FR_0_16[n]
    if n < 20:
        return "dix-" + _FR_0_16[n - 10]
    if n < 70:
        t, u = divmod(n, 10)
        if u == 0:
            return _FR_TENS[t]
        return _FR_TENS[t] + (" et un" if u == 1 else "-" + _FR_0_16[u])
    if n < 80:
        return "soixante et onze" if n == 71 else "soixante-" + _fr_below_100(n - 60)
    if n == 80:
        return "quatre-vingts"
    return "quatre-vingt-" + _fr_below_100(n - 80)


def _fr_below_1000(n: int, final: bool = True) -> str:
    """`final`: the group ends the number, so 'cents' / 'quatre-vingts' keep their plural s."""
    h, r = divmod(n, 100)
    if h == 0:
        text = _fr_below_100(r)
    elif r == 0:
        text = "cent" if h == 1 else _FR_0_16[h] + " cents"
    else:
        text = ("cent " if h == 1 else _FR_0_16[h] + " cent ") + _fr_below_100(r)
    if not final and text.endswith(("cents", "vingts")):
        text = text[:-1]
    return text


def cardinal_fr(n: int) -> str:
    if n < 0:
        return "moins " + cardinal_fr(-n)
    if n < 1000:
        return _fr_below_1000(n)
    if n < 10 ** 6:
        q, r = divmod(n, 1000)
        head = "mille" if q == 1 else _fr_below_1000(q, final=False) + " mille"
        return head + (" " + _fr_below_1000(r) if r else "")
    for scale, one, many in ((10 ** 9, "milliard", "milliards"), (10 ** 6, "million", "millions")):
        if n >= scale:
            q, r = divmod(n, scale)
            head = f"un {one}" if q == 1 else f"{cardinal_fr(q)} {many}"
            return head + (" " + cardinal_fr(r) if r else "")
    return str(n)


def ordinal_fr(n: int) -> str:
    if n == 1:
        return "premier"
    w = cardinal_fr(n)
    if w.endswith(("cents", "vingts")):
        w = w[:-1]
    if w.endswith("cinq"):
        return w + "uième"
    if w.endswith("neuf"):
        return w[:-1] + "vième"
    if w.endswith("e"):
        w = w[:-1]
    return w + "ième"


# --- German
_DE_0_19 = ["null", "eins", "zwei", "drei", "vier", "fünf", "sechs", "sieben", "acht", "neun", "zehn", "elf", "zwölf",
            "dreizehn", "vierzehn", "fünfzehn", "sechzehn", "siebzehn", "achtzehn", "neunzehn"]
_DE_TENS = ["", "", "zwanzig", "dreißig", "vierzig", "fünfzig", "sechzig", "siebzig", "achtzig", "neunzig"]


def _de_below_100(n: int, bare_one: bool = False) -> str:
    if n == 1 and bare_one:
        return "ein"
    if n < 20:
        return _DE_0_19[n]
    t, u = divmod(n, 10)
    if u == 0:
        return _DE_TENS[t]
    return ("ein" if u == 1 else _DE_0_19[u]) + "und" + _DE_TENS[t]


def _de_below_1000(n: int, bare_one: bool = False) -> str:
    h, r = divmod(n, 100)
    text = (_de_below_100(h, True) + "hundert") if h else ""
    if r or not h:
        text += _de_below_100(r, bare_one and not h)
    return text


def cardinal_de(n: int) -> str:
    if n < 0:
        return "minus " + cardinal_de(-n)
    if n < 1000:
        return _de_below_1000(n)
    if n < 10 ** 6:
        q, r = divmod(n, 1000)
        return _de_below_1000(q, True) + "tausend" + (_de_below_1000(r) if r else "")
    for scale, one, many in ((10 ** 9, "eine milliarde", "milliarden"), (10 ** 6, "eine million", "millionen")):
        if n >= scale:
            q, r = divmod(n, scale)
            head = one if q == 1 else f"{cardinal_de(q)} {many}"
            return head + (" " + cardinal_de(r) if r else "")
    return str(n)


_DE_ORD_IRREGULAR = {"eins": "erste", "drei": "dritte", "sieben": "siebte", "acht": "achte"}


def ordinal_de(n: int) -> str:
    if n <= 0:
        return cardinal_de(n)
    w = cardinal_de(n)
    r = n % 100
    if 0 < r < 20:
        stem = _DE_0_19[r]
        return w[:len(w) - len(stem)] + _DE_ORD_IRREGULAR.get(stem, stem + "te")
    return w + "ste"


# --- Italian
_IT_0_19 = ["zero", "uno", "due", "tre", "quattro", "cinque", "sei", "sette", "otto", "nove", "dieci", "undici", "dodici",
            "tredici", "quattordici", "quindici", "sedici", "diciassette", "diciotto", "diciannove"]
_IT_TENS = ["", "", "venti", "trenta", "quaranta", "cinquanta", "sessanta", "settanta", "ottanta", "novanta"]


def _it_below_100(n: int) -> str:
    if n < 20:
        return _IT_0_19[n]
    t, u = divmod(n, 10)
    tens = _IT_TENS[t]
    if u == 0:
        return tens
    if u in (1, 8):
        tens = tens[:-1]                       # ventuno, ventotto
    return tens + ("tré" if u == 3 else _IT_0_19[u])


def _it_below_1000(n: int) -> str:
    h, r = divmod(n, 100)
    text = ("" if h == 1 else _IT_0_19[h]) + "cento" if h else ""
    if h and 80 <= r < 90:
        text = text[:-1]                       # centottanta
    if r or not h:
        text += _it_below_100(r)
    return text


def cardinal_it(n: int) -> str:
    if n < 0:
        return "meno " + cardinal_it(-n)
    if n < 1000:
        return _it_below_1000(n)
    if n < 10 ** 6:
        q, r = divmod(n, 1000)
        return ("mille" if q == 1 else _it_below_1000(q) + "mila") + (_it_below_1000(r) if r else "")
    for scale, one, many in ((10 ** 9, "un miliardo", "miliardi"), (10 ** 6, "un milione", "milioni")):
        if n >= scale:
            q, r = divmod(n, scale)
            head = one if q == 1 else f"{cardinal_it(q)} {many}"
            return head + (" e " + cardinal_it(r) if r else "")
    return str(n)


_IT_ORD_1_10 = ["", "primo", "secondo", "terzo", "quarto", "quinto", "sesto", "settimo", "ottavo", "nono", "decimo"]


def ordinal_it(n: int) -> str:
    if n <= 0:
        return cardinal_it(n)
    if n <= 10:
        return _IT_ORD_1_10[n]
    w = cardinal_it(n)
    if w.endswith("tré"):
        return w[:-1] + "eesimo"
    if w.endswith("sei"):
        return w + "esimo"
    return w[:-1] + "esimo"


# --- Portuguese (num2words "pt" is European Portuguese: dezasseis, dezassete, dezanove)
_PT_0_19 = ["zero", "um", "dois", "três", "quatro", "cinco", "seis", "sete", "oito", "nove", "dez", "onze", "doze", "treze",
            "catorze", "quinze", "dezasseis", "dezassete", "dezoito", "dezanove"]
_PT_TENS = ["", "", "vinte", "trinta", "quarenta", "cinquenta", "sessenta", "setenta", "oitenta", "noventa"]
_PT_HUNDREDS = ["", "cento", "duzentos", "trezentos", "quatrocentos", "quinhentos", "seiscentos", "setecentos", "oitocentos",
                "novecentos"]


def _pt_below_1000(n: int) -> str:
    if n == 100:
        return "cem"
    parts = []
    if n >= 100:
        parts.append(_PT_HUNDREDS[n // 100])
        n %= 100
    if n >= 20:
        parts.append(_PT_TENS[n // 10])
        if n % 10:
            parts.append(_PT_0_19[n % 10])
    elif n or not parts:
        parts.append(_PT_0_19[n])
    return " e ".join(parts)


def _pt_join(head: str, r: int, tail: str) -> str:
    """'e' links a remainder that is below 100 or a round hundred: mil e um, mil e cem, mil duzentos e trinta."""
    if not r:
        return head
    return head + (" e " if (r < 100 or r % 100 == 0) else " ") + tail


def cardinal_pt(n: int) -> str:
    if n < 0:
        return "menos " + cardinal_pt(-n)
    if n < 1000:
        return _pt_below_1000(n)
    if n < 10 ** 6:
        q, r = divmod(n, 1000)
        return _pt_join("mil" if q == 1 else _pt_below_1000(q) + " mil", r, _pt_below_1000(r))
    for scale, one, many in ((10 ** 9, "mil milhões", "mil milhões"), (10 ** 6, "um milhão", "milhões")):
        if n >= scale:
            q, r = divmod(n, scale)
            head = one if q == 1 else f"{cardinal_pt(q)} {many}"
            return _pt_join(head, r, cardinal_pt(r))
    return str(n)


_PT_ORD_UNITS = ["", "primeiro", "segundo", "terceiro", "quarto", "quinto", "sexto", "sétimo", "oitavo", "nono"]
_PT_ORD_TENS = ["", "décimo", "vigésimo", "trigésimo", "quadragésimo", "quinquagésimo", "sexagésimo", "septuagésimo",
                "octogésimo", "nonagésimo"]
_PT_ORD_HUNDREDS = ["", "centésimo", "ducentésimo", "tricentésimo", "quadringentésimo", "quingentésimo", "seiscentésimo",
                    "septingentésimo", "octingentésimo", "noningentésimo"]


def ordinal_pt(n: int) -> str:
    if n <= 0 or n >= 1000:
        return cardinal_pt(n)
    parts = [_PT_ORD_HUNDREDS[n // 100], _PT_ORD_TENS[(n // 10) % 10], _PT_ORD_UNITS[n % 10]]
    return " ".join(p for p in parts if p)


_CARDINAL = {"en": cardinal_en, "es": cardinal_es, "fr": cardinal_fr, "de": cardinal_de, "it": cardinal_it, "pt": cardinal_pt}
_ORDINAL = {"en": ordinal_en, "es": ordinal_es, "fr": ordinal_fr, "de": ordinal_de, "it": ordinal_it, "pt": ordinal_pt}
_POINT = {"en": "point", "es": "punto", "fr": "virgule", "de": "komma", "it": "virgola", "pt": "vírgula"}
# (unit singular, unit plural, sub-unit singular, sub-unit plural)
_CURRENCY_WORDS = {
    "en": {"USD": ("dollar", "dollars", "cent", "cents"), "GBP": ("pound", "pounds", "penny", "pence"),
           "EUR": ("euro", "euros", "cent", "cents")},
    "es": {"USD": ("dólar", "dólares", "centavo", "centavos"), "GBP": ("libra", "libras", "penique", "peniques"),
           "EUR": ("euro", "euros", "céntimo", "céntimos")},
    "fr": {"USD": ("dollar", "dollars", "cent", "cents"), "GBP": ("livre", "livres", "penny", "pence"),
           "EUR": ("euro", "euros", "centime", "centimes")},
    "de": {"USD": ("dollar", "dollar", "cent", "cent"), "GBP": ("pfund", "pfund", "penny", "pence"),
           "EUR": ("euro", "euro", "cent", "cent")},
    "it": {"USD": ("dollaro", "dollari", "centesimo", "centesimi"), "GBP": ("sterlina", "sterline", "penny", "penny"),
           "EUR": ("euro", "euro", "centesimo", "centesimi")},
    "pt": {"USD": ("dólar", "dólares", "cêntimo", "cêntimos"), "GBP": ("libra", "libras", "péni", "pence"),
           "EUR": ("euro", "euros", "cêntimo", "cêntimos")},
}
_ONE_BEFORE_NOUN = {"es": "un", "de": "ein", "it": "un"}      # 'uno euro' -> 'un euro', 'eins euro' -> 'ein euro'

VERBALISED_LANGS = frozenset(_CARDINAL)


def decimal_words(value: float, lang: str) -> str:
    """3.05 -> 'three point zero five' (the fraction is read digit by digit, after Python's float repr, as num2words does)."""
    whole, _, frac = repr(float(value)).partition(".")
    if "e" in frac or "e" in whole:
        return _CARDINAL[lang](int(value))
    card = _CARDINAL[lang]
    return f"{card(int(whole))} {_POINT[lang]} " + " ".join(card(int(d)) for d in frac)


def currency_words(amount: float, code: str, lang: str) -> str:
    """5.5 USD -> 'five dollars, fifty cents' — always both parts, like num2words; integer amounts are trimmed by the caller."""
    one, many, c_one, c_many = _CURRENCY_WORDS[lang][code]
    units = int(amount)
    cents = int(round((amount - units) * 100))
    card = _CARDINAL[lang]

    def count(n: int) -> str:
        if n == 1 and lang in _ONE_BEFORE_NOUN:
            return _ONE_BEFORE_NOUN[lang]
        w = card(n)
        return _es_apocope(w) if lang == "es" else w
    return f"{count(units)} {one if units == 1 else many}{_AND_EQUIVALENTS[lang]}{count(cents)} {c_one if cents == 1 else c_many}"


def verbalise(number, ordinal: bool = False, lang: str = "en", to: str = "cardinal", **kwargs) -> str:
    """`num2words`-shaped entry point over the restated languages.  Languages without a restatement keep their digits
    (for an ordinal or an amount: the digits of the number alone)."""
    if lang == "cz":
        lang = "cs"
    if lang in _MORE.CARDINAL:                 # nl, tr, hu, ru, pl, cs: numwords_more.py (cardinals, fractions, some ordinals)
        if to == "currency":
            return _MORE.currency_words(float(number), kwargs.get("currency", "EUR"), lang)
        if ordinal or to == "ordinal":
            return _MORE.ORDINAL[lang](int(number)) if lang in _MORE.ORDINAL else str(number)
        return _MORE.decimal_words(number, lang) if isinstance(number, float) else _MORE.CARDINAL[lang](int(number))
    if lang not in VERBALISED_LANGS:
        if to == "currency":
            return repr(float(number)) if not float(number).is_integer() else str(int(number))
        return str(number)
    if to == "currency":
        return currency_words(float(number), kwargs.get("currency", "EUR"), lang)
    if ordinal or to == "ordinal":
        return _ORDINAL[lang](int(number))
    if isinstance(number, float):
        return decimal_words(number, lang)
    return _CARDINAL[lang](int(number))


# ---------------------------------------------------------------------------------------------------------------
# the cleaner pipeline
# ---------------------------------------------------------------------------------------------------------------
def expand_abbreviations_multilingual(text: str, lang: str = "en") -> str:
    """tokenizer.py:401-405."""
    for rx, rep in _ABBREV_RE.get(lang, ()):
        text = rx.sub(rep, text)
    return text


def expand_symbols_multilingual(text: str, lang: str = "en") -> str:
    """tokenizer.py:596-601: ' word ' for each symbol, double spaces squeezed after every replacement, strip."""
    words = _SYMBOL_WORDS.get(lang)
    if words:
        for sym, w in zip(_SYMBOL_ORDER, words):
            text = text.replace(sym, f" {w} ")
            text = text.replace("  ", " ")
    return text.strip()


def _expand_currency(m, lang: str, code: str, n2w: Verbaliser) -> str:
    """tokenizer.py:647-673."""
    amount = float(re.sub(r"[^\d.]", "", m.group(0).replace(",", ".")))
    full = n2w(amount, to="currency", currency=code, lang=lang if lang != "cs" else "cz")
    if amount.is_integer():
        cut = full.rfind(_AND_EQUIVALENTS.get(lang, ", "))
        if cut != -1:
            full = full[:cut]
    return full


def expand_numbers_multilingual(text: str, lang: str = "en", n2w: Optional[Verbaliser] = None) -> str:
    """tokenizer.py:681-700.  `n2w` defaults to `verbalise`; tests inject the marker function shared with the reference."""
    if lang == "zh":
        from .zh_textnorm import normalize as zh_normalize
        return zh_normalize(text)       # the reference's own zh_num2words.TextNorm, restated and pinned (zh_textnorm.py)
    n2w = n2w or verbalise
    n2w_lang = lang if lang != "cs" else "cz"
    if lang in ("en", "ru"):
        text = _COMMA_NUMBER_RE.sub(lambda m: m.group(0).replace(",", ""), text)
    else:
        text = _DOT_NUMBER_RE.sub(lambda m: m.group(0).replace(".", ""), text)
    try:                                # a failure in one currency pass skips the remaining ones (tokenizer.py:688-693)
        for code in ("GBP", "USD", "EUR"):
            text = _CURRENCY_RE[code].sub(lambda m, code=code: _expand_currency(m, lang, code, n2w), text)
    except Exception:
        pass
    if lang != "tr":
        text = _DECIMAL_RE.sub(lambda m: n2w(float(m.group(1).replace(",", ".")), lang=n2w_lang), text)
    if lang in _ORDINAL_RE:
        text = _ORDINAL_RE[lang].sub(lambda m: n2w(int(m.group(1)), ordinal=True, lang=n2w_lang), text)
    return _NUMBER_RE.sub(lambda m: n2w(int(m.group(0)), lang=n2w_lang), text)


def multilingual_cleaners(text: str, lang: str, n2w: Optional[Verbaliser] = None) -> str:
    """tokenizer.py:708-719."""
    text = text.replace('"', "")
    if lang == "tr":
        text = text.replace("İ", "i").replace("Ö", "ö").replace("Ü", "ü")
    text = text.lower()
    text = expand_numbers_multilingual(text, lang, n2w)
    text = expand_abbreviations_multilingual(text, lang)
    text = expand_symbols_multilingual(text, lang)
    return _WS.sub(" ", text)


def basic_cleaners(text: str) -> str:
    return _WS.sub(" ", text.lower())


_TRANSLIT_CACHE: Dict[str, object] = {}
_TRANSLIT_WARNED = set()


def _transliterator(lang: str):
    """zh / ko / ja romanisation exactly as the reference wires it (tokenizer.py:727-739,797-803) when the third-party
    package it uses is installed: pypinyin (TONE3, neutral tone 5), hangul_romanize (academic rule), cutlet (romaji, then
    lowercase).  None when the package is absent — the text then passes through unromanised and a warning says so once."""
    if lang in _TRANSLIT_CACHE:
        return _TRANSLIT_CACHE[lang]
    fn = None
    try:
        if lang == "zh":
            import pypinyin
            fn = lambda t: "".join(p[0] for p in pypinyin.pinyin(t, style=pypinyin.Style.TONE3, heteronym=False,      # noqa: E731
                                                                  neutral_tone_with_five=True))
        elif lang == "ko":
            from hangul_romanize import Transliter
            from hangul_romanize.rule import academic
            tr = Transliter(academic)
            fn = tr.translit
        elif lang == "ja":
            import cutlet
            katsu = cutlet.Cutlet()
            fn = lambda t: katsu.romaji(t).lower()                                                                     # noqa: E731
    except ImportError:
        fn = None
    if fn is None and lang not in _TRANSLIT_WARNED:
        _TRANSLIT_WARNED.add(lang)
        import warnings
        pkg = {"zh": "pypinyin", "ko": "hangul_romanize", "ja": "cutlet"}[lang]
        warnings.warn(f"text front-end: {pkg} is not installed, '{lang}' text is passed to the BPE without romanisation "
                      f"(the reference romanises it, tokenizer.py:805-819)", RuntimeWarning, stacklevel=3)
    _TRANSLIT_CACHE[lang] = fn
    return fn


def preprocess_text(text: str, lang: str, n2w: Optional[Verbaliser] = None) -> str:
    """tokenizer.py:805-819.  zh / ko / ja romanisation runs when its third-party package is importable (`_transliterator`)."""
    base = lang.split("-")[0]
    if base in _CLEANED_LANGS:
        text = multilingual_cleaners(text, base, n2w)
        if base in ("zh", "ko"):
            fn = _transliterator(base)
            if fn is not None:
                text = fn(text)
        return text
    if base == "ja":
        fn = _transliterator("ja")
        if fn is not None:
            return fn(text)                       # japanese_cleaners: romaji, then lowercase (tokenizer.py:732-735)
    return basic_cleaners(text)


def format_for_bpe(text: str, lang: str) -> str:
    """tokenizer.py:913-917: language tag in front, spaces as the [SPACE] token."""
    base = lang.split("-")[0]
    code = "zh-cn" if base == "zh" else base
    return f"[{code}]{preprocess_text(text, lang)}".replace(" ", "[SPACE]")
