"""Number words for Dutch, Turkish, Hungarian, Russian, Polish and Czech — the languages for which the reference's cleaners call
the third-party `num2words` (tokenizer.py:643-679) and which `textnorm.py` did not restate in round 1.

PARITY UNPINNED: `num2words` is not installed in this image and not vendored by the reference, so there is nothing to run these
against; they restate the library's published conventions (word lists, plural rules, where "one" is dropped, how a fraction is
read) from the languages' grammar.  What IS pinned is everything around them (which spans are numbers, ordinals, amounts, in which
order they are expanded): `tests/test_text_golden.py` runs our cleaners and the reference's with the same marker verbaliser.

Cardinals for all six; ordinals for Dutch, Turkish, Hungarian and (masculine nominative) Russian; fractions: Dutch digit by digit after "komma" (the base class's reading),
Russian / Polish / Czech as an integer after "запятая" / "przecinek" / "celá" with leading zeros spoken (those three modules
share that code), Hungarian "egész ... tized / század / ezred"; amounts of money as "<amount> <unit>, <cents> <sub-unit>".  Polish and
Czech ordinals keep their digits (and are then read as cardinals)."""
from typing import Callable, Dict, List, Tuple


# --------------------------------------------------------------------------------------------------- Dutch
_NL_LOW = ["nul", "een", "twee", "drie", "vier", "vijf", "zes", "zeven", "acht", "negen", "tien", "elf", "twaalf", "dertien",
           "veertien", "vijftien", "zestien", "zeventien", "achttien", "negentien"]
_NL_TENS = ["", "", "twintig", "dertig", "veertig", "vijftig", "zestig", "zeventig", "tachtig", "negentig"]
_NL_BIG = [(10 ** 12, "biljoen"), (10 ** 9, "miljard"), (10 ** 6, "miljoen")]


def _nl_below_100(n: int) -> str:
    if n < 20:
        return _NL_LOW[n]
    t, u = divmod(n, 10)
    if u == 0:
        return _NL_TENS[t]
    unit = _NL_LOW[u]
    return unit + ("ën" if unit.endswith("e") else "en") + _NL_TENS[t]          # tweeëntwintig, eenentwintig


def _nl_below_1000(n: int) -> str:
    h, r = divmod(n, 100)
    head = "" if h == 0 else ("honderd" if h == 1 else _NL_LOW[h] + "honderd")
    return head + (_nl_below_100(r) if r or not head else "")


def cardinal_nl(n: int) -> str:
    """1234 -> 'duizendtweehonderdvierendertig', 2500000 -> 'twee miljoen vijfhonderdduizend'."""
    if n < 0:
        return "min " + cardinal_nl(-n)
    parts: List[str] = []
    rest = n
    for scale, name in _NL_BIG:
        if rest >= scale:
            q, rest = divmod(rest, scale)
            parts.append(f"{cardinal_nl(q)} {name}")
    low = ""
    th, r = divmod(rest, 1000)
    if th:
        low = ("" if th == 1 else _nl_below_1000(th)) + "duizend"
    if r or (not low and not parts):
        low += _nl_below_1000(r)
    if low:
        parts.append(low)
    return " ".join(parts)


_NL_ORD_ENDINGS = [("nul", "nuld"), ("een", "eerst"), ("twee", "tweed"), ("drie", "derd"), ("vier", "vierd"), ("vijf", "vijfd"),
                   ("zes", "zesd"), ("zeven", "zevend"), ("acht", "achtst"), ("negen", "negend"), ("tien", "tiend"),
                   ("elf", "elfd"), ("twaalf", "twaalfd"), ("ig", "igst"), ("erd", "erdst"), ("end", "endst"),
                   ("joen", "joenst"), ("rd", "rdst")]


def ordinal_nl(n: int) -> str:
    word = cardinal_nl(n)
    for key, repl in _NL_ORD_ENDINGS:
        if word.endswith(key):
            word = word[: len(word) - len(key)] + repl
            break
    return word + "e"


# --------------------------------------------------------------------------------------------------- Turkish (written closed up)
_TR_ONES = ["", "bir", "iki", "üç", "dört", "beş", "altı", "yedi", "sekiz", "dokuz"]
_TR_TENS = ["", "on", "yirmi", "otuz", "kırk", "elli", "altmış", "yetmiş", "seksen", "doksan"]
_TR_BIG = [(10 ** 12, "trilyon"), (10 ** 9, "milyar"), (10 ** 6, "milyon")]


def _tr_below_1000(n: int) -> str:
    h, r = divmod(n, 100)
    t, u = divmod(r, 10)
    return ("" if h == 0 else ("yüz" if h == 1 else _TR_ONES[h] + "yüz")) + _TR_TENS[t] + _TR_ONES[u]


def cardinal_tr(n: int) -> str:
    """2345 -> 'ikibinüçyüzkırkbeş', 1000000 -> 'birmilyon' (no spaces, 'bir' dropped before yüz and bin only)."""
    if n < 0:
        return "eksi" + cardinal_tr(-n)
    if n == 0:
        return "sıfır"
    out = ""
    rest = n
    for scale, name in _TR_BIG:
        if rest >= scale:
            q, rest = divmod(rest, scale)
            out += cardinal_tr(q) + name
    th, r = divmod(rest, 1000)
    if th:
        out += ("" if th == 1 else _tr_below_1000(th)) + "bin"
    return out + _tr_below_1000(r)


_TR_ORD_SUFFIX = [("bir", "inci"), ("iki", "nci"), ("üç", "üncü"), ("dört", "üncü"), ("beş", "inci"), ("altı", "ncı"), ("yedi", "nci"),
                  ("sekiz", "inci"), ("dokuz", "uncu"), ("on", "uncu"), ("yirmi", "nci"), ("otuz", "uncu"), ("kırk", "ıncı"),
                  ("elli", "nci"), ("altmış", "ıncı"), ("yetmiş", "inci"), ("seksen", "inci"), ("doksan", "ıncı"), ("yüz", "üncü"),
                  ("bin", "inci"), ("milyon", "uncu"), ("milyar", "ıncı"), ("trilyon", "uncu"), ("sıfır", "ıncı")]


def ordinal_tr(n: int) -> str:
    word = cardinal_tr(n)
    if word.endswith("dört"):
        return word[:-1] + "düncü"                       # consonant softening: dört -> dördüncü
    for stem, suffix in sorted(_TR_ORD_SUFFIX, key=lambda kv: -len(kv[0])):
        if word.endswith(stem):
            return word + suffix
    return word


# --------------------------------------------------------------------------------------------------- Hungarian
_HU_ONES = ["", "egy", "kettő", "három", "négy", "öt", "hat", "hét", "nyolc", "kilenc"]
_HU_TENS_ALONE = ["", "tíz", "húsz", "harminc", "negyven", "ötven", "hatvan", "hetven", "nyolcvan", "kilencven"]
_HU_TENS_PREFIX = ["", "tizen", "huszon", "harminc", "negyven", "ötven", "hatvan", "hetven", "nyolcvan", "kilencven"]
_HU_BIG = [(10 ** 12, "billió"), (10 ** 9, "milliárd"), (10 ** 6, "millió")]


def _hu_below_1000(n: int, prefix: bool) -> str:
    """prefix=True: the group multiplies a following word ('két' + 'ezer'), so 2 is 'két'."""
    h, r = divmod(n, 100)
    t, u = divmod(r, 10)
    out = ""
    if h:
        out += ("" if h == 1 else ("két" if h == 2 else _HU_ONES[h])) + "száz"
    if t:
        out += _HU_TENS_PREFIX[t] if u else _HU_TENS_ALONE[t]
    if u:
        out += "két" if (u == 2 and prefix) else _HU_ONES[u]
    return out


def cardinal_hu(n: int) -> str:
    """1999 -> 'ezerkilencszázkilencvenkilenc', 2001 -> 'kétezer-egy' (groups are hyphenated above two thousand)."""
    if n < 0:
        return "mínusz " + cardinal_hu(-n)
    if n == 0:
        return "nulla"
    groups: List[str] = []
    rest = n
    for scale, name in _HU_BIG:
        if rest >= scale:
            q, rest = divmod(rest, scale)
            groups.append(_hu_below_1000(q, True) + name)
    th, r = divmod(rest, 1000)
    if th:
        groups.append(("" if th == 1 and not groups else _hu_below_1000(th, True)) + "ezer")
    if r:
        groups.append(_hu_below_1000(r, False))
    return ("-" if n > 2000 else "").join(groups)


_HU_FRACTION = {1: "tized", 2: "század", 3: "ezred"}
# the last component of the cardinal takes its ordinal form (longest match first); 1 and 2 alone are suppletive
_HU_ORD_LAST = [("milliárd", "milliárdodik"), ("millió", "milliomodik"), ("billió", "billiomodik"), ("kilencven", "kilencvenedik"),
                ("nyolcvan", "nyolcvanadik"), ("hetven", "hetvenedik"), ("hatvan", "hatvanadik"), ("ötven", "ötvenedik"),
                ("negyven", "negyvenedik"), ("harminc", "harmincadik"), ("húsz", "huszadik"), ("tíz", "tizedik"), ("száz", "századik"),
                ("ezer", "ezredik"), ("kilenc", "kilencedik"), ("nyolc", "nyolcadik"), ("három", "harmadik"), ("kettő", "kettedik"),
                ("négy", "negyedik"), ("egy", "egyedik"), ("hét", "hetedik"), ("hat", "hatodik"), ("öt", "ötödik")]


def ordinal_hu(n: int) -> str:
    """1 -> 'első', 2 -> 'második', 21 -> 'huszonegyedik', 100 -> 'századik', 2001 -> 'kétezer-egyedik'."""
    if n == 1:
        return "első"
    if n == 2:
        return "második"
    word = cardinal_hu(n)
    for tail, repl in _HU_ORD_LAST:
        if word.endswith(tail):
            return word[: len(word) - len(tail)] + repl
    return word


# --------------------------------------------------------------------------------------------------- Russian / Polish / Czech
# one engine, three tables: (ones, ones feminine, teens, tens, hundreds, scale names with three plural forms, plural rule,
# whether 'one' is dropped before a scale word, zero, the word between whole part and fraction)
def _plural_ru(n: int) -> int:
    if n % 100 in range(11, 20):
        return 2
    return 0 if n % 10 == 1 else (1 if n % 10 in (2, 3, 4) else 2)


def _plural_west(n: int) -> int:            # Polish, Czech
    if n == 1:
        return 0
    return 1 if (1 < n % 10 < 5 and not 10 < n % 100 < 20) else 2


_SLAVIC: Dict[str, dict] = {
    "ru": dict(
        ones=["", "один", "два", "три", "четыре", "пять", "шесть", "семь", "восемь", "девять"],
        fem={1: "одна", 2: "две"},
        teens=["десять", "одиннадцать", "двенадцать", "тринадцать", "четырнадцать", "пятнадцать", "шестнадцать", "семнадцать",
               "восемнадцать", "девятнадцать"],
        tens=["", "", "двадцать", "тридцать", "сорок", "пятьдесят", "шестьдесят", "семьдесят", "восемьдесят", "девяносто"],
        hundreds=["", "сто", "двести", "триста", "четыреста", "пятьсот", "шестьсот", "семьсот", "восемьсот", "девятьсот"],
        scales=[("тысяча", "тысячи", "тысяч"), ("миллион", "миллиона", "миллионов"), ("миллиард", "миллиарда", "миллиардов"),
                ("триллион", "триллиона", "триллионов")],
        plural=_plural_ru, drop_one=False, zero="ноль", point="запятая", minus="минус"),
    "pl": dict(
        ones=["", "jeden", "dwa", "trzy", "cztery", "pięć", "sześć", "siedem", "osiem", "dziewięć"],
        fem={},
        teens=["dziesięć", "jedenaście", "dwanaście", "trzynaście", "czternaście", "piętnaście", "szesnaście", "siedemnaście",
               "osiemnaście", "dziewiętnaście"],
        tens=["", "", "dwadzieścia", "trzydzieści", "czterdzieści", "pięćdziesiąt", "sześćdziesiąt", "siedemdziesiąt",
              "osiemdziesiąt", "dziewięćdziesiąt"],
        hundreds=["", "sto", "dwieście", "trzysta", "czterysta", "pięćset", "sześćset", "siedemset", "osiemset", "dziewięćset"],
        scales=[("tysiąc", "tysiące", "tysięcy"), ("milion", "miliony", "milionów"), ("miliard", "miliardy", "miliardów"),
                ("bilion", "biliony", "bilionów")],
        plural=_plural_west, drop_one=True, zero="zero", point="przecinek", minus="minus"),
    "cs": dict(
        ones=["", "jedna", "dva", "tři", "čtyři", "pět", "šest", "sedm", "osm", "devět"],
        fem={},
        teens=["deset", "jedenáct", "dvanáct", "třináct", "čtrnáct", "patnáct", "šestnáct", "sedmnáct", "osmnáct", "devatenáct"],
        tens=["", "", "dvacet", "třicet", "čtyřicet", "padesát", "šedesát", "sedmdesát", "osmdesát", "devadesát"],
        hundreds=["", "sto", "dvě stě", "tři sta", "čtyři sta", "pět set", "šest set", "sedm set", "osm set", "devět set"],
        scales=[("tisíc", "tisíce", "tisíc"), ("milion", "miliony", "milionů"), ("miliarda", "miliardy", "miliard"),
                ("bilion", "biliony", "bilionů")],
        plural=_plural_west, drop_one=True, zero="nula", point="celá", minus="mínus"),
}


def _slavic_cardinal(n: int, lang: str) -> str:
    T = _SLAVIC[lang]
    if n < 0:
        return f"{T['minus']} {_slavic_cardinal(-n, lang)}"
    if n == 0:
        return T["zero"]
    chunks: List[int] = []
    while n:
        n, c = divmod(n, 1000)
        chunks.append(c)
    if len(chunks) - 1 > len(T["scales"]):
        raise OverflowError("number too large to verbalise")
    words: List[str] = []
    for i in range(len(chunks) - 1, -1, -1):
        x = chunks[i]
        if x == 0:
            continue
        h, r = divmod(x, 100)
        t, u = divmod(r, 10)
        if h:
            words.append(T["hundreds"][h])
        if t == 1:
            words.append(T["teens"][u])
        else:
            if t:
                words.append(T["tens"][t])
            if u and not (T["drop_one"] and i > 0 and x == 1):           # 'tysiąc', 'tisíc' — but 'одна тысяча'
                words.append(T["fem"].get(u, T["ones"][u]) if i == 1 else T["ones"][u])
        if i > 0:
            words.append(T["scales"][i - 1][T["plural"](x)])
    return " ".join(words)


def cardinal_ru(n: int) -> str:
    """21000 -> 'двадцать одна тысяча', 1000 -> 'одна тысяча'."""
    return _slavic_cardinal(n, "ru")


_RU_ORD_WORD = {"один": "первый", "одна": "первый", "два": "второй", "две": "второй", "три": "третий", "четыре": "четвертый", "пять": "пятый",
                "шесть": "шестой", "семь": "седьмой", "восемь": "восьмой", "девять": "девятый", "десять": "десятый",
                "одиннадцать": "одиннадцатый", "двенадцать": "двенадцатый", "тринадцать": "тринадцатый", "четырнадцать": "четырнадцатый",
                "пятнадцать": "пятнадцатый", "шестнадцать": "шестнадцатый", "семнадцать": "семнадцатый", "восемнадцать": "восемнадцатый",
                "девятнадцать": "девятнадцатый", "двадцать": "двадцатый", "тридцать": "тридцатый", "сорок": "сороковой",
                "пятьдесят": "пятидесятый", "шестьдесят": "шестидесятый", "семьдесят": "семидесятый", "восемьдесят": "восьмидесятый",
                "девяносто": "девяностый", "сто": "сотый", "двести": "двухсотый", "триста": "трехсотый", "четыреста": "четырехсотый",
                "пятьсот": "пятисотый", "шестьсот": "шестисотый", "семьсот": "семисотый", "восемьсот": "восьмисотый",
                "девятьсот": "девятисотый", "ноль": "нулевой"}
_RU_THOUSANDS_PREFIX = {1: "", 2: "двух", 3: "трех", 4: "четырех", 5: "пяти", 6: "шести", 7: "семи", 8: "восьми", 9: "девяти", 10: "десяти"}


def ordinal_ru(n: int) -> str:
    """Masculine nominative: 21 -> 'двадцать первый', 100 -> 'сотый', 2000 -> 'двухтысячный'.  Only the last word of the cardinal
    changes; round thousands above ten thousand and round millions keep their digits (the caller then reads them as cardinals)."""
    if n % 1000 == 0 and n > 0:
        k = n // 1000
        return _RU_THOUSANDS_PREFIX[k] + "тысячный" if k in _RU_THOUSANDS_PREFIX else str(n)
    words = cardinal_ru(n).split(" ")
    words[-1] = _RU_ORD_WORD.get(words[-1], words[-1])
    return " ".join(words)


def cardinal_pl(n: int) -> str:
    """1000 -> 'tysiąc', 2000 -> 'dwa tysiące', 5000 -> 'pięć tysięcy'."""
    return _slavic_cardinal(n, "pl")


def cardinal_cs(n: int) -> str:
    """21 -> 'dvacet jedna', 1000 -> 'tisíc', 2000 -> 'dva tisíce'."""
    return _slavic_cardinal(n, "cs")


# --------------------------------------------------------------------------------------------------- tables for textnorm.verbalise
_LIMIT = 10 ** 15           # the scale tables end at 10^12 (x 999): longer digit strings (ids, phone numbers) are read digit by digit


def _bounded(card: Callable[[int], str]) -> Callable[[int], str]:
    def f(n: int) -> str:
        if abs(n) >= _LIMIT:
            return " ".join(card(int(d)) for d in str(abs(n)))
        return card(n)
    f.__doc__ = card.__doc__
    return f


CARDINAL: Dict[str, Callable[[int], str]] = {k: _bounded(c) for k, c in (("nl", cardinal_nl), ("tr", cardinal_tr), ("hu", cardinal_hu),
                                                                          ("ru", cardinal_ru), ("pl", cardinal_pl), ("cs", cardinal_cs))}
ORDINAL: Dict[str, Callable[[int], str]] = {k: (lambda n, o=o, k=k: o(n) if abs(n) < _LIMIT else CARDINAL[k](n))
                                            for k, o in (("nl", ordinal_nl), ("tr", ordinal_tr), ("hu", ordinal_hu), ("ru", ordinal_ru))}


# (unit forms, sub-unit forms): two forms (one / many) for nl, hu, tr; three (1 / 2-4 / 5+, picked by the language's plural rule)
# for ru, pl, cs.  Read as "<amount> <unit>, <cents> <sub-unit>": ", " is the separator the reference trims integer amounts at
# (tokenizer.py:651-673, `and_equivalents` of these six languages)
_CURRENCY: Dict[str, Dict[str, Tuple[Tuple[str, ...], Tuple[str, ...]]]] = {
    "nl": {"EUR": (("euro", "euro"), ("cent", "cent")), "USD": (("dollar", "dollar"), ("cent", "cent")),
           "GBP": (("pond", "pond"), ("penny", "pence"))},
    "hu": {"EUR": (("euró", "euró"), ("cent", "cent")), "USD": (("dollár", "dollár"), ("cent", "cent")),
           "GBP": (("font", "font"), ("penny", "penny"))},
    "tr": {"EUR": (("euro", "euro"), ("sent", "sent")), "USD": (("dolar", "dolar"), ("sent", "sent")),
           "GBP": (("sterlin", "sterlin"), ("peni", "peni"))},
    "ru": {"EUR": (("евро", "евро", "евро"), ("цент", "цента", "центов")), "USD": (("доллар", "доллара", "долларов"), ("цент", "цента", "центов")),
           "GBP": (("фунт", "фунта", "фунтов"), ("пенс", "пенса", "пенсов"))},
    "pl": {"EUR": (("euro", "euro", "euro"), ("cent", "centy", "centów")), "USD": (("dolar", "dolary", "dolarów"), ("cent", "centy", "centów")),
           "GBP": (("funt", "funty", "funtów"), ("pens", "pensy", "pensów"))},
    "cs": {"EUR": (("euro", "euro", "euro"), ("cent", "centy", "centů")), "USD": (("dolar", "dolary", "dolarů"), ("cent", "centy", "centů")),
           "GBP": (("libra", "libry", "liber"), ("pence", "pence", "pencí"))},
}


def currency_words(amount: float, code: str, lang: str) -> str:
    """5.5 EUR -> 'vijf euro, vijftig cent' / 'пять евро, пятьдесят центов' — always both parts; the caller trims whole amounts."""
    unit, sub = _CURRENCY[lang][code]
    whole = int(amount)
    cents = int(round((amount - whole) * 100))
    card = CARDINAL[lang]

    def form(n: int, forms: Tuple[str, ...]) -> str:
        if len(forms) == 3:
            return forms[_SLAVIC[lang]["plural"](n)]
        return forms[0] if n == 1 else forms[1]
    return f"{card(whole)} {form(whole, unit)}, {card(cents)} {form(cents, sub)}"


def _split_float(value: float) -> Tuple[str, str]:
    whole, _, frac = repr(float(value)).partition(".")
    return whole, frac


def decimal_words(value: float, lang: str) -> str:
    whole, frac = _split_float(value)
    if "e" in whole or "e" in frac:
        return CARDINAL[lang](int(value))
    card = CARDINAL[lang]
    if lang in _SLAVIC:                       # the fraction as a number, its leading zeros spoken
        zeros = len(frac) - len(frac.lstrip("0"))
        tail = " ".join([_SLAVIC[lang]["zero"]] * zeros + ([card(int(frac))] if frac.strip("0") else []))
        return f"{card(int(whole))} {_SLAVIC[lang]['point']} {tail or _SLAVIC[lang]['zero']}"
    if lang == "hu":
        unit = _HU_FRACTION.get(len(frac))
        if unit is None:
            return f"{card(int(whole))} egész " + " ".join(card(int(d)) for d in frac)
        return f"{card(int(whole))} egész {card(int(frac))} {unit}"
    point = {"nl": "komma", "tr": "virgül"}[lang]
    return f"{card(int(whole))} {point} " + " ".join(card(int(d)) for d in frac)
