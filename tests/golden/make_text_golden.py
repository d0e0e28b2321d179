"""Generates tests/golden/text_cleaners.json from the REFERENCE's own text front-end
(/root/reference/src/auralis/models/xttsv2/config/tokenizer.py, imported unmodified by oracle/ref_text.py with
num2words replaced by a marker function and spaCy's sentencizer by the punctuation rule).

    python tests/golden/make_text_golden.py          (container only: needs /root/reference)

Every record is (function, lang, input, reference output).  tests/test_text_golden.py replays them against
auralis_b200/textnorm.py + text.py with the same marker function injected.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_text  # noqa: E402

LANGS = ["en", "es", "fr", "de", "it", "pt", "pl", "ar", "zh", "cs", "ru", "nl", "tr", "hu", "ko"]

# per-language sentences touching every abbreviation, every symbol and the ordinal suffixes of that language
SAMPLES = {
    "en": ['Mrs. Jones and Mr. Smith met Dr. Who at St. James Co. with Jr. Maj. Gen. Drs. Rev. Lt. Hon. Sgt. Capt. Esq. Ltd. Col. Ft. Knox.',
           'It costs $5.50 & £3, or 20€; 50% off #1 @ home at 30°. The 1st, 2nd, 3rd and 24th of 1,234,567 items: 3.14 and 0.5.',
           'Pay $1,200.75 or 1200$ "now", 7£ later, €9 tomorrow. 1000000 people; 12,5 is odd.'],
    "es": ['La Sra. y el Sr. con el Dr. y la Dra. en St. Co. Jr. Ltd. pagaron 5,50€ & 3$ @ casa.',
           'El 1º y la 2ª, el 3er piso, 4o 5a 6os 7as: 1.234.567 cosas al 50% #2 a 30° por £8 y 3,14.'],
    "fr": ['Mme. Dupont et Mr. Martin avec Dr. Co. St. Jr. Ltd. ont payé 12,50€ & 7$ @ Paris.',
           'Le 1er, la 1re, le 2e, la 3ème, 4º 5ª: 1.234 euros, 50% de #3 à 30° pour £9 et 2,5.'],
    "de": ['Fr. Müller und Dr. Schmidt bei St. Co. Jr. zahlten 9,99€ & 4$ @ Berlin.',
           'Der 1. Mai und der 3. Platz, 2nd 4th 5º 6ª: 1.234.567 Stück, 50% von #4 bei 30° für £10 und 3,14. Ende 5.'],
    "it": ['Sig. Rossi e Dr. Bianchi di St. Co. Jr. Ltd. pagano 7,20€ & 3$ @ Roma.',
           'Il 1º, il 2° posto, la 3ª, 4o 5a 6i 7e: 1.234 cose, 50% di #5 a 30° per £11 e 2,75.'],
    "pt": ['A Sra. e o Sr. com o Dr. e a Dra. em St. Co. Jr. Ltd. pagaram 3,30€ & 2$ @ Lisboa.',
           'O 1º e a 2ª, 3o 4a 5os 6as: 1.234.567 coisas, 50% de #6 a 30° por £12 e 1,5.'],
    "pl": ['P. Kowalska i M. Nowak z Dr. Sw. Jr. zapłacili 4,40€ & 6$ @ Kraków.',
           'Miejsce 1º 2ª 3st 4nd 5rd 6th: 1.234 rzeczy, 50% z #7 przy 30° za £13 i 2,5.'],
    "ar": ['دفع 5,50€ & 3$ @ المنزل 50% #8 30° £14 و 3,14 ثم 12ون 13ين 14ث 15ر 16ى و 1.234.'],
    "zh": ['价格是 5.50€ & 3$ @ 家 50% #9 30° £15 和 3.14 以及 1,234 个。'],
    "cs": ['Dr. Novák a Ing. Svoboda s P. Dvořákem zaplatili 8,80€ & 5$ @ Praha.',
           'Dne 1. května a 3. místo: 1.234 věcí, 50% z #10 při 30° za £16 a 2,5. Konec 7.'],
    "ru": ['г-жа Иванова и г-н Петров с д-р Сидоровым заплатили 6,60€ & 4$ @ Москва.',
           '1-й 2-я 3-е 4-ое 5-ье 6-го: 1,234,567 вещей, 50% от #11 при 30° за £17 и 2.5.'],
    "nl": ['Dhr. Jansen en Mevr. de Vries met Dr. Bakker en Jhr. Six betaalden 2,20€ & 8$ @ Amsterdam.',
           'De 1ste, 2de, 3e: 1.234 dingen, 50% van #12 bij 30° voor £18 en 2,5.'],
    "tr": ['B. Yılmaz ve Byk. Kaya ile Dr. Demir 1,10€ & 9$ @ İstanbul ÖDEDİ ÜÇ.',
           '1. 2inci 3nci 4uncu 5üncü: 1.234 şey, 50% #13 30° £19 ve 2,5 kaldı.'],
    "hu": ['Dr. Nagy és B. Kovács meg Nőv. Szabó fizettek 9,90€ & 1$ @ Budapest.',
           '1. 2adik 3edik 4odik 5ödik 6ödike 7ik: 1.234 dolog, 50% #14 30° £20 és 2,5.'],
    "ko": ['가격은 5,50€ & 3$ @ 집 50% #15 30° £21 그리고 3,14 와 1번째 2번 3차 4째 그리고 1.234.'],
}

WORDS = ("the quick brown fox jumps over a lazy dog while seven bright stars slowly cross this quiet northern sky and nobody "
         "really knows why any river would ever run backwards into those hills").split()


def long_text(rng, n_sent, long_every=0):
    out = []
    for i in range(n_sent):
        n = rng.randint(4, 22)
        s = " ".join(rng.choice(WORDS) for _ in range(n)).capitalize()
        if long_every and i % long_every == long_every - 1:                 # an over-long sentence with inner punctuation
            s += ", " + "; ".join(" ".join(rng.choice(WORDS) for _ in range(rng.randint(6, 14))) for _ in range(6))
            s += " - " + " ".join(rng.choice(WORDS) for _ in range(30))
        out.append(s + rng.choice([".", ".", "!", "?", "..."]))
    return " ".join(out)


def main():
    ref = ref_text.load()
    recs = []
    for lang in LANGS:
        for s in SAMPLES[lang]:
            recs.append({"fn": "expand_abbreviations_multilingual", "lang": lang, "in": s.lower(),
                         "out": ref.expand_abbreviations_multilingual(s.lower(), lang)})
            recs.append({"fn": "expand_symbols_multilingual", "lang": lang, "in": s.lower(),
                         "out": ref.expand_symbols_multilingual(s.lower(), lang)})
            # (zh numbers go through the reference's own zh_num2words.TextNorm, not num2words)
            recs.append({"fn": "expand_numbers_multilingual", "lang": lang, "in": s.lower(),
                         "out": ref.expand_numbers_multilingual(s.lower(), lang)})
            recs.append({"fn": "multilingual_cleaners", "lang": lang, "in": s,
                         "out": ref.multilingual_cleaners(s, lang)})
    recs.append({"fn": "basic_cleaners", "lang": "xx", "in": "MiXed   Case\n\tText", "out": ref.basic_cleaners("MiXed   Case\n\tText")})
    rng = random.Random(20240917)
    for i, (lang, limit) in enumerate([("en", 250), ("en", 120), ("de", 253), ("it", 213), ("pt", 203), ("ar", 166), ("ja", 71),
                                       ("zh", 82), ("es", 239), ("en", 60)]):
        txt = long_text(rng, rng.randint(6, 40), long_every=(0 if i % 3 == 0 else 5))
        recs.append({"fn": "split_sentence", "lang": lang, "limit": limit, "in": txt, "out": ref.split_sentence(txt, lang, limit)})
    for target in (10, 50, 120, 249):
        txt = long_text(rng, 12, long_every=3)
        recs.append({"fn": "find_best_split_point", "lang": "en", "limit": target, "in": txt,
                     "out": ref.find_best_split_point(txt, target, 30)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "text_cleaners.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"source": "reference tokenizer.py imported unmodified; num2words = oracle.ref_text.marker_num2words; "
                             "spaCy sentencizer = auralis_b200.text.sentencize", "records": recs}, f, ensure_ascii=False, indent=0)
    print(f"{len(recs)} records -> {path}")


if __name__ == "__main__":
    main()
