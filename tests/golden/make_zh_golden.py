"""Generates tests/golden/zh_textnorm.json from the REFERENCE's own Chinese normaliser
(/root/reference/src/auralis/models/xttsv2/components/tts/layers/xtts/zh_num2words.py, imported unmodified; pure stdlib).
    python tests/golden/make_zh_golden.py          (container only)
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from zh_fixture import load_reference_zh, sentences  # noqa: E402


def main():
    z = load_reference_zh()
    norm = z.TextNorm()
    recs = [{"in": s, "out": norm(s)} for s in sentences(400, 20240923)]
    rng = random.Random(5)
    nums = ["0", "2", "10", "12", "20", "22", "100", "101", "110", "200", "202", "220", "1000", "1001", "1010", "1100", "2000", "2002", "2200",
            "10000", "10500", "20000", "22222", "100200", "1000000", "100000000", "120000000", "2000000000000", "0.5", "00.5", "3.14", "12.50"]
    nums += ["".join(rng.choice("0012359") for _ in range(rng.randint(1, 30))) for _ in range(150)]
    numrecs = [{"in": n, "cardinal": z.num2chn(n), "digits": z.num2chn(n, alt_two=False, use_units=False)} for n in nums]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zh_textnorm.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"source": "reference zh_num2words.TextNorm() / num2chn, imported unmodified", "sentences": recs, "numbers": numrecs},
                  f, ensure_ascii=False, indent=0)
    print(f"{len(recs)} sentences + {len(numrecs)} numbers -> {path}")


if __name__ == "__main__":
    main()
