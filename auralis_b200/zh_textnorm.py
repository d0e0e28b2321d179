"""Chinese non-standard-word normalisation in front of the BPE: digits, dates, money, telephone numbers, fractions and
percentages are rewritten as Chinese numerals, then every punctuation mark becomes an ASCII comma.

Restates what `expand_numbers_multilingual(text, "zh")` does in the reference — `zh_num2words.TextNorm()` with its
default flags, i.e. `normalize_nsw` followed by the punctuation translation
(`/root/reference/src/auralis/models/xttsv2/components/tts/layers/xtts/zh_num2words.py:649-750` numeral writer,
`:928-1016` the nine rewriting passes, `:63-68` punctuation, `:1083-1116` driver; called from
`config/tokenizer.py:681-683`).  That module is the reference's own code and imports only the standard library, so this
restatement is pinned against it directly: golden records in `tests/golden/zh_textnorm.json` and a live fuzz when
/root/reference is mounted (`tests/test_zh_textnorm.py`).

The passes are order-dependent and each rewrites the FIRST occurrence of the matched substring in the current text
(`str.replace(..., 1)`), exactly as the reference does — a match found late in the string can therefore rewrite an
identical substring earlier in it; that behaviour is kept.
"""
from __future__ import annotations

import re
import string
from typing import List, Optional, Tuple

_DIGITS = "零一二三四五六七八九"
_LIANG = "两"
_POINT = "点"
# (power of ten, character): 十 百 千 万, then 亿 = 10^8 and every further unit 10^4 apart ("mid" numbering)
_UNITS: List[Tuple[int, str]] = [(1, "十"), (2, "百"), (3, "千"), (4, "万")] + [((i + 2) * 4, c) for i, c in enumerate("亿兆京垓秭穰沟涧正载")]

# measure words / currency units the passes key on — data of the reference (zh_num2words.py:51-60), needed verbatim for parity
_CURRENCY_UNITS = "((亿|千万|百万|万|千|百)|(亿|千万|百万|万|千|百|)元|(亿|千万|百万|万|千|百|)块|角|毛|分)"
_QUANTIFIERS = ("(匹|张|座|回|场|尾|条|个|首|阙|阵|网|炮|顶|丘|棵|只|支|袭|辆|挑|担|颗|壳|窠|曲|墙|群|腔|砣|座|客|贯|扎|捆|刀|令|打|手|罗|坡|山|岭|江|溪|钟|队|"
                "单|双|对|出|口|头|脚|板|跳|枝|件|贴|针|线|管|名|位|身|堂|课|本|页|家|户|层|丝|毫|厘|分|钱|两|斤|担|铢|石|钧|锱|忽|(千|毫|微)克|毫|厘|分|寸|尺|"
                "丈|里|寻|常|铺|程|(千|分|厘|毫|微)米|撮|勺|合|升|斗|石|盘|碗|碟|叠|桶|笼|盆|盒|杯|钟|斛|锅|簋|篮|盘|桶|罐|瓶|壶|卮|盏|箩|箱|煲|啖|袋|钵|年|月|"
                "日|季|刻|时|周|天|秒|分|旬|纪|岁|世|更|夜|春|夏|秋|冬|代|伏|辈|丸|泡|粒|颗|幢|堆|条|根|支|道|面|片|张|颗|块)")
_CN_PUNCS = ("！？｡。" "＂＃＄％＆＇（）＊＋，－／：；＜＝＞＠［＼］＾＿｀｛｜｝～｟｠｢｣､、〃《》「」『』【】〔〕〖〗〘〙〚〛〜〝〞〟〰〾〿–—‘’‛“”„‟…‧﹏·〈〉-")
_PUNCS = _CN_PUNCS + string.punctuation
_PUNCS_TO_COMMA = str.maketrans(_PUNCS, "," * len(_PUNCS))

_RE_DATE = re.compile(r"\D+((([089]\d|(19|20)\d{2})年)?(\d{1,2}月(\d{1,2}[日号])?)?)")
_RE_MONEY = re.compile(r"\D+((\d+(\.\d+)?)[多余几]?" + _CURRENCY_UNITS + r"(\d" + _CURRENCY_UNITS + r"?)?)")
_RE_MOBILE = re.compile(r"\D((\+?86 ?)?1([38]\d|5[0-35-9]|7[678]|9[89])\d{8})\D")
_RE_FIXED = re.compile(r"\D((0(10|2[1-3]|[3-9]\d{2})-?)?[1-9]\d{6,7})\D")
_RE_FRACTION = re.compile(r"(\d+/\d+)")
_RE_PERCENT = re.compile(r"(\d+(\.\d+)?%)")
_RE_QUANTITY = re.compile(r"(\d+(\.\d+)?)[多余几]?" + _QUANTIFIERS)
_RE_DIGITS = re.compile(r"(\d{4,32})")
_RE_NUMBER = re.compile(r"(\d+(\.\d+)?)")
_RE_X2Y = re.compile(r"(([a-zA-Z]+)二([a-zA-Z]+))")


# symbols of a numeral under construction: ("d", value) or ("u", power, char)
def _group(value: str, zeros: bool = True) -> list:
    """zh_num2words.py:660-674: the recursive split at the largest unit below the (zero-stripped) length; the HIGH part keeps
    its leading zeros, the LOW part is taken from the stripped string."""
    stripped = value.lstrip("0")
    if not stripped:
        return []
    if len(stripped) == 1:
        d = ("d", int(stripped))
        return [("d", 0), d] if (zeros and len(value) != len(stripped)) else [d]
    power, char = next(u for u in reversed(_UNITS) if u[0] < len(stripped))
    return _group(value[:-power]) + [("u", power, char)] + _group(stripped[-power:])


def num2chn(number: str, alt_two: bool = True, use_units: bool = True) -> str:
    """'10500' -> '一万零五百', '202' -> '两百零二', '0.5' -> '零点五', '12' -> '十二'; use_units=False reads digit by digit."""
    parts = number.split(".")
    if len(parts) > 2:
        raise ValueError(f"invalid input num string with more than one dot: {number}")
    int_s, dec_s = parts[0], (parts[1] if len(parts) == 2 else "")
    syms = _group(int_s) if (use_units and len(int_s) > 1) else [("d", int(c)) for c in int_s]
    if dec_s:
        syms = syms + [("p",)] + [("d", int(c)) for c in dec_s]
    out = []
    for i, s in enumerate(syms):
        if s[0] == "d":
            ch = _DIGITS[s[1]]
            if alt_two and s[1] == 2:                      # 两 in front of 百 / 千 / 万 ..., never in front of or after 十
                nxt = syms[i + 1] if i + 1 < len(syms) else None
                prv = syms[i - 1] if i > 0 else None
                if nxt is not None and nxt[0] == "u" and (prv is None or prv[0] == "u"):
                    if nxt[1] != 1 and (prv is None or prv[1] != 1):
                        ch = _LIANG
            out.append(ch)
        elif s[0] == "u":
            out.append(s[2])
        else:
            out.append(_POINT)
    res = "".join(out)
    if res.startswith(_POINT):
        return _DIGITS[0] + res
    if len(res) >= 2 and res[1] == "十" and res[0] == "一":   # 一十二 -> 十二
        res = res[1:]
    return res


def _digits(s: str) -> str:
    return num2chn(s, alt_two=False, use_units=False)


def _date(date: str) -> str:
    """zh_num2words.py:864-886: the year digit by digit, month and day as cardinals."""
    if "年" in date:
        year, other = date.strip().split("年", 1)
        year = _digits(year) + "年"
    else:
        other, year = date, ""
    month = day = ""
    if other:
        if "月" in other:
            month, day = other.strip().split("月", 1)
            month = num2chn(month) + "月"
        else:
            day = date
        if day:
            day = num2chn(day[:-1]) + day[-1]
    return year + month + day


def _money(money: str) -> str:
    for m in _RE_NUMBER.findall(money):
        money = money.replace(m[0], num2chn(m[0]))
    return money


def _telephone(tel: str, fixed: bool) -> str:
    parts = tel.split("-") if fixed else tel.strip("+").split()
    return "".join(_digits(p) for p in parts)


def normalize_nsw(raw: str) -> str:
    """zh_num2words.py:928-1016."""
    text = "^" + raw + "$"
    for m in _RE_DATE.findall(text):
        text = text.replace(m[0], _date(m[0]), 1)
    for m in _RE_MONEY.findall(text):
        text = text.replace(m[0], _money(m[0]), 1)
    for m in _RE_MOBILE.findall(text):
        text = text.replace(m[0], _telephone(m[0], fixed=False), 1)
    for m in _RE_FIXED.findall(text):
        text = text.replace(m[0], _telephone(m[0], fixed=True), 1)
    for m in _RE_FRACTION.findall(text):
        num, den = m.split("/")
        text = text.replace(m, num2chn(den) + "分之" + num2chn(num), 1)
    text = text.replace("％", "%")
    for m in _RE_PERCENT.findall(text):
        text = text.replace(m[0], "百分之" + num2chn(m[0].strip().strip("%")), 1)
    for m in _RE_QUANTITY.findall(text):
        text = text.replace(m[0], num2chn(m[0]), 1)
    for m in _RE_DIGITS.findall(text):
        text = text.replace(m, _digits(m), 1)
    for m in _RE_NUMBER.findall(text):
        text = text.replace(m[0], num2chn(m[0]), 1)
    for m in _RE_X2Y.findall(text):                          # P2P, B2B ... got their 2 rewritten above: put it back
        text = text.replace(m[0], m[1] + "2" + m[2], 1)
    return text.lstrip("^").rstrip("$")


def normalize(text: str) -> str:
    """`TextNorm()(text)` with the default flags: rewrite the non-standard words, then punctuation -> ','."""
    return normalize_nsw(text).translate(_PUNCS_TO_COMMA)
