"""TEST INFRASTRUCTURE: generator of Chinese test sentences that exercise every pass of the zh normaliser, and a loader
for the reference's own module (pure standard library)."""
import importlib
import random

from oracle import ref_import

FIXED = [
    "我有2个苹果和12.5%的股份", "2023年10月5日，价格是3.50元", "电话13812345678或010-12345678", "1/3的人", "第22章 2222 20000 202 10500 0.5 12 100200",
    "P2P和B2B, 2两 2百 2千2百", "他花了1.2亿元买了3万多块地", "98年5月1号出生，08年上大学", "比分是3:2，温度-5度，占比100％", "+86 13912345678",
    "02187654321和021-87654321", "房间号1208，门牌2号", "重量2.50千克，长度1200米", "10个人，11只狗，12张票，20年，200天，2000岁", "约30多个国家，5余人，10几岁",
    "1001夜 1010 1100 10010", "0.05和00.5以及5.", "价格：￥99.9、$15.5", "A1B2C3", "", "没有数字。只有标点！？", "3角5分，2毛钱，5块2",
]
FRAGS = ["年", "月", "日", "号", "元", "块", "角", "分", "个", "只", "米", "千克", "%", "％", "/", ".", "-", "+86 ", " ", "，", "。", "的", "是", "约", "多", "余", "几",
         "万", "亿", "百", "千", "P", "B", "x", "第", "章"]


def sentences(n: int, seed: int):
    rng = random.Random(seed)
    out = list(FIXED)
    for _ in range(n):
        parts = []
        for _ in range(rng.randint(1, 9)):
            r = rng.random()
            if r < 0.45:
                k = rng.choice([1, 1, 2, 2, 3, 4, 5, 8, 11, 13])
                parts.append("".join(rng.choice("0123456789") for _ in range(k)))
            elif r < 0.55:
                parts.append(rng.choice(["13812345678", "15900001111", "01012345678", "021-87654321", "2023年10月5日", "98年", "12月25号", "3.14", "0.5"]))
            else:
                parts.append(rng.choice(FRAGS))
        out.append("".join(parts))
    return out


def load_reference_zh():
    if not ref_import.available():
        raise RuntimeError("reference tree not mounted")
    ref_import.load()
    return importlib.import_module("auralis.models.xttsv2.components.tts.layers.xtts.zh_num2words")
