"""Text front-end of the hot path: sentence splitting into <= char-limit chunks and token ids.

Restates the splitting rules of `/root/reference/src/auralis/models/xttsv2/config/tokenizer.py`
(`find_best_split_point` :51-116, `split_sentence` :119-236, `char_limits` :784-789,
`batch_encode_with_split` :821-869) without spaCy: sentence boundaries are found with a punctuation
sentencizer equivalent to spaCy's rule-based ``sentencizer`` for the scripts XTTS supports.

Token ids: with a real ``tokenizer.json`` (HF `tokenizers`) in the GPT model directory the BPE of the
reference is used (`XTTSTokenizerFast`, tokenizer.py:742-942) behind the per-language cleaners of `textnorm.py`
(abbreviation / symbol / number passes of all 15 cleaned languages pinned against the reference's functions; number
words restated for en/es/fr/de/it/pt, digits kept elsewhere — SURVEY §8f-1).  Without one (this build has no network)
ids are synthetic: ceil(chars/3.2) ids drawn uniformly from the text vocabulary, seeded by the chunk's
hash — deterministic, same chunk -> same ids.
"""
from __future__ import annotations

import hashlib
import math
import os
import re
from typing import List, Optional

import numpy as np

from .textnorm import basic_cleaners, format_for_bpe, preprocess_text  # noqa: F401  (re-exported)

CHAR_LIMITS = {"en": 250, "de": 253, "fr": 273, "es": 239, "it": 213, "pt": 203, "pl": 224, "zh": 82, "ar": 166,
               "cs": 186, "ru": 182, "nl": 251, "tr": 226, "ja": 71, "hu": 224, "ko": 95}

_MARKERS = [
    (r"[.!?؟။။။]+[\s]*", 1.0), (r"[\n\r]+\s*[\n\r]+", 1.0), (r"[:|;；：；][\s]*", 0.9), (r"[,，،、][\s]*", 0.8),
    (r"[)}\]）】』»›》\s]+", 0.7), (r"[-—−]+[\s]*", 0.7), (r"\s+[&+=/\s]+\s+", 0.6), (r"[\s]+", 0.5),
]


def find_best_split_point(text: str, target_pos: int, window_size: int = 30) -> int:
    """tokenizer.py:51-116."""
    start = max(0, target_pos - window_size)
    end = min(len(text), target_pos + window_size)
    window = text[start:end]
    best_pos, best_score = target_pos, 0.0
    for pattern, priority in _MARKERS:
        for match in re.finditer(pattern, window):
            pos = start + match.end()
            score = priority * (1 - (abs(pos - target_pos) / (window_size * 2)))
            if score > best_score:
                best_score, best_pos = score, pos
    return best_pos


from .sentencizer import sentencize as _sentencize      # spaCy's sentencizer + blank-language tokenizer rules, restated


def sentencize(text: str, lang: str = "en") -> List[str]:
    """Sentences as spaCy's rule-based `sentencizer` yields them for the language object the reference picks
    (`config/tokenizer.py:25-48`); see sentencizer.py."""
    return _sentencize(text, lang)


def split_sentence(text: str, lang: str, text_split_length: int = 250) -> List[str]:
    """tokenizer.py:119-236 (same packing of sentences into chunks, same long-sentence handling,
    same trailing '.' -> ' ' replacement, App. B.12)."""
    text = text.strip()
    if len(text) <= text_split_length:
        return [text]
    splits, current, current_length = [], [], 0
    for sentence_text in sentencize(text, lang):
        n = len(sentence_text)
        if current_length + n <= text_split_length:
            current.append(sentence_text)
            current_length += n + 1
        elif n > text_split_length:
            if current:
                splits.append(" ".join(current))
                current, current_length = [], 0
            remaining = sentence_text
            while len(remaining) > text_split_length:
                pos = find_best_split_point(remaining, text_split_length, window_size=30)
                splits.append(remaining[:pos].strip())
                remaining = remaining[pos:].strip()
            if remaining:
                current, current_length = [remaining], len(remaining)
        else:
            splits.append(" ".join(current))
            current, current_length = [sentence_text], n
    if current:
        splits.append(" ".join(current))
    return [s[:-1] + " " if s.endswith(".") else s for s in splits if s]


_WS = re.compile(r"\s+")


class XTTSTokenizer:
    """Chunk -> ids.  ``bos``/``eos`` are added by the engine (XTTSv2.py:519-522)."""

    def __init__(self, n_text_tokens: int, max_text_tokens: int, tokenizer_file: Optional[str] = None):
        self.n_text_tokens = n_text_tokens
        self.max_text_tokens = max_text_tokens
        self.char_limits = dict(CHAR_LIMITS)
        self.tok = None
        self.bos_token_id, self.eos_token_id = 0, 1          # synthetic stand-ins for [START]/[STOP]
        if tokenizer_file and os.path.exists(tokenizer_file):
            from tokenizers import Tokenizer
            from tokenizers.pre_tokenizers import WhitespaceSplit
            self.tok = Tokenizer.from_file(tokenizer_file)
            self.tok.pre_tokenizer = WhitespaceSplit()                           # tokenizer.py:763
            b, e = self.tok.token_to_id("[START]"), self.tok.token_to_id("[STOP]")
            if b is not None and e is not None:
                self.bos_token_id, self.eos_token_id = b, e

    @property
    def synthetic(self) -> bool:
        return self.tok is None

    def encode_chunk(self, chunk: str, lang: str) -> List[int]:
        base = lang.split("-")[0]
        if self.tok is not None:
            txt = format_for_bpe(chunk, lang)                                    # tokenizer.py:871-942
            ids = self.tok.encode(txt, add_special_tokens=False).ids
        else:
            n = max(1, int(math.ceil(len(chunk) / 3.2)))
            h = hashlib.sha256(f"{base}|{chunk}".encode("utf-8")).digest()
            rng = np.random.RandomState(int.from_bytes(h[:4], "little"))
            ids = rng.randint(2, self.n_text_tokens, size=n).tolist()
        return ids[: self.max_text_tokens]

    def batch_encode_with_split(self, text: str, lang: str) -> List[List[int]]:
        """tokenizer.py:821-869 for a single string (App. B.11): split by the language's char limit, encode."""
        base = lang.split("-")[0]
        limit = self.char_limits.get(base, 250)
        return [self.encode_chunk(c, lang) for c in split_sentence(text, base, limit) if c.strip()]
