"""TTSRequest — same fields and defaults as the reference
(`/root/reference/src/auralis/common/definitions/requests.py:135-204`).

Differences, all outside the hot path: language auto-detection uses ``langid`` when it is installed and
otherwise falls back to "en" with a warning (the reference hard-requires langid); ``enhance_speech``
(off by default, requests.py:171; CPU DSP on the reference wav) is not provided: when requested, the original speaker
files are used and a warning is issued — the reference's own behaviour when its enhancer fails.
"""
from __future__ import annotations

import uuid
import warnings
from dataclasses import dataclass, field
from functools import lru_cache
from typing import AsyncGenerator, Callable, List, Optional, Union

SUPPORTED_LANGUAGES = ("en", "es", "fr", "de", "it", "pt", "pl", "tr", "ru", "nl", "cs", "ar", "zh-cn", "hu", "ko",
                       "ja", "hi", "auto", "")


@lru_cache(maxsize=1024)
def get_language(text: str) -> str:
    """requests.py:97-113."""
    try:
        import langid
    except ImportError:
        warnings.warn("langid is not installed: language='auto' falls back to 'en'")
        return "en"
    detected = langid.classify(text)[0].strip()
    return "zh-cn" if detected == "zh" else detected


def validate_language(language: str) -> str:
    if language not in SUPPORTED_LANGUAGES:
        raise ValueError(f"Language {language} not supported. Must be one of {SUPPORTED_LANGUAGES}")
    return language


@dataclass
class AudioPreprocessingConfig:
    """Placeholder with the reference's field names (enhancer.py:12-31); only carried, never applied here."""
    sample_rate: int = 22050
    normalize: bool = True
    trim_silence: bool = True
    remove_noise: bool = True
    enhance_speech: bool = True


@dataclass
class TTSRequest:
    text: Union[AsyncGenerator[str, None], str, List[str]]
    speaker_files: Union[Union[str, List[str]], Union[bytes, List[bytes]]]
    context_partial_function: Optional[Callable] = None
    start_time: Optional[float] = None
    enhance_speech: bool = False
    audio_config: AudioPreprocessingConfig = field(default_factory=AudioPreprocessingConfig)
    language: str = "auto"
    request_id: str = field(default_factory=lambda: uuid.uuid4().hex)
    load_sample_rate: int = 22050
    sound_norm_refs: bool = False
    # voice conditioning (requests.py:179-181)
    max_ref_length: int = 60
    gpt_cond_len: int = 30
    gpt_cond_chunk_len: int = 4
    # generation (requests.py:183-190)
    stream: bool = False
    temperature: float = 0.75
    top_p: float = 0.85
    top_k: int = 50
    repetition_penalty: float = 5.0
    length_penalty: float = 1.0     # carried, not consumed by the engine
    do_sample: bool = True          # carried, not consumed by the engine
    # additions (not in the reference): reproducible sampling
    seed: Optional[int] = None

    def __post_init__(self):
        if self.language == "auto" and len(self.text) > 0:
            self.language = get_language(self.text if isinstance(self.text, str) else " ".join(self.text))
        validate_language(self.language)
        if self.enhance_speech and isinstance(self.speaker_files, list):
            # requests.py:199-203 + 214-247: the reference cleans the reference audio on the CPU (librosa / pyloudnorm) and
            # falls back to the original file when that fails.  That DSP is outside the hot path and not provided here:
            # same fallback, said once.
            warnings.warn("enhance_speech: reference-audio enhancement is not provided, the original speaker files are used",
                          RuntimeWarning, stacklevel=2)

    def infer_language(self):
        if self.language == "auto":
            self.language = get_language(self.text)

    def copy(self) -> "TTSRequest":
        """requests.py:250-277."""
        fields = {k: getattr(self, k) for k in (
            "text", "speaker_files", "enhance_speech", "audio_config", "language", "request_id", "load_sample_rate",
            "sound_norm_refs", "max_ref_length", "gpt_cond_len", "gpt_cond_chunk_len", "stream", "temperature", "top_p",
            "top_k", "repetition_penalty", "length_penalty", "do_sample", "seed")}
        return TTSRequest(**fields)
