"""TTSOutput — the boundary type of the hot path (array/sample_rate/token_length/start_time).

Mirrors `/root/reference/src/auralis/common/definitions/output.py:17-38,95-111` for the fields and
``combine_outputs``; the audio utilities (mp3/opus/aac encoders, phase-vocoder speed change, playback)
are CPU post-processing outside the hot path: wav/pcm/flac-free paths are provided
with the standard library, the rest raise with a clear message when their optional dependency is absent.
"""
from __future__ import annotations

import io
import wave
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Union

import numpy as np


@dataclass
class TTSOutput:
    array: Union[np.ndarray, bytes]
    sample_rate: int = 24000
    bit_depth: int = 32
    bit_rate: int = 192
    compression: int = 10
    channel: int = 1
    start_time: Optional[float] = None
    end_time: Optional[float] = None
    token_length: Optional[int] = None

    def __post_init__(self):
        if isinstance(self.array, bytes):          # output.py:30-38
            self.array = np.frombuffer(self.array, dtype=np.int16)
            self.array = self.array.astype(np.float32) / 32768.0
            fade_length = 100
            fade_in = np.linspace(0, 1, fade_length)
            self.array[:fade_length] *= fade_in

    @staticmethod
    def combine_outputs(outputs: List["TTSOutput"]) -> "TTSOutput":
        """output.py:95-111."""
        combined_audio = np.concatenate([out.array for out in outputs])
        return TTSOutput(array=combined_audio, sample_rate=outputs[0].sample_rate)

    def to_tensor(self):
        import torch
        if isinstance(self.array, np.ndarray):
            return torch.from_numpy(self.array)
        return self.array

    def to_bytes(self, format: str = "wav", sample_width: int = 2) -> bytes:
        """output.py:119-187; wav and pcm are native here, compressed codecs need torchaudio+ffmpeg."""
        wav = np.clip(np.asarray(self.array, dtype=np.float32), -1.0, 1.0)
        if format == "pcm":
            if sample_width == 2:
                return (wav * 32767).astype(np.int16).tobytes()
            if sample_width == 4:
                return (wav.astype(np.float64) * 2147483647).astype(np.int32).tobytes()
            return (wav * 127).astype(np.int8).tobytes()
        if format == "wav":
            buf = io.BytesIO()
            with wave.open(buf, "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(self.sample_rate)
                w.writeframes((wav * 32767).astype(np.int16).tobytes())
            return buf.getvalue()
        if format in ("flac", "mp3", "opus", "aac"):
            try:
                import torch
                import torchaudio
                buffer = io.BytesIO()
                torchaudio.save(buffer, torch.from_numpy(wav)[None], self.sample_rate,
                                format={"aac": "adts"}.get(format, format))
                return buffer.getvalue()
            except Exception as e:   # codec backends are optional
                raise RuntimeError(f"format {format!r} needs a torchaudio codec backend: {e}") from e
        raise ValueError(f"Unsupported format: {format}. Supported formats are: mp3, opus, aac, flac, wav, pcm")

    def save(self, filename: Union[str, Path], sample_rate: Optional[int] = None, format: Optional[str] = None) -> None:
        out = self if not sample_rate or sample_rate == self.sample_rate else self.resample(sample_rate)
        fmt = format or (Path(filename).suffix.lstrip(".") or "wav")
        with open(filename, "wb") as f:
            f.write(out.to_bytes(fmt))

    def resample(self, new_sample_rate: int) -> "TTSOutput":
        if new_sample_rate == self.sample_rate:
            return self
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(new_sample_rate), int(self.sample_rate))
        y = resample_poly(np.asarray(self.array, np.float32), new_sample_rate // g, self.sample_rate // g)
        return TTSOutput(array=y.astype(np.float32), sample_rate=new_sample_rate)

    def change_speed(self, speed_factor: float) -> "TTSOutput":
        if speed_factor <= 0:
            raise ValueError("Speed factor must be positive")
        if speed_factor == 1.0:
            return self
        try:
            import librosa
        except ImportError as e:
            raise RuntimeError("change_speed needs librosa (CPU post-processing, outside the hot path)") from e
        wav = np.asarray(self.array, np.float32)
        D = librosa.stft(wav, n_fft=2048, hop_length=512)
        y = librosa.istft(librosa.phase_vocoder(D, rate=speed_factor, hop_length=512), hop_length=512)
        return TTSOutput(array=librosa.util.normalize(y, norm=np.inf), sample_rate=self.sample_rate)

    @property
    def duration_s(self) -> float:
        return len(self.array) / float(self.sample_rate)
