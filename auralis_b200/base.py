"""Engine plugin API — the drop-in boundary (`/root/reference/src/auralis/models/base.py:46-224`,
`models/registry.py:1-4`).  Same abstract methods and registration; no torch.nn.Module base because the
weights live inside the native library, not in Python."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from pathlib import Path
from typing import AsyncGenerator, Optional, Union

from .output import TTSOutput
from .requests import TTSRequest


@dataclass
class ConditioningConfig:
    """models/base.py:46-55."""
    speaker_embeddings: bool = False
    gpt_like_decoder_conditioning: bool = False


class BaseAsyncTTSEngine(ABC):
    @abstractmethod
    async def get_generation_context(self, request: TTSRequest):
        raise NotImplementedError

    @abstractmethod
    async def process_tokens_to_speech(self, generator, speaker_embeddings, multimodal_data=None,
                                       request: TTSRequest = None) -> AsyncGenerator[TTSOutput, None]:
        raise NotImplementedError

    @property
    def conditioning_config(self) -> ConditioningConfig:
        raise NotImplementedError

    @property
    def device(self):
        """models/base.py:131-138 — where the weights live (a torch.device; the weights themselves are inside the library)."""
        raise NotImplementedError

    @property
    def dtype(self):
        """models/base.py:140-146."""
        raise NotImplementedError

    @abstractmethod
    def get_memory_usage_curve(self):
        """models/base.py:148-159 (abstract there too): device memory as a function of the concurrency level."""
        raise NotImplementedError

    @staticmethod
    def get_memory_percentage(memory: int) -> Optional[float]:
        """models/base.py:161-180: fraction of a GPU's memory `memory` more bytes would bring it to (None: no GPU fits)."""
        import torch
        for i in range(torch.cuda.device_count()):
            free_memory, total_memory = torch.cuda.mem_get_info(i)
            used_memory = total_memory - free_memory
            estimated = (memory + used_memory) / total_memory
            if 0 < estimated < 1:
                return estimated
        return None

    @classmethod
    def from_pretrained(cls, *args, **kwargs) -> "BaseAsyncTTSEngine":
        raise NotImplementedError

    @staticmethod
    def load_audio(audio_path: Union[str, Path], sampling_rate: int = 22050):
        """models/base.py:200-224: mono, resampled, clipped — a torch tensor [1, samples] like the reference's."""
        import torch
        from .engine import load_audio
        return torch.from_numpy(load_audio(audio_path, sampling_rate))[None]

    async def shutdown(self):
        pass


MODEL_REGISTRY = {}


def register_model(name, model):
    MODEL_REGISTRY[name] = model
