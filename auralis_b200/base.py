"""Engine plugin API — the drop-in boundary (`/root/reference/src/auralis/models/base.py:46-224`,
`models/registry.py:1-4`).  Same abstract methods and registration; no torch.nn.Module base because the
weights live inside the native library, not in Python."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import AsyncGenerator

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

    @classmethod
    def from_pretrained(cls, *args, **kwargs) -> "BaseAsyncTTSEngine":
        raise NotImplementedError

    async def shutdown(self):
        pass


MODEL_REGISTRY = {}


def register_model(name, model):
    MODEL_REGISTRY[name] = model
