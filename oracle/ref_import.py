"""TEST INFRASTRUCTURE — container-only helper that imports the *unmodified*
reference modules from /root/reference/src through empty stub parent packages
(their real ``__init__`` files import vLLM 0.6.4 / librosa / spacy, which are
absent).  Used only by ``tests/golden/make_golden.py`` and by the non-GPU tests
that pin the oracle restatement against the reference when /root/reference is
mounted.  Nothing on the GPU box can call this (no /root/reference there).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_SRC = os.environ.get("XTTS_REFERENCE_SRC", "/root/reference/src")     # override to test without the tree


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "auralis"))


def _stub(name: str, path: str) -> None:
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__path__ = [path]          # namespace-like: sub-modules resolve from disk
    m.__package__ = name
    sys.modules[name] = m


def load():
    """Returns a namespace with the reference classes used as ground truth."""
    if not available():
        raise RuntimeError("reference tree not mounted")
    base = os.path.join(REF_SRC, "auralis")
    _stub("auralis", base)
    _stub("auralis.common", os.path.join(base, "common"))
    _stub("auralis.models", os.path.join(base, "models"))
    _stub("auralis.models.xttsv2", os.path.join(base, "models", "xttsv2"))
    _stub("auralis.models.xttsv2.components", os.path.join(base, "models", "xttsv2", "components"))
    _stub("auralis.models.xttsv2.components.tts", os.path.join(base, "models", "xttsv2", "components", "tts"))
    _stub("auralis.models.xttsv2.components.tts.layers",
          os.path.join(base, "models", "xttsv2", "components", "tts", "layers"))
    _stub("auralis.models.xttsv2.components.tts.layers.xtts",
          os.path.join(base, "models", "xttsv2", "components", "tts", "layers", "xtts"))
    ns = types.SimpleNamespace()
    hd = importlib.import_module("auralis.models.xttsv2.components.tts.layers.xtts.hifigan_decoder")
    le = importlib.import_module("auralis.models.xttsv2.components.tts.layers.xtts.latent_encoder")
    pe = importlib.import_module("auralis.models.xttsv2.components.tts.layers.xtts.perceiver_encoder")
    ut = importlib.import_module("auralis.common.utilities")
    ns.HifiDecoder = hd.HifiDecoder
    ns.HifiganGenerator = hd.HifiganGenerator
    ns.ResNetSpeakerEncoder = hd.ResNetSpeakerEncoder
    ns.ConditioningEncoder = le.ConditioningEncoder
    ns.PerceiverResampler = pe.PerceiverResampler
    ns.wav_to_mel_cloning = ut.wav_to_mel_cloning
    return ns
