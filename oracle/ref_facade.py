"""TEST INFRASTRUCTURE — container-only loader for the reference's user-facing classes, imported unmodified:
`auralis.core.tts.TTS`, `TTSRequest`, `TTSOutput` (`/root/reference/src/auralis/core/tts.py`, `common/definitions/*.py`).
Their third-party imports that are absent from this image (sounddevice, IPython, torio, langid, librosa, soundfile,
pyloudnorm, colorama, vLLM 0.6.4) are replaced by empty stubs for the import of those modules only; none of the stubbed
functionality is reached by the differential tests (tests/test_facade_vs_reference.py), which drive both façades with the
same fake engine."""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

from . import ref_import


def _mod(name: str, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load():
    if not ref_import.available():
        raise RuntimeError("reference tree not mounted")
    ref_import.load()
    base = os.path.join(ref_import.REF_SRC, "auralis")
    for pkg in ("core", "common/definitions", "common/scheduling", "common/logging", "common/metrics"):
        ref_import._stub("auralis." + pkg.replace("/", "."), os.path.join(base, pkg))
    saved = {}

    def put(name, mod):
        saved[name] = sys.modules.get(name)
        sys.modules[name] = mod

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            raise RuntimeError("stubbed third-party call")

    put("sounddevice", _mod("sounddevice"))
    put("IPython", _mod("IPython"))
    put("IPython.display", _mod("IPython.display", Audio=_Anything, display=lambda *a, **k: None))
    put("torio", _mod("torio"))
    put("torio.io", _mod("torio.io", CodecConfig=_Anything))
    put("langid", _mod("langid", classify=lambda t: ("en", 1.0)))
    put("librosa", _mod("librosa"))
    put("soundfile", _mod("soundfile"))
    put("pyloudnorm", _mod("pyloudnorm"))
    put("vllm", _mod("vllm", RequestOutput=_Anything))
    if "auralis.common.logging.logger" not in sys.modules:
        lg = types.ModuleType("auralis.common.logging.logger")
        lg.setup_logger = lambda name=None, *a, **k: logging.getLogger("auralis.ref")
        lg.set_vllm_logging_level = lambda *a, **k: None
        sys.modules["auralis.common.logging.logger"] = lg
    else:
        sys.modules["auralis.common.logging.logger"].set_vllm_logging_level = lambda *a, **k: None
    try:
        tts = importlib.import_module("auralis.core.tts")
        req = importlib.import_module("auralis.common.definitions.requests")
        out = importlib.import_module("auralis.common.definitions.output")
        basem = importlib.import_module("auralis.models.base")
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return types.SimpleNamespace(TTS=tts.TTS, TTSRequest=req.TTSRequest, TTSOutput=out.TTSOutput,
                                 BaseAsyncTTSEngine=basem.BaseAsyncTTSEngine, tts_module=tts)
