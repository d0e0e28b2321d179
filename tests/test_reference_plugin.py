"""CPU, container only: INTEGRATION.md path B executed — `auralis_b200.reference_plugin` builds an engine class on the
REFERENCE's own `BaseAsyncTTSEngine` (ABC + torch.nn.Module, abstract `get_memory_usage_curve`), registers it in the
reference's registry and the reference's unmodified `TTS` drives it: one-shot, streaming, and a failing chunk.  The native
library is replaced by the fake native layer of test_host_logic (no GPU here); everything above the C ABI is the real code."""
import asyncio

import numpy as np
import pytest

from oracle import ref_import

pytestmark = [pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted"), pytest.mark.timeout(120)]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_facade
    return ref_facade.load()


def _plugged(ref, **kw):
    import importlib
    from auralis_b200 import reference_plugin
    from test_host_logic import _host_engine
    ref_base = importlib.import_module("auralis.models.base")
    cls = reference_plugin.make_reference_engine(ref_base, ref.TTSOutput)
    assert issubclass(cls, ref.BaseAsyncTTSEngine)
    inner = _host_engine(max_speakers=2, **kw)
    eng = cls(inner)                                   # TypeError here if an abstract method of the reference's ABC is missing
    tts = ref.TTS(scheduler_max_concurrency=4)
    tts._ensure_event_loop()
    tts.tts_engine = eng
    return tts, eng


def _wav(level):
    import io, wave
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(22050)
        w.writeframes((np.full(4000, level) * 32767).astype(np.int16).tobytes())
    return buf.getvalue()


TEXT = ("The quick brown fox jumps over the lazy dog near the quiet river bank. " * 5).strip()


def test_registers_in_the_reference_registry(ref):
    import importlib
    from auralis_b200 import reference_plugin
    cls = reference_plugin.register("xtts_b200_test")
    reg = importlib.import_module("auralis.models.registry").MODEL_REGISTRY
    assert reg["xtts_b200_test"] is cls and issubclass(cls, ref.BaseAsyncTTSEngine)


def test_reference_tts_one_shot_and_streaming(ref):
    tts, eng = _plugged(ref)
    try:
        import torch
        assert eng.dtype == torch.float32 and eng.device.type == "cuda" and eng.max_gb_for_vllm_model > 0
        req = ref.TTSRequest(text=TEXT, speaker_files=[_wav(0.2)], language="en")
        out = tts.generate_speech(req)
        assert isinstance(out, ref.TTSOutput) and out.array.dtype == np.float32
        n_chunks = len(eng.inner.prepare_text_tokens(TEXT, "en"))
        assert n_chunks >= 2 and out.array.shape[0] == 8 * n_chunks          # the fake native layer: 8 samples per chunk
        sreq = ref.TTSRequest(text=TEXT, speaker_files=[_wav(0.2)], language="en", stream=True)
        pieces = list(tts.generate_speech(sreq))
        assert len(pieces) == n_chunks and all(isinstance(p, ref.TTSOutput) for p in pieces)
        np.testing.assert_array_equal(np.concatenate([p.array for p in pieces]), out.array)
        assert eng.inner.native.cond_calls == 1                               # the speaker was conditioned once, then cached
    finally:
        tts.loop.run_until_complete(tts.shutdown())


def test_reference_tts_failing_chunk_reaches_the_caller(ref):
    tts, eng = _plugged(ref, fail_ids=(2,))
    try:
        req = ref.TTSRequest(text=TEXT, speaker_files=[_wav(0.3)], language="en")
        with pytest.raises(Exception) as ei:
            tts.generate_speech(req)
        assert "failed" in str(ei.value) or "chunk" in str(ei.value)
    finally:
        tts.loop.run_until_complete(tts.shutdown())
