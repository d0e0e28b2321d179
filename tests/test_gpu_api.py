"""GPU: the drop-in surface — TTS().from_pretrained / TTSRequest / generate_speech — end to end on the small
geometry, checked against the oracle chunk by chunk (greedy: token ids bit-exact, waveform fp32 tolerance)."""
import numpy as np
import pytest
import torch

from auralis_b200 import TTS, TTSRequest
from auralis_b200.weights import save_model_dir
from oracle import xtts_oracle as O

pytestmark = pytest.mark.gpu


def _wav_bytes(seconds, f0, seed):
    import io, wave
    x = O.synthetic_reference_wav(seconds, 22050, f0, seed).numpy()
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(22050)
        w.writeframes((np.clip(x, -1, 1) * 32767).astype(np.int16).tobytes())
    return buf.getvalue()


@pytest.fixture(scope="module")
def tts(tmp_path_factory, dims_small, state_small):
    d = tmp_path_factory.mktemp("model")
    save_model_dir(str(d), dims_small, state_small[0], state_small[1])
    t = TTS(scheduler_max_concurrency=16).from_pretrained(str(d), precision="fp32", max_concurrency=8)
    yield t
    t.loop.run_until_complete(t.shutdown())


TEXT = ("The quick brown fox jumps over the lazy dog near the quiet river bank. " * 6).strip()


def test_generate_speech_matches_oracle(tts, dims_small, state_small):
    spk = _wav_bytes(2.5, 150.0, 4)
    req = TTSRequest(text=TEXT, speaker_files=spk, language="en", temperature=0.0, repetition_penalty=5.0)
    out = tts.generate_speech(req)
    eng = tts.tts_engine
    chunks = eng.prepare_text_tokens(TEXT, "en")
    assert len(chunks) >= 2
    cond, g = tts.loop.run_until_complete(eng.get_audio_conditioning(spk, req.max_ref_length, req.gpt_cond_len, req.gpt_cond_chunk_len))
    orc = O.GPTOracle(state_small[0], state_small[1], dims_small)
    sp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=dims_small.gpt.max_audio_tokens,
                          stop_token=dims_small.gpt.stop_audio_token)
    waves = []
    for ids in chunks:
        toks, lats = orc.generate(torch.from_numpy(np.asarray(cond[0])), ids, sp)
        waves.append(O.vocoder(lats, torch.from_numpy(np.asarray(g).reshape(-1)), state_small[1], dims_small).numpy())
    ref = np.concatenate(waves)
    assert out.array.dtype == np.float32 and out.sample_rate == 24000
    assert out.array.shape == ref.shape
    assert np.abs(out.array - ref).max() < 5e-4, np.abs(out.array - ref).max()


def test_streaming_and_async(tts):
    spk = _wav_bytes(2.5, 150.0, 4)
    req = TTSRequest(text=TEXT, speaker_files=spk, language="en", stream=True, seed=3)
    chunks = list(tts.generate_speech(req))
    assert len(chunks) >= 2 and all(c.token_length > 0 and c.array.size > 0 for c in chunks)
    full = tts.generate_speech(TTSRequest(text=TEXT, speaker_files=spk, language="en", seed=3, request_id=req.request_id))
    np.testing.assert_array_equal(np.concatenate([c.array for c in chunks]), full.array)     # seeded => reproducible
    outs = tts.generate_speech_batch([TTSRequest(text=TEXT[: 80 * (i + 1)], speaker_files=spk, language="en") for i in range(5)])
    assert len(outs) == 5 and all(o.array.size > 0 for o in outs)


def test_prepared_speaker_partial(tts):
    spk = _wav_bytes(1.5, 110.0, 8)
    req = TTSRequest(text="Short sentence for a cached voice.", speaker_files=spk, language="en", temperature=0.0)
    fn = tts.loop.run_until_complete(tts.prepare_for_streaming_generation(req))
    a = tts.generate_speech(TTSRequest(text=req.text, speaker_files=spk, language="en", temperature=0.0, context_partial_function=fn))
    b = tts.generate_speech(req)
    np.testing.assert_array_equal(a.array, b.array)
