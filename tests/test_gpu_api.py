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


CFG5_TEXTS = {
    "en": "The harbour lights were already burning when the last ferry came in. Nobody spoke on the quay.",
    "fr": "Les lumières du port brûlaient déjà quand le dernier bac est arrivé. Personne ne parlait sur le quai.",
    "es": "Las luces del puerto ya ardían cuando llegó el último transbordador. Nadie hablaba en el muelle.",
    "de": "Die Hafenlichter brannten schon, als die letzte Fähre einlief. Niemand sprach auf dem Kai.",
    "it": "Le luci del porto erano già accese quando arrivò l'ultimo traghetto. Nessuno parlava sul molo.",
    "pt": "As luzes do porto já ardiam quando chegou a última balsa. Ninguém falava no cais.",
    "ja": "最後のフェリーが入ってきたとき、港の明かりはもう灯っていた。桟橋では誰も話さなかった。",
    "zh-cn": "最后一班渡轮进港的时候，港口的灯已经亮了。码头上没有人说话。",
}


def test_cfg5_multilingual_speaker_mix_matches_oracle(tts, dims_small, state_small):
    """BASELINE configs[4] shape at test geometry: one batch mixing 8 languages and 8 speakers, greedy decode;
    every request's waveform must match the oracle run chunk by chunk with that request's own speaker and language
    (per-language split limits, language tag in the token stream), i.e. no cross-talk inside the batch."""
    spks = [_wav_bytes(1.5 + 0.2 * i, 90.0 + 20.0 * i, 20 + i) for i in range(8)]
    reqs = [TTSRequest(text=" ".join([t] * (3 if i % 2 == 0 else 1)), speaker_files=spks[i], language=lang, temperature=0.0,
                       repetition_penalty=5.0) for i, (lang, t) in enumerate(CFG5_TEXTS.items())]     # 1..3 chunks each
    outs = tts.generate_speech_batch(reqs)
    eng = tts.tts_engine
    orc = O.GPTOracle(state_small[0], state_small[1], dims_small)
    sp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=dims_small.gpt.max_audio_tokens,
                          stop_token=dims_small.gpt.stop_audio_token)
    worst = 0.0
    for req, out in zip(reqs, outs):
        cond, g = tts.loop.run_until_complete(eng.get_audio_conditioning(req.speaker_files, req.max_ref_length, req.gpt_cond_len,
                                                                         req.gpt_cond_chunk_len))
        waves = []
        for ids in eng.prepare_text_tokens(req.text, req.language):
            toks, lats = orc.generate(torch.from_numpy(np.asarray(cond[0])), ids, sp)
            waves.append(O.vocoder(lats, torch.from_numpy(np.asarray(g).reshape(-1)), state_small[1], dims_small).numpy())
        ref = np.concatenate(waves)
        assert out.array.shape == ref.shape, req.language
        mse = float(np.mean((out.array - ref) ** 2))
        worst = max(worst, float(np.abs(out.array - ref).max()))
        assert mse < 1e-8, (req.language, mse)
    assert worst < 5e-4, worst
