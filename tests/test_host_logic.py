"""CPU: host-side mirror of the reference interface — TTSRequest/TTSOutput, text splitting, scheduler ordering,
timeouts and error propagation, and the TTS façade driven by a fake engine (no GPU)."""
import asyncio
import time

import numpy as np
import pytest

from auralis_b200 import TTS, TTSOutput, TTSRequest
from auralis_b200.base import BaseAsyncTTSEngine, ConditioningConfig
from auralis_b200.scheduler import TwoPhaseScheduler
from auralis_b200.text import XTTSTokenizer, find_best_split_point, split_sentence


def test_request_defaults_match_reference():
    r = TTSRequest(text="hi", speaker_files=["a.wav"], language="en")
    assert (r.temperature, r.top_p, r.top_k, r.repetition_penalty) == (0.75, 0.85, 50, 5.0)     # requests.py:185-188
    assert (r.max_ref_length, r.gpt_cond_len, r.gpt_cond_chunk_len, r.load_sample_rate) == (60, 30, 4, 22050)
    assert r.stream is False and r.enhance_speech is False and len(r.request_id) == 32
    c = r.copy()
    assert c.text == r.text and c.request_id == r.request_id
    with pytest.raises(ValueError):
        TTSRequest(text="hi", speaker_files=["a.wav"], language="xx")


def test_output_combine_and_bytes():
    a = TTSOutput(array=np.ones(10, np.float32) * 0.5, token_length=3)
    b = TTSOutput(array=np.zeros(5, np.float32))
    c = TTSOutput.combine_outputs([a, b])
    assert c.array.shape == (15,) and c.sample_rate == 24000
    assert len(c.to_bytes("pcm")) == 30 and c.to_bytes("wav")[:4] == b"RIFF"
    pcm = (np.arange(200, dtype=np.int16) * 100).tobytes()
    d = TTSOutput(array=pcm)                                   # bytes input: int16 -> float, 100-sample fade-in
    assert d.array.dtype == np.float32 and d.array[0] == 0.0
    assert c.resample(12000).array.shape[0] in (7, 8)
    assert c.get_info() == (15, 24000, 15 / 24000)                 # output.py:248-256


def test_accumulator_equals_combine_outputs():
    rng = np.random.RandomState(0)
    for sizes in ([5], [3, 0, 7], [1000] * 5, [10] * 20, [7, 900, 2, 5000, 1, 1, 1, 1, 1, 64]):
        outs = [TTSOutput(array=rng.randn(n).astype(np.float32)) for n in sizes]
        acc = TTSOutput.Accumulator()
        for o in outs:
            acc.add(o)
        want = TTSOutput.combine_outputs(outs)
        got = acc.result()
        assert got.sample_rate == want.sample_rate and got.array.dtype == want.array.dtype
        np.testing.assert_array_equal(got.array, want.array)
    with pytest.raises(ValueError):
        TTSOutput.Accumulator().result()


def test_output_file_and_tensor_round_trip(tmp_path):
    import torch
    x = (np.sin(np.arange(480) * 0.05) * 0.5).astype(np.float32)
    o = TTSOutput(array=x)
    f = tmp_path / "a.wav"
    o.save(f)
    back = TTSOutput.from_file(f)
    assert back.sample_rate == 24000 and np.array_equal(back.array, x)           # bit_depth 32 (default) = IEEE-float WAV, lossless
    assert f.read_bytes()[20:22] == b"\x03\x00"                                 # format tag 3, what torchaudio writes for the reference
    o16 = TTSOutput(array=x, bit_depth=16)
    o16.save(tmp_path / "c.wav")
    back16 = TTSOutput.from_file(tmp_path / "c.wav")
    assert np.abs(back16.array - x).max() < 1e-4                                  # int16 quantisation
    (tmp_path / "d.wav").write_bytes(o.to_bytes("wav"))                           # to_bytes('wav') stays 16-bit PCM (output.py:141-150)
    assert np.abs(TTSOutput.from_file(tmp_path / "d.wav").array - x).max() < 1e-4
    o.save(tmp_path / "b.wav", sample_rate=12000)
    assert TTSOutput.from_file(tmp_path / "b.wav").get_info()[:2] == (240, 12000)
    t = TTSOutput.from_tensor(torch.from_numpy(x)[None], 16000)
    assert t.sample_rate == 16000 and np.array_equal(t.array, x) and torch.equal(t.to_tensor(), torch.from_numpy(x))


def test_split_sentence_rules():
    assert split_sentence("short text.", "en", 250) == ["short text."]
    t = ("This is a sentence of some length that keeps going. " * 12).strip()
    chunks = split_sentence(t, "en", 250)
    assert all(len(c) <= 250 for c in chunks) and len(chunks) >= 3
    assert all(not c.endswith(".") for c in chunks)           # trailing '.' -> ' ' (tokenizer.py:234)
    long_one = "word " * 120                                    # one 600-char "sentence": split at whitespace
    chunks = split_sentence(long_one, "en", 250)
    assert all(len(c) <= 250 for c in chunks) and "".join(c.replace(" ", "") for c in chunks) == "word" * 120
    assert find_best_split_point("aaa, bbb ccc", 5, 30) == 5   # the comma+space ends exactly at the target


def test_tokenizer_synthetic_is_deterministic_and_bounded():
    tok = XTTSTokenizer(6681, 402)
    a = tok.batch_encode_with_split("Hello there. " * 40, "en")
    b = tok.batch_encode_with_split("Hello there. " * 40, "en")
    assert a == b and len(a) >= 2
    assert all(0 <= i < 6681 for c in a for i in c) and all(len(c) <= 402 for c in a)
    assert tok.batch_encode_with_split("こんにちは。" * 30, "ja") != []      # ja limit = 71 chars


class FakeEngine(BaseAsyncTTSEngine):
    """Yields a waveform whose values encode (chunk index) after a delay that is LONGER for earlier chunks."""

    def __init__(self, fail_at=None, delay=0.02):
        self.fail_at, self.delay, self.started = fail_at, delay, []

    def get_memory_usage_curve(self):
        return None

    @property
    def conditioning_config(self):
        return ConditioningConfig(True, True)

    async def get_audio_conditioning(self, speaker_files, *a, **k):
        return np.zeros((1, 32, 8), np.float32), np.zeros((1, 4, 1), np.float32)

    async def get_generation_context(self, request, gpt_cond_latent=None, speaker_embeddings=None):
        n = max(1, len(request.text) // 10)

        async def gen(i):
            self.started.append(i)
            await asyncio.sleep(self.delay * (n - i))
            if self.fail_at == i:
                raise RuntimeError(f"chunk {i} exploded")
            yield i
        return [gen(i) for i in range(n)], [f"r_{i}" for i in range(n)], np.zeros((1, 4, 1)), [np.zeros(1)] * n

    async def process_tokens_to_speech(self, generator, speaker_embeddings, multimodal_data=None, request=None):
        async for i in generator:
            yield TTSOutput(array=np.full(4, float(i), np.float32), token_length=1, start_time=request.start_time)


def test_generate_speech_orders_chunks_and_concatenates():
    tts = TTS(scheduler_max_concurrency=8).from_engine(FakeEngine())
    out = tts.generate_speech(TTSRequest(text="x" * 55, speaker_files=["s.wav"], language="en"))
    np.testing.assert_array_equal(out.array, np.repeat(np.arange(5, dtype=np.float32), 4))


def test_streaming_yields_in_order():
    tts = TTS(scheduler_max_concurrency=8).from_engine(FakeEngine())
    req = TTSRequest(text="x" * 40, speaker_files=["s.wav"], language="en", stream=True)
    got = [int(c.array[0]) for c in tts.generate_speech(req)]
    assert got == [0, 1, 2, 3]


def test_async_and_batch_api():
    tts = TTS(scheduler_max_concurrency=4).from_engine(FakeEngine(delay=0.005))
    reqs = [TTSRequest(text="y" * (10 * (i + 1)), speaker_files=["s.wav"], language="en") for i in range(4)]
    outs = tts.generate_speech_batch(reqs)
    assert [o.array.shape[0] for o in outs] == [4, 8, 12, 16]

    async def stream():
        r = TTSRequest(text="z" * 30, speaker_files=["s.wav"], language="en", stream=True)
        return [int(c.array[0]) async for c in await tts.generate_speech_async(r)]
    assert tts.loop.run_until_complete(stream()) == [0, 1, 2]


def test_failed_chunk_fails_the_request():
    tts = TTS().from_engine(FakeEngine(fail_at=1))
    with pytest.raises(RuntimeError, match="chunk 1 exploded"):
        tts.generate_speech(TTSRequest(text="x" * 30, speaker_files=["s.wav"], language="en"))


def test_concurrency_limit_and_timeouts():
    eng = FakeEngine(delay=0.03)
    tts = TTS(scheduler_max_concurrency=2).from_engine(eng)
    tts.generate_speech(TTSRequest(text="x" * 40, speaker_files=["s.wav"], language="en"))
    assert eng.started[:2] == [0, 1]                            # lazily started, two at a time (App. B.15)
    sch = TwoPhaseScheduler(2, request_timeout=0.05)

    async def slow_first(_):
        await asyncio.sleep(1.0)

    async def run():
        async for _ in sch.run(None, "r", slow_first, None):
            pass
    with pytest.raises(TimeoutError):
        asyncio.new_event_loop().run_until_complete(run())


def test_split_requests_and_prepared_speaker():
    r = TTSRequest(text="a" * 250, speaker_files=["s.wav"], language="en")
    parts = TTS.split_requests(r, max_length=100)
    assert [len(p.text) for p in parts] == [100, 100, 50] and len({p.request_id for p in parts}) == 3
    tts = TTS().from_engine(FakeEngine(delay=0.001))
    fn = tts.loop.run_until_complete(tts.prepare_for_streaming_generation(r))
    r2 = TTSRequest(text="b" * 20, speaker_files=["s.wav"], language="en", context_partial_function=fn)
    assert tts.generate_speech(r2).array.shape == (8,)


def test_tokenizer_with_real_vocab_file(tmp_path):
    """the HF-`tokenizers` path (what XTTSTokenizerFast wraps, tokenizer.py:742-942): [lang] prefix, [SPACE] for blanks,
    [START]/[STOP] ids come from the vocab."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    vocab = {"[STOP]": 0, "[UNK]": 1, "[SPACE]": 2, "[START]": 3, "[en]": 4, "h": 5, "e": 6, "l": 7, "o": 8, "w": 9, "r": 10, "d": 11}
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[], unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tok.add_special_tokens(["[STOP]", "[UNK]", "[SPACE]", "[START]", "[en]"])     # matched whole, as in the real vocab
    f = tmp_path / "tokenizer.json"
    tok.save(str(f))
    t = XTTSTokenizer(6681, 402, str(f))
    assert not t.synthetic and (t.bos_token_id, t.eos_token_id) == (3, 0)
    ids = t.encode_chunk("Hello World", "en")
    assert ids[0] == 4 and 2 in ids and ids.count(7) == 3          # [en] h e l l o [SPACE] w o r l d
    assert t.batch_encode_with_split("hello world", "en") == [ids]


def test_speaker_slots_pin_evict_and_pending():
    """speakers.py: LRU eviction never takes a pinned or still-computing slot; stale (key, slot) pairs are detected."""
    from auralis_b200.speakers import SpeakerSlots, SpeakerSlotsFull
    t = SpeakerSlots(2)
    s_a, f_a, own = t.acquire("a")
    assert own and f_a is not None and not t.holds("a", s_a)          # not valid until the owner says so
    s_a2, f_a2, own2 = t.acquire("a")
    assert (s_a2, own2) == (s_a, False) and f_a2 is f_a                # a second request waits on the same future
    t.ready("a")
    assert f_a.result(timeout=1) is True and t.holds("a", s_a)
    assert t.acquire("a") == (s_a, None, False)                       # plain hit
    s_b, _, _ = t.acquire("b"); t.ready("b")
    assert s_b != s_a and t.pin("a", s_a) and t.pinned(s_a) == 1
    s_c, _, own_c = t.acquire("c"); t.ready("c")                       # evicts b (a is pinned although older)
    assert own_c and s_c == s_b and not t.holds("b", s_b) and t.holds("a", s_a)
    assert not t.pin("b", s_b)                                         # stale pair: caller must re-register
    assert t.pin("c", s_c)
    with pytest.raises(SpeakerSlotsFull):
        t.acquire("d")                                                 # both slots pinned
    t.unpin(s_c)
    s_d, f_d, _ = t.acquire("d")                                       # c evicted; d is still being computed ...
    assert s_d == s_c
    t.unpin(s_a)
    s_e, _, _ = t.acquire("e")                                         # ... so the only candidate is a
    assert s_e == s_a
    t.failed("d", RuntimeError("boom"))                                # owner failed: entry forgotten, waiters see the error
    with pytest.raises(RuntimeError):
        f_d.result(timeout=1)
    assert len(t) == 1 and t.acquire("d")[2] is True


def test_engine_repins_recycled_speaker_slot():
    """engine.py: a conditioning pair whose native slot was recycled is uploaded again instead of selecting the other
    speaker's voice; the slot stays pinned until the native completion is seen."""
    import asyncio
    from auralis_b200.config import XTTSDims
    from auralis_b200.engine import XTTSv2Engine
    from auralis_b200.speakers import SpeakerSlots

    class FakeNative:
        def __init__(self):
            self.slots = {}

        def set_speaker(self, slot, c, g):
            self.slots[slot] = (c.copy(), g.copy())

    dims = XTTSDims.small()
    eng = object.__new__(XTTSv2Engine)
    eng.dims, eng.native, eng._spk = dims, FakeNative(), SpeakerSlots(1)
    n, h, d = dims.gpt.n_cond_latents, dims.gpt.hidden, dims.voc.d_vector
    ca, ga = np.full((1, n, h), 1.0, np.float32), np.full((1, d, 1), 1.0, np.float32)
    cb, gb = np.full((1, n, h), 2.0, np.float32), np.full((1, d, 1), 2.0, np.float32)
    A = eng.register_speaker(ca, ga)
    assert A[0].slot == 0 and A[0].key == A[1].key
    B = eng.register_speaker(cb, gb)                                   # one slot only: A is evicted
    assert B[0].slot == 0 and eng.native.slots[0][0].flat[0] == 2.0

    async def go():
        slot = await eng._pin_speaker(*A)                              # stale pair -> uploaded again
        assert slot == 0 and eng.native.slots[0][0].flat[0] == 1.0 and eng._spk.pinned(0) == 1
        # B cannot take the slot while A's chunk is in flight; it gets it once the completion unpins
        task = asyncio.ensure_future(eng._pin_speaker(*B, timeout_s=5))
        await asyncio.sleep(0.05)
        assert not task.done() and eng.native.slots[0][0].flat[0] == 1.0
        eng._spk.unpin(0)
        assert await task == 0 and eng.native.slots[0][0].flat[0] == 2.0
    asyncio.new_event_loop().run_until_complete(go())


class _FakeNativeEngine:
    """Stands in for native.NativeEngine: same methods, a worker thread that 'synthesises' each chunk as a waveform
    filled with (speaker value at submission time) — so a voice swap or an unpinned slot shows up in the audio."""

    def __init__(self, dims, max_speakers, delay=0.01, fail_ids=()):
        import queue
        import threading
        self.dims, self.delay, self.fail_ids = dims, delay, set(fail_ids)
        self.spk, self.q, self.done = {}, queue.Queue(), queue.Queue()
        self.results, self.closed, self.cond_calls = {}, False, 0
        self.lock = threading.Lock()
        self.lib = type("L", (), {"xtts_last_error": staticmethod(lambda: b"fake failure")})()
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        from auralis_b200.native import XttsResult
        while not self.closed:
            try:
                sid, ids, slot, early = self.q.get(timeout=0.05)
            except Exception:
                continue
            time.sleep(self.delay)
            with self.lock:
                v = float(self.spk[slot][1].flat[0])          # the slot's CURRENT contents, like the real engine
            toks = np.arange(len(ids), dtype=np.int32)
            wav = np.full(8, v, np.float32) + np.arange(8, dtype=np.float32) * 1e-3
            skip = 0
            if early > 0 and len(ids) > early:                # first-audio piece, then the remainder (include/xtts_b200.h)
                skip = 3
                self.results.setdefault(sid, []).append((toks[:early], wav[:skip]))
                p = XttsResult()
                p.seq_id, p.status, p.n_tokens, p.n_samples = sid, 1, early, skip
                self.done.put(p)
                time.sleep(self.delay)
            self.results.setdefault(sid, []).append((toks, wav[skip:]))
            r = XttsResult()
            r.seq_id, r.status, r.n_tokens, r.n_samples = sid, (-2 if sid in self.fail_ids else 0), len(ids), 8 - skip
            self.done.put(r)

    def condition(self, slot, wav22, wav16, cond_len=30, chunk_len=4):
        time.sleep(0.02)
        self.cond_calls += 1
        v = float(np.round(np.abs(wav22).mean() * 1000))
        n, h, d = self.dims.gpt.n_cond_latents, self.dims.gpt.hidden, self.dims.voc.d_vector
        with self.lock:
            self.spk[slot] = (np.full((n, h), v, np.float32), np.full((d,), v, np.float32))

    def set_speaker(self, slot, c, g):
        with self.lock:
            self.spk[slot] = (np.array(c, np.float32).reshape(self.dims.gpt.n_cond_latents, -1), np.array(g, np.float32).reshape(-1))

    def get_speaker(self, slot):
        self.get_calls = getattr(self, "get_calls", 0) + 1
        with self.lock:
            c, g = self.spk[slot]
            return c.copy(), g.copy()

    def submit(self, sid, ids, slot, sp):
        self.q.put((sid, list(ids), slot, int(getattr(sp, "early_tokens", 0))))

    def cancel(self, sid):
        self.cancelled = getattr(self, "cancelled", []) + [sid]

    def poll(self, timeout_ms=50):
        try:
            return self.done.get(timeout=timeout_ms / 1000.0)
        except Exception:
            return None

    def fetch(self, r, want_wav=True, want_latents=False):
        toks, wav = self.results[r.seq_id].pop(0)             # oldest piece of that id first, like xtts_fetch
        return toks, wav, None

    def close(self):
        self.closed = True


def _host_engine(max_speakers=2, n_devices=1, **kw):
    """The real XTTSv2Engine host code (engine.py) over the fake native layer (`n_devices` fake GPUs)."""
    import threading
    from auralis_b200.config import XTTSDims
    from auralis_b200.engine import XTTSv2Engine
    from auralis_b200.speakers import SpeakerSlots
    dims = XTTSDims.small()
    eng = object.__new__(XTTSv2Engine)
    eng.dims, eng.precision, eng.max_concurrency, eng.max_speakers = dims, "fp32", 8, max_speakers
    eng.natives = [_FakeNativeEngine(dims, max_speakers, **kw) for _ in range(n_devices)]
    eng.native = eng.natives[0]
    eng.devices, eng.device_index = list(range(n_devices)), 0
    eng.tokenizer = XTTSTokenizer(dims.gpt.n_text_tokens, dims.gpt.max_text_tokens)
    eng.mel_bos_token_id, eng.mel_eos_token_id = dims.gpt.start_audio_token, dims.gpt.stop_audio_token
    eng._spks = [SpeakerSlots(max_speakers) for _ in range(n_devices)]
    eng._spk = eng._spks[0]
    eng._spk_arrays = {}
    eng.early_emit_tokens = 0
    eng._next_id, eng._id_lock, eng._waiters, eng._wlock = 1, threading.Lock(), {}, threading.Lock()
    eng._load = [0] * n_devices
    eng._stop = eng._paused = False
    eng._parked = 0
    eng._pollers = [threading.Thread(target=eng._poll_loop, args=(i,), daemon=True) for i in range(n_devices)]
    [t.start() for t in eng._pollers]
    return eng


def _wav(level):
    import io
    import wave
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(22050)
        w.writeframes((np.full(2205, level * 32767.0)).astype(np.int16).tobytes())
    return buf.getvalue()


def test_engine_host_path_many_speakers_few_slots():
    """6 requests x 3 speakers through TTS -> XTTSv2Engine -> (fake) native with only 2 speaker slots: every request
    must come back in its own speaker's voice, the same reference is conditioned once even when requested concurrently,
    and no slot stays pinned afterwards."""
    eng = _host_engine(max_speakers=2)
    tts = TTS(scheduler_max_concurrency=8).from_engine(eng)
    levels = [0.1, 0.2, 0.3]
    spk = [_wav(v) for v in levels]
    text = ("Sentence number one is here. " * 12).strip()              # several chunks per request
    reqs = [TTSRequest(text=text, speaker_files=spk[i % 3], language="en") for i in range(6)]
    outs = tts.generate_speech_batch(reqs)
    for i, o in enumerate(outs):
        want = float(np.round(levels[i % 3] * 1000))
        assert o.array.size >= 16 and np.all(np.abs(o.array - want) <= 1.0), (i, o.array[:4], want)
    assert eng.native.cond_calls <= 6                                   # 3 speakers (+ re-conditioning after an eviction), never one per request and retry
    assert eng.native.get_calls <= 12                                   # the concurrent first requests read back (+ retries after an eviction)
    assert all(eng._spk.pinned(s) == 0 for s in range(2)) and not eng._waiters
    # later calls are served from the host cache without touching the native layer ...
    n_get, n_cond = eng.native.get_calls, eng.native.cond_calls
    pairs = [tts.loop.run_until_complete(eng.get_audio_conditioning(b, 60, 30, 4)) for b in spk]      # the requests' parameters
    assert (eng.native.get_calls, eng.native.cond_calls) == (n_get, n_cond)
    # ... and a cached pair whose slot was recycled meanwhile (3 speakers, 2 slots) still synthesises in its own voice
    stale = [i for i, (c, _) in enumerate(pairs) if not eng._spk.holds(c.key, c.slot)]
    assert stale
    for i in stale:
        fn = partial_ctx(eng, *pairs[i])
        out = tts.generate_speech(TTSRequest(text="short one", speaker_files=spk[i], language="en", context_partial_function=fn))
        assert np.all(np.abs(out.array - float(np.round(levels[i] * 1000))) <= 1.0)
    tts.loop.run_until_complete(tts.shutdown())


def partial_ctx(eng, cond, g):
    from functools import partial
    return partial(eng.get_generation_context, gpt_cond_latent=cond, speaker_embeddings=g)


def test_engine_host_path_native_failure_propagates_and_unpins():
    eng = _host_engine(max_speakers=2, fail_ids={2})
    tts = TTS(scheduler_max_concurrency=8).from_engine(eng)
    text = ("Sentence number one is here. " * 12).strip()
    with pytest.raises(Exception, match="failed"):
        tts.generate_speech(TTSRequest(text=text, speaker_files=_wav(0.5), language="en"))
    deadline = time.time() + 2.0
    while time.time() < deadline and (eng._waiters or any(eng._spk.pinned(s) for s in range(2))):
        time.sleep(0.01)                                                # the other chunks finish in the background
    assert not eng._waiters and all(eng._spk.pinned(s) == 0 for s in range(2))
    tts.loop.run_until_complete(tts.shutdown())


def test_engine_early_first_audio_piece():
    """engine option early_emit_tokens: a streaming request gets its first chunk in two pieces (leading tokens first);
    the pieces concatenate to the unsplit audio, token counts add up, later chunks and non-streaming requests are not split."""
    eng = _host_engine(max_speakers=2)
    tts = TTS(scheduler_max_concurrency=8).from_engine(eng)
    text = ("Sentence number one is here. " * 12).strip()
    spk = _wav(0.4)
    plain = list(tts.generate_speech(TTSRequest(text=text, speaker_files=spk, language="en", stream=True)))
    eng.early_emit_tokens = 4
    early = list(tts.generate_speech(TTSRequest(text=text, speaker_files=spk, language="en", stream=True)))
    assert len(early) == len(plain) + 1                                 # only the first chunk is split
    assert early[0].array.size == 3 and early[0].token_length == 4     # the fake's partial piece
    assert early[0].token_length + early[1].token_length == plain[0].token_length
    np.testing.assert_array_equal(np.concatenate([c.array for c in early]), np.concatenate([c.array for c in plain]))
    whole = tts.generate_speech(TTSRequest(text=text, speaker_files=spk, language="en"))        # stream=False: never split
    np.testing.assert_array_equal(whole.array, np.concatenate([c.array for c in plain]))
    assert not eng._waiters and all(eng._spk.pinned(s) == 0 for s in range(2))
    tts.loop.run_until_complete(tts.shutdown())


def test_abandoned_chunk_is_cancelled_natively():
    """a chunk whose consumer goes away (closed generator / cancelled coroutine) must not keep decoding:
    engine.py calls xtts_cancel for it; a chunk consumed to its final result is not cancelled."""
    import asyncio
    from auralis_b200 import native
    eng = _host_engine(delay=0.05)
    loop = asyncio.new_event_loop()

    async def go():
        cond, g = await eng.get_audio_conditioning(_wav(0.2), 30, 6, 6)
        sp = native.Sampling()
        gen = eng._chunk_generator("r_0", [0, 5, 6, 1], cond, g, sp)
        task = asyncio.ensure_future(gen.__anext__())
        await asyncio.sleep(0.01)                      # submitted, not finished
        task.cancel()
        try:
            await task
        except asyncio.CancelledError:
            pass
        await gen.aclose()
        assert getattr(eng.native, "cancelled", []) == [1]
        gen2 = eng._chunk_generator("r_1", [0, 5, 6, 1], cond, g, sp)
        out = [o async for o in gen2]
        assert len(out) == 1 and getattr(eng.native, "cancelled", []) == [1]     # consumed to the end: no cancel
    loop.run_until_complete(go())
    loop.run_until_complete(eng.shutdown())


def test_in_process_data_parallel_engines():
    """`XTTSv2Engine(devices=[...])`: chunks are spread over the engines by work in flight, the speaker computed on GPU 0
    is uploaded to the other GPU from the host arrays on first use (no voice swap), the request comes back in order, and
    the per-GPU load counters return to zero."""
    from auralis_b200 import TTS, TTSRequest
    eng = _host_engine(max_speakers=2, n_devices=2, delay=0.01)
    tts = TTS(scheduler_max_concurrency=64).from_engine(eng)
    text = " ".join(f"Sentence number {i} talks about nothing in particular, at some length, to fill a chunk." for i in range(14))
    n_chunks = len(eng.prepare_text_tokens(text, "en"))
    assert n_chunks >= 6
    seen = [[], []]
    for d, ne in enumerate(eng.natives):
        orig = ne.submit

        def rec(sid, ids, slot, sp, _o=orig, _d=d):
            seen[_d].append(sid)
            return _o(sid, ids, slot, sp)
        ne.submit = rec
    out = tts.generate_speech(TTSRequest(text=text, speaker_files=_wav(0.2), language="en"))
    assert out.array.shape[0] == 8 * n_chunks
    assert len(seen[0]) + len(seen[1]) == n_chunks and min(len(seen[0]), len(seen[1])) >= n_chunks // 2 - 1
    v = float(np.round(0.2 * 32767) / 32767 * 1000 // 1)                    # the fake's voice value for this reference
    assert np.allclose(np.round(out.array.reshape(n_chunks, 8)[:, 0]), np.round(out.array[0]))      # one voice on both GPUs
    assert eng.natives[0].cond_calls == 1 and eng.natives[1].cond_calls == 0  # conditioned once, uploaded to the 2nd GPU
    assert eng._load == [0, 0] and not eng._waiters
    tts.loop.run_until_complete(tts.shutdown())


def test_generation_metrics_tracker_matches_the_reference_rule():
    """metrics.py restates performance.py:12-151: one update per chunk that has a start_time, a log line (and a window reset)
    once log_interval has passed, the three rates from the window totals; native counters are appended when available."""
    from auralis_b200.metrics import TTSMetricsTracker, track
    from auralis_b200 import TTSOutput
    t = TTSMetricsTracker(log_interval=0.0)
    t.window_start -= 2.0

    class Eng:
        def stats(self):
            return {"decode_steps": 7, "kernel_launches": 1500, "gpt_ms": 12.0, "vocoder_ms": 3.0}
    out = TTSOutput(array=np.zeros(24000, np.float32), token_length=21, start_time=time.time())
    assert t.update_metrics(21, 1.0) is True and t.window_requests == 1 and t.window_tokens == 21
    assert 9 < t.tokens_per_second < 11 and 0.4 < t.requests_per_second < 0.6 and 1900 < t.ms_per_second_of_audio < 2200
    line = t.line(Eng())
    assert line.startswith("Generation metrics | Throughput: ") and "tokens/s" in line and "7 decode steps" in line
    t.reset_window()
    assert t.window_requests == 0 and t.window_tokens == 0
    track(out, Eng(), t)                                    # logs + resets (interval 0)
    assert t.window_requests == 0
    track(TTSOutput(array=np.zeros(10, np.float32), token_length=1, start_time=None), None, t)   # no start_time: not counted
    assert t.window_requests == 0
