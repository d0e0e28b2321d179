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
