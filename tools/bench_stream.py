"""BASELINE.json configs[2]: continuous-batching stream — 256 mixed-length (128..4096 chars) requests, 4 shared speakers,
1 GPU; reports p50/p99 time-to-first-audio (first streamed chunk of each request) and aggregate audio-s/s.
All requests are submitted at t=0 through the public async streaming API (`generate_speech_async(stream=True)`).
    python tools/bench_stream.py [n_requests] [max_concurrency] [early_emit_tokens]
early_emit_tokens > 0 (e.g. 58): the first chunk of every request delivers the audio of its first tokens as soon as they are
decoded (engine option) — 58 tokens = 2.7 s of audio, about what the rest of a 605-token chunk takes to decode.
"""
import asyncio, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_text, synthetic_wav_bytes, SEED
from auralis_b200 import TTS, TTSRequest
from auralis_b200.config import XTTSDims
from auralis_b200.engine import XTTSv2Engine
from auralis_b200.weights import synth_state

n_req = int(sys.argv[1]) if len(sys.argv) > 1 else 256
conc = int(sys.argv[2]) if len(sys.argv) > 2 else 256
early = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dims = XTTSDims.full()
state = synth_state(dims, SEED)
eng = XTTSv2Engine(dims, state[0], state[1], device=0, precision="fp16", max_concurrency=conc, max_speakers=8,
                   early_emit_tokens=early)
tts = TTS(scheduler_max_concurrency=100000).from_engine(eng)
spk = [synthetic_wav_bytes(6.0, 100.0 + 25.0 * i, 7 + i) for i in range(4)]
loop = tts.loop
for b in spk:
    loop.run_until_complete(eng.get_audio_conditioning(b, 60, 30, 4))
rng = np.random.RandomState(11)
lens = rng.randint(128, 4097, size=n_req)
texts = [make_text(int(n), 5000 + i) for i, n in enumerate(lens)]

async def one(i, t0):
    req = TTSRequest(text=texts[i], speaker_files=spk[i % 4], language="en", stream=True, seed=SEED + i)
    gen = await tts.generate_speech_async(req)
    first, samples, chunks = None, 0, 0
    async for chunk in gen:
        if first is None:
            first = time.perf_counter() - t0
        samples += chunk.array.shape[0]; chunks += 1
    return first, samples, chunks, time.perf_counter() - t0

async def main():
    t0 = time.perf_counter()
    res = await asyncio.gather(*[one(i, t0) for i in range(n_req)])
    return res, time.perf_counter() - t0

# warm-up on a small slice (graph capture, pools)
loop.run_until_complete(asyncio.gather(*[one(i, time.perf_counter()) for i in range(min(8, n_req))]))
res, wall = loop.run_until_complete(main())
ttfa = np.array([r[0] for r in res]); done = np.array([r[3] for r in res])
audio_s = sum(r[1] for r in res) / 24000.0
out = {"config": f"cfg3: {n_req} requests, {int(lens.min())}..{int(lens.max())} chars, 4 speakers, 1 GPU, max_concurrency {conc}, early_emit_tokens {early}",
       "chunks": int(sum(r[2] for r in res)), "wall_s": wall, "audio_s": audio_s, "audio_s_per_s": audio_s / wall,
       "ttfa_p50_s": float(np.percentile(ttfa, 50)), "ttfa_p99_s": float(np.percentile(ttfa, 99)), "ttfa_min_s": float(ttfa.min()),
       "request_latency_p50_s": float(np.percentile(done, 50)), "request_latency_p99_s": float(np.percentile(done, 99)),
       "note": "random-init weights never emit the stop token, so every <=250-char chunk is 605 tokens = 28.1 s of audio"}
print(json.dumps(out))
loop.run_until_complete(tts.shutdown())
