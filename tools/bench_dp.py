"""Data parallelism INSIDE the product: `XTTSv2Engine(devices=[0..N-1])` — one process, one native engine (full replica, own
scheduler thread and streams) per GPU, every text chunk dispatched to the engine with the least work in flight, results
re-assembled in request order by the façade.  Waveforms go GPU -> pinned host memory of this process directly: there is no
gather step at all (the NCCL all-gather of the torchrun layout exists because there the consumer is another process).

    python tools/bench_dp.py <n_gpus> [mode ...]       modes: weak strong book   (default: all three)

  weak    32 x 1k-char requests per GPU        (bench.py's workload, for comparison with the torchrun numbers)
  strong  64 x 1k-char requests in TOTAL       (the north-star headline shape; 8 requests ~ 41 chunks per GPU at N = 8)
  book    BASELINE cfg4: ~500 k characters as 100 requests of 5 000 chars, balanced chunk by chunk (~2 000 chunks)
Every number is end to end through the public API (host text in, host float32 waveforms out), host wall clock."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_text, synthetic_wav_bytes, SEED
from auralis_b200 import TTS, TTSRequest
from auralis_b200.config import XTTSDims
from auralis_b200.engine import XTTSv2Engine
from auralis_b200.weights import synth_state

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
modes = sys.argv[2:] or ["weak", "strong", "book"]
dims = XTTSDims.full()
state = synth_state(dims, SEED)
t0 = time.perf_counter()
eng = XTTSv2Engine(dims, state[0], state[1], devices=list(range(n_gpus)), precision="fp16", max_concurrency=256,
                   max_speakers=8, voc_segment=96, tune_malloc=True)
print(f"[bench_dp] {n_gpus} engines up in {time.perf_counter() - t0:.1f}s", file=sys.stderr, flush=True)
tts = TTS(scheduler_max_concurrency=100000).from_engine(eng)
spk = [synthetic_wav_bytes(6.0, 100.0 + 25.0 * i, 7 + i) for i in range(4)]
for b in spk:
    tts.loop.run_until_complete(eng.get_audio_conditioning(b, 60, 30, 4))


def run(texts, seed0):
    reqs = [TTSRequest(text=t, speaker_files=spk[i % 4], language="en", temperature=0.75, top_p=0.85, top_k=50,
                       repetition_penalty=5.0, seed=seed0 + i) for i, t in enumerate(texts)]
    t0 = time.perf_counter()
    outs = tts.generate_speech_batch(reqs)
    dt = time.perf_counter() - t0
    return sum(o.array.shape[0] for o in outs) / 24000.0, dt


out = {"n_gpus": n_gpus, "layout": "one process, one native engine per GPU, chunk-level least-loaded dispatch, no gather"}
if "weak" in modes:
    texts = [make_text(1000, i) for i in range(32 * n_gpus)]
    run(texts, 1)
    a, dt = run(texts, 2)
    out["weak"] = {"requests": len(texts), "audio_s": a, "wall_s": dt, "audio_s_per_s": a / dt}
    print(f"[bench_dp] weak: {a / dt:.1f} audio-s/s", file=sys.stderr, flush=True)
if "strong" in modes:
    texts = [make_text(1000, 100 + i) for i in range(64)]
    run(texts, 3)
    a, dt = run(texts, 4)
    out["strong"] = {"requests": 64, "audio_s": a, "wall_s": dt, "audio_s_per_s": a / dt}
    print(f"[bench_dp] strong: {a / dt:.1f} audio-s/s in {dt:.2f}s", file=sys.stderr, flush=True)
if "book" in modes:
    texts = [make_text(5000, 9000 + i) for i in range(100)]
    run(texts[:2], 5)
    a, dt = run(texts, 6)
    out["book"] = {"chars": sum(len(t) for t in texts), "requests": len(texts), "audio_s": a, "audio_hours": a / 3600.0, "wall_s": dt,
                   "audio_s_per_s": a / dt}
    print(f"[bench_dp] book: {a / 3600:.2f} h of audio in {dt:.2f}s", file=sys.stderr, flush=True)
st = eng.stats()
out["native_counters"] = {k: st[k] for k in ("tokens_generated", "samples_generated", "decode_steps", "kernel_launches")}
out["per_gpu_load_left"] = list(eng._load)
print(json.dumps(out))
tts.loop.run_until_complete(tts.shutdown())
