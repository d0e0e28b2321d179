"""BASELINE.json configs[3]: a Harry-Potter-length synthetic book (~500 k chars) split by the scheduler into requests and
chunks, sharded data-parallel over the ranks (LPT over request length, no data-path collective), waveforms gathered on
rank 0 over NCCL; reports end-to-end wall-clock and aggregate audio-s/s through the public API with host buffers.

    python tools/bench_book.py [chars] [chars_per_request]                         # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_book.py ...

The reference's CPU path on the same text is bench.py's `cpu_baseline` / `--impl reference` figure (audio-s/s on the
box's host cores): book_audio_s / that figure is its wall-clock for the same book.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_text, synthetic_wav_bytes, SEED
from auralis_b200 import TTS, TTSRequest, parallel
from auralis_b200.config import XTTSDims
from auralis_b200.engine import XTTSv2Engine
from auralis_b200.weights import synth_state

total_chars = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
per_req = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000
rank, world, local = parallel.init_from_env()
dims = XTTSDims.full()
state = synth_state(dims, SEED)
eng = XTTSv2Engine(dims, state[0], state[1], device=local, precision="fp16", max_concurrency=256, max_speakers=4)
tts = TTS(scheduler_max_concurrency=100000).from_engine(eng)
spk = synthetic_wav_bytes(6.0, 120.0, 7)
tts.loop.run_until_complete(eng.get_audio_conditioning(spk, 60, 30, 4))

# the "book": chapters of per_req chars (every rank builds the same list; LPT assigns them)
n_req = (total_chars + per_req - 1) // per_req
texts = [make_text(min(per_req, total_chars - i * per_req), 9000 + i) for i in range(n_req)]
costs = [float(len(t)) for t in texts]


def synth(indices):
    reqs = [TTSRequest(text=texts[i], speaker_files=spk, language="en", seed=SEED + i) for i in indices]
    outs = tts.generate_speech_batch(reqs)
    return {i: o.array for i, o in zip(indices, outs)}


def barrier():
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


# warm-up (graph capture for the steady batch sizes, pools) on a small slice of this rank's share
synth(parallel.lpt_partition(costs, world)[rank][:1])
barrier()
t0 = time.perf_counter()
waves = parallel.run_sharded(texts, costs, synth, torch.device("cuda", local), dst=0)
barrier()
wall = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    wall = float(t[0])
if rank == 0:
    audio_s = sum(w.shape[0] for w in waves) / 24000.0
    print(json.dumps({"config": f"cfg4: synthetic book, {sum(len(t) for t in texts)} chars in {n_req} requests of <= {per_req} chars, "
                                f"{world} GPU(s) data-parallel (LPT shards), waveforms gathered on rank 0",
                      "n_gpus": world, "wall_s": wall, "audio_s": audio_s, "audio_hours": audio_s / 3600.0,
                      "audio_s_per_s": audio_s / wall,
                      "note": "random-init weights never emit the stop token: every <=250-char chunk runs the full 605 tokens"}))
tts.loop.run_until_complete(tts.shutdown())
if world > 1:
    torch.distributed.destroy_process_group()
