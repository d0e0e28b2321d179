"""Decode-step time at the bench's batch (163 rows) and context (~500 cached tokens) as a function of the branch stagger:
ms per decode step = (run to 400 tokens - run to 300 tokens) / 100.   python tools/stagger_probe.py [rows]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from auralis_b200 import native
from auralis_b200.config import XTTSDims
from auralis_b200.weights import synth_state

dims = XTTSDims.full()
gs, cs = synth_state(dims, 1234)
g = torch.Generator().manual_seed(500)
cond = torch.randn(32, 1024, generator=g); dv = torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0)
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 163
eng = native.NativeEngine(dims, precision=1, max_batch=max(NB, 8), max_speakers=2)
eng.load_state(gs, cs); eng.set_speaker(0, cond.numpy(), dv.numpy())
rng = np.random.RandomState(1)
ids = [[0] + rng.randint(2, 6000, size=78).tolist() + [1] for _ in range(NB)]
def jobs(nt):
    return [(i, ids[i], 0, native.Sampling(temperature=0.75, top_p=0.85, top_k=50, max_tokens=nt, seed=1, seq_seed=i, vocode=False, stop_token=4095)) for i in range(NB)]
eng.run_batch(jobs(8), timeout_s=600, want_wav=False)
def run(label, **opts):
    for k, v in opts.items(): eng.set_option(k, v)
    eng.run_batch(jobs(12), timeout_s=600, want_wav=False)          # graph capture for this configuration
    best = 1e9
    for rep in range(2):
        t = []
        for nt in (300, 400):
            t0 = time.time(); eng.run_batch(jobs(nt), timeout_s=600, want_wav=False); t.append(time.time() - t0)
        best = min(best, 1e3 * (t[1] - t[0]) / 100)
    print(f"{label:48s} {best:6.3f} ms/decode-step", flush=True)
for extra in sys.argv[2:]:
    eng.set_option(extra.split("=")[0], int(extra.split("=")[1]))
run("warm", microbatches=2)
import itertools
cfgs = [(aw, mb) for aw in (4, 2, 1) for mb in (2, 1)]
best = {}
def measure(**opts):
    for k, v in opts.items(): eng.set_option(k, v)
    eng.run_batch(jobs(12), timeout_s=600, want_wav=False)
    t = []
    for nt in (300, 400):
        t0 = time.time(); eng.run_batch(jobs(nt), timeout_s=600, want_wav=False); t.append(time.time() - t0)
    return 1e3 * (t[1] - t[0]) / 100
for rep in range(3):                                         # interleaved repeats: clock / thermal drift hits every config alike
    for (aw, mb) in (cfgs if rep % 2 == 0 else cfgs[::-1]):
        v = measure(attn_warps=aw, microbatches=mb, dep_flags=0, branch_stagger_us=0)
        best.setdefault((aw, mb), []).append(v)
for k, v in best.items():
    print(f"attn_warps {k[0]}, {k[1]} branch(es): " + " ".join(f"{x:6.3f}" for x in v) + f"   min {min(v):6.3f} ms/decode-step", flush=True)
eng.close()
