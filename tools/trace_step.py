"""Timeline of ONE decode step as it really runs (PDL, concurrent branches): %globaltimer stamps from the kernels themselves
(xtts_debug_trace).   python tools/trace_step.py [rows] [ctx_tokens] [microbatches] [extra engine opts k=v ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from auralis_b200 import native
from auralis_b200.config import XTTSDims
from auralis_b200.weights import synth_state

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 163
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 300
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dims = XTTSDims.full()
gs, cs = synth_state(dims, 1234)
g = torch.Generator().manual_seed(500)
cond = torch.randn(32, 1024, generator=g); dv = torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0)
eng = native.NativeEngine(dims, precision=1, max_batch=max(rows, 8), max_speakers=2)
eng.load_state(gs, cs); eng.set_speaker(0, cond.numpy(), dv.numpy())
eng.set_option("microbatches", mb)
for kv in sys.argv[4:]:
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
rng = np.random.RandomState(1)
ntok = max(8, ctx - 113 + 24)
jobs = [(i, [0] + rng.randint(2, 6000, size=78).tolist() + [1], 0,
         native.Sampling(temperature=0.75, top_p=0.85, top_k=50, max_tokens=ntok, seed=1, seq_seed=i, vocode=False, stop_token=4095)) for i in range(rows)]
eng.set_option("hold_admission", 1)
for j in jobs:
    eng.submit(*j)
eng.set_option("hold_admission", 0)
import time
# let the decode reach the wanted context, then trace ~20 steps
while eng.stats().decode_steps < ntok - 30:
    time.sleep(0.005)
eng.trace_start()
s0 = eng.stats().decode_steps
time.sleep(0.06)
tr = eng.trace_stop()
s1 = eng.stats().decode_steps
n = 0
while n < rows:
    r = eng.poll(1000)
    if r is not None:
        eng.fetch(r, want_wav=False); n += 1
names = {1: "gemm", 2: "attn", 3: "reduce_ln", 4: "ln", 5: "head", 6: "sample", 7: "rows", 8: "conv"}
print(f"traced {len(tr)} records over ~{s1 - s0} steps; rows {rows}, branches {mb}")
# one step = from a 'rows' kernel entry to the next
starts = [i for i in range(len(tr)) if tr[i, 1] == 7 and tr[i, 2] == 0 and tr[i, 3] == 0]
if len(starts) >= 4:
    a, b = starts[2], starts[3]
    step = tr[a:b]
    t0 = step[0, 0]
    print(f"step duration {(tr[b, 0] - t0) / 1e3:.1f} us, {len(step)} records")
    # per-kernel-instance durations (first CTA entry -> last CTA exit), in launch order, first 2 layers
    ev = [(int(t - t0), names.get(int(k), str(k)), int(ph), int(last), int(grid)) for t, k, ph, last, grid in step]
    lim = 0
    for e in ev:
        print(f"{e[0] / 1e3:9.2f} us  {e[1]:10s} phase {e[2]} {'last' if e[3] else 'first'} grid {e[4]}")
        lim += 1
        if lim > 260:
            break
    # aggregate: per kernel type, mean (entry first -> exit last) and mean wait (entry -> dependency resolved)
    for kid, nm in names.items():
        ent = step[(step[:, 1] == kid) & (step[:, 2] == 0) & (step[:, 3] == 0)][:, 0]
        dep = step[(step[:, 1] == kid) & (step[:, 2] == 1) & (step[:, 3] == 0)][:, 0]
        ext = step[(step[:, 1] == kid) & (step[:, 2] == 2) & (step[:, 3] == 1)][:, 0]
        if len(ent):
            msg = f"{nm:10s} n={len(ent):4d}"
            if len(dep) == len(ent):
                msg += f"  entry->dep-resolved {np.mean(dep - ent) / 1e3:6.2f} us"
            if len(ext) == len(ent):
                msg += f"  entry->exit(last CTA) {np.mean(np.sort(ext) - np.sort(ent)) / 1e3:6.2f} us"
            print(msg)
eng.close()
