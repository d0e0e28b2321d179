"""Where does a decode step's wall time go?  Same engine, same batch, toggling CUDA graphs / split-K."""
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
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 41
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 96
eng = native.NativeEngine(dims, precision=1, max_batch=max(NB, 8), max_speakers=2)
eng.load_state(gs, cs); eng.set_speaker(0, cond.numpy(), dv.numpy())
rng = np.random.RandomState(1)
def jobs(nt):
    return [(i, [0] + rng.randint(2, 6000, size=78).tolist() + [1], 0,
             native.Sampling(temperature=0.75, top_p=0.85, top_k=50, max_tokens=nt, seed=1, seq_seed=i, vocode=False)) for i in range(NB)]
eng.run_batch(jobs(8), timeout_s=600, want_wav=False)     # warm
def run(label, **opts):
    for k, v in opts.items(): eng.set_option(k, v)
    # two runs with different token counts: the difference isolates the decode steps from prefill / fixed costs
    t = []
    for nt in (NT // 2, NT):
        eng.set_option("reset_stats", 0)
        t0 = time.time(); eng.run_batch(jobs(nt), timeout_s=600, want_wav=False); t.append(time.time() - t0)
    st = eng.stats()
    print(f"{label:34s} batch {NB}: {1e3 * (t[1] - t[0]) / (NT - NT // 2):6.3f} ms/decode-step  (runs {t[0]*1e3:.0f} / {t[1]*1e3:.0f} ms; engine gpt_ms {st.gpt_ms:.0f})", flush=True)
eng.set_option("microbatch_min_rows", 8)
for nmb in (1, 2, 3, 4):
    run(f"graphs + split-K, {nmb} branch(es)", cuda_graphs=1, splitk=1, microbatches=nmb)
run("eager  + split-K, 1 branch", cuda_graphs=0, splitk=1, microbatches=1)
run("eager  + split-K, 2 branches", cuda_graphs=0, splitk=1, microbatches=2)
if os.environ.get("STEP_PROBE_NOSPLITK"):
    run("graphs, no split-K", cuda_graphs=1, splitk=0, microbatches=1)
eng.set_option("cuda_graphs", 1); eng.set_option("splitk", 1); eng.set_option("microbatches", 1)
eng.set_option("profile", 1)
eng.run_batch(jobs(NT // 2), timeout_s=600, want_wav=False)
prof = eng.kernel_profile(); eng.set_option("profile", 0)
steps = NT // 2 - 1
tot = sum(v["ms"] for v in prof.values())
print(f"eager profile: kernel time {tot:.1f} ms over prefill + {steps} steps")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"   {k:28s} {v['ms']:8.2f} ms  {v['launches']:6d} launches  {1e3*v['ms']/v['launches']:7.1f} us avg")
