"""Quick full-size timing probe (not the bench): vocoder on one 605-latent chunk, GPT batch decode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from auralis_b200 import native
from auralis_b200.config import XTTSDims
from auralis_b200.weights import synth_state

dims = XTTSDims.full()
t0 = time.time(); gs, cs = synth_state(dims, 1234); print("synth", time.time() - t0, flush=True)
g = torch.Generator().manual_seed(500)
cond = torch.randn(32, 1024, generator=g); dv = torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0)
modes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1").split(",")]
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 32
NT = int(sys.argv[3]) if len(sys.argv) > 3 else 64
for prec in modes:
    t0 = time.time()
    eng = native.NativeEngine(dims, precision=prec, max_batch=max(NB, 4), max_speakers=2)
    eng.load_state(gs, cs); eng.set_speaker(0, cond.numpy(), dv.numpy())
    print(f"[prec {prec}] engine up in {time.time() - t0:.1f}s", flush=True)
    lat = np.random.RandomState(0).randn(605, 1024).astype(np.float32)
    eng.vocode(lat, 0)
    t0 = time.time(); n = 3
    for _ in range(n): wav = eng.vocode(lat, 0)
    dt = (time.time() - t0) / n
    print(f"[prec {prec}] vocoder 605 latents: {dt*1e3:.1f} ms -> {wav.size/24000/dt:.1f} audio-s/s ({621.36e6*2634/dt/1e12:.2f} TFLOP/s)", flush=True)
    rng = np.random.RandomState(1)
    for nb in (1, NB):
        jobs = [(i, [0] + rng.randint(2, 6000, size=60).tolist() + [1], 0,
                 native.Sampling(temperature=0.75, top_p=0.85, top_k=50, max_tokens=NT, seed=1, seq_seed=i, vocode=False)) for i in range(nb)]
        eng.set_option("reset_stats", 0)
        t0 = time.time(); res = eng.run_batch(jobs, timeout_s=600, want_wav=False); dt = time.time() - t0
        st = eng.stats()
        print(f"[prec {prec}] GPT batch {nb} x {NT} tokens: {dt*1e3:.0f} ms -> {nb*NT/dt:.0f} tok/s; "
              f"{st.decode_steps} steps, {dt*1e3/max(1,st.decode_steps):.2f} ms/step, launches {st.kernel_launches}", flush=True)
    eng.close()
