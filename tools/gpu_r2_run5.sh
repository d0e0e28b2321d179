#!/bin/bash
# round 2, run 5: branch stagger sweep; ncu capture of the CTA-pair GEMM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=20 run r2e_stagger 400 python tools/stagger_probe.py 163
cat > /tmp/gemm_probe.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from auralis_b200 import native
from auralis_b200.config import XTTSDims
eng = native.NativeEngine(XTTSDims.small(), precision=1, max_batch=4, max_speakers=2)
rng = np.random.RandomState(1)
for (M, N, K) in ((4096, 4096, 1024), (2048, 3072, 1024), (2048, 4096, 1024), (2048, 1024, 4096)):
    A = rng.randn(M, K).astype(np.float32); W = (rng.randn(N, K) * 0.05).astype(np.float32)
    _, ms = eng.debug_gemm(1, A, W, None, None, False, iters=10)
    print(f"2cta M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
PY
run r2e_gemm_shapes 200 python /tmp/gemm_probe.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2cta -s 2 -c 2 -o gpurun_out/r2e_prof_gemm2cta -f python /tmp/gemm_probe.py > gpurun_out/r2e_ncu.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/r2e_ncu.log
ls -la gpurun_out/*.ncu-rep
