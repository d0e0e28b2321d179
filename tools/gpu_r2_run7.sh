#!/bin/bash
# round 2, run 7: 8-warp decode attention, counter dependencies with back-off, CTA-pair GEMM with two epilogue warpgroups
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=12 run r2g_gemm 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -s -k "gemm"
run r2g_gpt  600 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_bench_regime.py -q --no-header -x
TAIL=20 run r2g_probe 600 python tools/stagger_probe.py 163
