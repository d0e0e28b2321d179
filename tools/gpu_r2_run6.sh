#!/bin/bash
# round 2, run 6: counter dependencies along the decode chain, coalesced CTA-pair GEMM epilogue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=12 run r2f_gemm 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -s -k "gemm"
run r2f_gpt  600 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_windows.py tests/test_gpu_api.py -q --no-header -x
if grep -q "failed\|rror" gpurun_out/r2f_gpt.log; then echo "gpt tests failed with dep_flags"; fi
TAIL=20 run r2f_probe 600 python tools/stagger_probe.py 163
TAIL=300 run r2f_trace2 300 python tools/trace_step.py 163 415 2
