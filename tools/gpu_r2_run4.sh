#!/bin/bash
# round 2, run 4: CTA-pair GEMM, co-major transposed-conv epilogue, timeline of the decode step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=30 run r2d_gemm 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -s -k "gemm"
run r2d_voc  300 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_windows.py -q --no-header -x
run r2d_vprobe  300 python tools/voc_probe.py 605
TAIL=330 run r2d_trace2  300 python tools/trace_step.py 163 415 2
TAIL=60 run r2d_trace1  300 python tools/trace_step.py 163 415 1
run r2d_suite 600 python -m pytest tests -q -m gpu --no-header --deselect tests/test_gpu_bench_regime.py
