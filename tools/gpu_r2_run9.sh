#!/bin/bash
# round 2, run 9: CTA-pair GEMM with TMA-store epilogue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=12 run r2i_gemm 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -s -k "gemm"
TAIL=40 run r2i_gemm_probe 300 python tools/gemm_probe.py
run r2i_gpt  600 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_bench_regime.py tests/test_gpu_conditioning.py tests/test_gpu_api.py -q --no-header -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2cta -s 1 -c 1 -o gpurun_out/r2i_prof_gemm2cta -f python tools/gemm_probe.py > gpurun_out/r2i_ncu.log 2>&1; echo "ncu exit $?"
