#!/bin/bash
# round 2, run 25: CTA-pair GEMM with 16 epilogue warps — tests, shapes x variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" > gpurun_out/r2y_test.log 2>&1; echo "test exit $?"; tail -n 5 gpurun_out/r2y_test.log
timeout 600 python tools/gemm_probe.py > gpurun_out/r2y_gemm_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/r2y_gemm_probe.log
