#!/bin/bash
# round 2, run 27: decode GEMMs with a ring that fills the SM (one CTA per SM, 1.6 instead of 3.2 ring passes per GEMM)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu -k "capped_attention or microbatch" > gpurun_out/r2aa_test.log 2>&1; echo "test exit $?"; tail -n 6 gpurun_out/r2aa_test.log
STEP_CFGS='[{"gemm_deep_ring":0,"gemm_bn":0,"microbatches":2},{"gemm_deep_ring":1,"gemm_bn":0,"microbatches":2},{"gemm_deep_ring":1,"gemm_bn":64,"microbatches":2},{"gemm_deep_ring":0,"gemm_bn":64,"microbatches":2},{"gemm_deep_ring":0,"gemm_bn":0,"microbatches":1},{"gemm_deep_ring":1,"gemm_bn":0,"microbatches":1},{"gemm_deep_ring":1,"gemm_bn":64,"microbatches":1},{"gemm_deep_ring":1,"gemm_bn":0,"microbatches":3}]' timeout 900 python tools/attn_probe.py 163 2 > gpurun_out/r2aa_probe.log 2>&1; echo "probe exit $?"; tail -n 12 gpurun_out/r2aa_probe.log
