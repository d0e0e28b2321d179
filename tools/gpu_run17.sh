#!/bin/bash
# run 17: ncu --set full of the decode-shaped GEMMs (M = 163) and the split-K reduce + LayerNorm kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_bf16_tc_kernel<32>' -s 900 -c 4 \
    -o gpurun_out/prof17_gemm_decode -f python tools/profile_kernels.py 163 6 8 > gpurun_out/prof17_gemm.log 2>&1; echo "gemm exit $?"; tail -n 3 gpurun_out/prof17_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:residual_reduce_ln -s 400 -c 2 \
    -o gpurun_out/prof17_ln -f python tools/profile_kernels.py 163 6 8 > gpurun_out/prof17_ln.log 2>&1; echo "ln exit $?"; tail -n 3 gpurun_out/prof17_ln.log
