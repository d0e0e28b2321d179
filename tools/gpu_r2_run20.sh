#!/bin/bash
# round 2, run 20: ncu --set full of the two decode-attention kernels at the bench's batch and ~400 cached tokens
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
XTTS_OPTS=attn_bulk=1,microbatches=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_decode_bulk -s 8400 -c 2 -f -o gpurun_out/r2t_attn_bulk python tools/profile_kernels.py 163 300 8 > gpurun_out/r2t_ncu_bulk.log 2>&1; echo "ncu bulk exit $?"; tail -n 3 gpurun_out/r2t_ncu_bulk.log
XTTS_OPTS=attn_bulk=0,microbatches=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_decode_kernel -s 8400 -c 2 -f -o gpurun_out/r2t_attn_reg python tools/profile_kernels.py 163 300 8 > gpurun_out/r2t_ncu_reg.log 2>&1; echo "ncu reg exit $?"; tail -n 3 gpurun_out/r2t_ncu_reg.log
ls -la gpurun_out/*.ncu-rep
