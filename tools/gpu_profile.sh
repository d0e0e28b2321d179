#!/bin/bash
# ncu evidence: (1) launch list with device time per launch for a reduced bench step, (2) full-metric captures of the
# dominant kernels (tensor-core conv, paged decode attention, tcgen05 GEMM)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_kernels.py 32 6 605 > gpurun_out/launches.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 20 -c 3 -o gpurun_out/prof_conv1d_tc \
    python tools/profile_kernels.py 2 2 605 > gpurun_out/prof_conv.log 2>&1
echo "conv capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 30 -c 2 -o gpurun_out/prof_attn_decode \
    python tools/profile_kernels.py 64 12 8 > gpurun_out/prof_attn.log 2>&1
echo "attn capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 200 -c 4 -o gpurun_out/prof_gemm_tc \
    python tools/profile_kernels.py 64 4 8 > gpurun_out/prof_gemm.log 2>&1
echo "gemm capture exit $?"
ls -la gpurun_out/*.ncu-rep
