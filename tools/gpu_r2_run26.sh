#!/bin/bash
# round 2, run 26: register attention, first page early / two pages in flight — parity tests, step-time probe
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu -k "bulk_copy or capped_attention or microbatch" > gpurun_out/r2z_test.log 2>&1; echo "test exit $?"; tail -n 6 gpurun_out/r2z_test.log
timeout 900 python tools/attn_probe.py 163 2 > gpurun_out/r2z_probe.log 2>&1; echo "probe exit $?"; tail -n 16 gpurun_out/r2z_probe.log
