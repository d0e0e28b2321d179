#!/bin/bash
# round 2, run 21: bulk attention with 4 / 8 / 16 consumer warps — parity test, step-time probe, timeline
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu -k "bulk_copy" > gpurun_out/r2u_test.log 2>&1; echo "test exit $?"; tail -n 6 gpurun_out/r2u_test.log
timeout 900 python tools/attn_probe.py 163 2 > gpurun_out/r2u_probe.log 2>&1; echo "probe exit $?"; tail -n 20 gpurun_out/r2u_probe.log
timeout 300 python tools/trace_step.py 163 415 1 attn_bulk=1 attn_warps=16 attn_stages=16 > gpurun_out/r2u_trace_bulk1.log 2>&1; echo "trace exit $?"; tail -n 9 gpurun_out/r2u_trace_bulk1.log
