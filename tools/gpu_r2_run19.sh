#!/bin/bash
# round 2, run 19: bulk attention + L2 prefetch one item ahead — parity test, step-time probe, timeline
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu -k "bulk_copy" > gpurun_out/r2s_test.log 2>&1; echo "test exit $?"; tail -n 6 gpurun_out/r2s_test.log
timeout 900 python tools/attn_probe.py 163 2 > gpurun_out/r2s_probe.log 2>&1; echo "probe exit $?"; tail -n 20 gpurun_out/r2s_probe.log
timeout 300 python tools/trace_step.py 163 415 1 attn_bulk=1 > gpurun_out/r2s_trace_bulk1.log 2>&1; echo "trace exit $?"; tail -n 9 gpurun_out/r2s_trace_bulk1.log
