#!/bin/bash
# round 2, run 17: cp.async.bulk streaming-rate microbenchmark + bulk attention parity test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 tools/probes/bulk_probe > gpurun_out/r2q_bulk_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/r2q_bulk_probe.log
timeout 600 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu -k "bulk_copy" > gpurun_out/r2q_test.log 2>&1; echo "test exit $?"; tail -n 6 gpurun_out/r2q_test.log
