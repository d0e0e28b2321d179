#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== full-config bench (32 requests x 1k chars)"
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; tail -n 12 gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json
echo "=== reference arm"
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; tail -c 1200 gpurun_out/bench_ref.json
bash tools/gpu_profile.sh
