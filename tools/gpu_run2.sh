#!/bin/bash
# second GPU session: conditioning + API tests, re-check of the scheduler change, first bench line, ncu captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 30 gpurun_out/$name.log; }
run t2_cond     600 python -m pytest tests/test_gpu_conditioning.py -q -m gpu --no-header -s
run t2_api      300 python -m pytest tests/test_gpu_api.py -q -m gpu --no-header -s -x
run t2_gpt      300 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "small" --no-header -s
run t2_perf     400 python tools/perf_probe.py 1 64 48
run t2_bench    1500 python bench.py --gpus 1 --steps 1 --warmup 1 --requests 8
tail -n 3 gpurun_out/t2_bench.log > gpurun_out/bench_small_line.json
