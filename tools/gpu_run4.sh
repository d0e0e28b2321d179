#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 22 gpurun_out/$name.log; }
run t4_small    400 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_vocoder.py tests/test_gpu_api.py -q -m gpu -k "not full and not max_length" --no-header -x
run t4_full     700 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_gpt.py -q -m gpu -k "tc_full or max_length or bf16_full or (golden and full)" --no-header -s
run t4_perf     300 python tools/perf_probe.py 1 64 48
run t4_bench    600 python bench.py --gpus 1 --steps 1 --warmup 1 --requests 8
