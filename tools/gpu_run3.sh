#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 25 gpurun_out/$name.log; }
run t3_voc_tc   300 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -k "tc_small or (small and oracle)" --no-header -s
run t3_gpt      300 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_api.py -q -m gpu -k "small or api or speech or streaming or prepared" --no-header -s
run t3_full     600 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_gpt.py -q -m gpu -k "tc_full or max_length or bf16_full" --no-header -s
run t3_perf     300 python tools/perf_probe.py 1 64 48
run t3_bench    600 python bench.py --gpus 1 --steps 1 --warmup 1 --requests 8
