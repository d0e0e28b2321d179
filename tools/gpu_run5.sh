#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 22 gpurun_out/$name.log; }
run t5_voc      300 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -k "tc_full or max_length or tc_small" --no-header -s
run t5_small    300 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_kernels.py -q -m gpu -k "small or tcgen05 or gemm" --no-header -x
run t5_step     400 python tools/step_probe.py 41 96
run t5_bench    600 python bench.py --gpus 1 --steps 1 --warmup 1 --requests 8
run t5_full     600 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "bf16_full" --no-header -s
