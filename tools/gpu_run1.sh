#!/bin/bash
# first GPU session: correctness of every kernel family, each group under its own timeout
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 25 gpurun_out/$name.log; }
run t_gemm_f32   300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm_f32 or sampler" -x --no-header -s
run t_gemm_tc    300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tcgen05" --no-header -s
run t_voc_small  300 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -k "small" --no-header -s
run t_gpt_small  400 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "small and not bf16" --no-header -s
run t_gpt_bf16   400 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "bf16" --no-header -s
run t_full       900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_gpt.py -q -m gpu -k "full" --no-header -s
run t_smoke      300 python __graft_entry__.py smoke
run t_perf       600 python tools/perf_probe.py 0,1 32 48
