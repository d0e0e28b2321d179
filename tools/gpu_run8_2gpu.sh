#!/bin/bash
# 2-GPU torchrun check of the data-parallel bench path (what the driver does at N>1)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit $?"; tail -n 14 gpurun_out/bench_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpt_tokens_per_s','gpu_launches','scaling')}, d['e2e'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
echo "ref exit $?"; tail -c 300 gpurun_out/bench_2gpu_ref.json
