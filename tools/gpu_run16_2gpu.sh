#!/bin/bash
# run 16: 2-GPU torchrun check of the current build (data-parallel bench arm + cfg4 book shards)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench16_2gpu.json 2> gpurun_out/bench16_2gpu.err
echo "exit $?"; tail -n 8 gpurun_out/bench16_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/bench16_2gpu.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpt_tokens_per_s','gpu_launches','scaling')}, d['e2e'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    tools/bench_book.py 500000 5000 > gpurun_out/bench16_book_2gpu.json 2> gpurun_out/bench16_book_2gpu.err
echo "book exit $?"; tail -n 2 gpurun_out/bench16_book_2gpu.err; cat gpurun_out/bench16_book_2gpu.json
