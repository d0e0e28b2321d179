#!/bin/bash
# round 2, run 31 (4 GPUs): the torchrun bench line on the final build (weak scaling point between 2 and 8)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 --no-extras > gpurun_out/r2ae_bench_4gpu.json 2> gpurun_out/r2ae_bench_4gpu.err; echo "bench exit $?"; tail -n 4 gpurun_out/r2ae_bench_4gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/r2ae_bench_4gpu.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','dtype')}, 'e2e', d['e2e']['value'])"
