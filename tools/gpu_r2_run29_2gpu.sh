#!/bin/bash
# round 2, run 29 (2 GPUs): the torchrun bench line on the final build (fp16 default), both arms as the driver launches them
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2ac_bench_2gpu.json 2> gpurun_out/r2ac_bench_2gpu.err; echo "bench exit $?"; tail -n 6 gpurun_out/r2ac_bench_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/r2ac_bench_2gpu.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','dtype')}, 'e2e', d['e2e']['value'], 'ragged', d.get('ragged',{}).get('value'), 'strong', (d.get('strong') or {}).get('value'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2ac_ref_2gpu.json 2> gpurun_out/r2ac_ref_2gpu.err; echo "ref exit $?"; tail -n 2 gpurun_out/r2ac_ref_2gpu.json | cut -c1-300
