#!/bin/bash
# round 2, run 10 (2 GPUs): torchrun bench with extras (async gather, strong-scaling arm) + in-process data parallelism
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-3} --warmup 2 > gpurun_out/r2j_bench_${N}gpu.json 2> gpurun_out/r2j_bench_${N}gpu.err; echo "bench exit $?"; tail -n 8 gpurun_out/r2j_bench_${N}gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/r2j_bench_${N}gpu.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'ragged', d.get('ragged',{}).get('value'), 'strong', d.get('strong'))"
timeout 900 python tools/bench_dp.py $N ${DP_MODES:-weak strong book} > gpurun_out/r2j_dp_${N}gpu.json 2> gpurun_out/r2j_dp_${N}gpu.err; echo "bench_dp exit $?"; tail -n 6 gpurun_out/r2j_dp_${N}gpu.err; cat gpurun_out/r2j_dp_${N}gpu.json
