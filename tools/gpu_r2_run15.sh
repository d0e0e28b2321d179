#!/bin/bash
# round 2, run 15: what the driver runs at round end — smoke(), the default bench line (all arms), the CPU reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/r2o_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench exit $?"; tail -n 12 gpurun_out/r2o_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step','dtype','fp32_value','gpu_launches')}, 'e2e', d['e2e']['value'], 'ragged', d.get('ragged',{}).get('value'), d.get('cfg3_ttfa',{}).get('early_58'), 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks']);r=d['roofline'];print({k:r[k] for k in ('bound','achieved','peak','frac','kernel','traffic')})"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2o_ref.json 2> gpurun_out/r2o_ref.err; echo "ref exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/r2o_ref.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'])"
