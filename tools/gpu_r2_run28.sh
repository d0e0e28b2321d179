#!/bin/bash
# round 2, run 28: what the driver runs at round end — the whole GPU suite, smoke(), the default bench line, the CPU reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ab_suite.log 2>&1; echo "suite exit $?"; tail -n 4 gpurun_out/r2ab_suite.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2ab_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r2ab_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench exit $?"; tail -n 8 gpurun_out/r2ab_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2ab_bench.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step','dtype','fp32_value','gpu_launches')}, 'e2e', d['e2e']['value'], 'ragged', d.get('ragged',{}).get('value'), d.get('cfg3_ttfa',{}).get('early_58'), 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks']);r=d['roofline'];print({k:r[k] for k in ('bound','achieved','peak','frac','kernel','traffic')})"
