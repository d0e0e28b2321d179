#!/bin/bash
# run 19 (6 GPU-minutes left in the round): validate the session's host-side changes on the box — GPU suite, default
# bench (CUDA-event timing, per-replay attention bytes, roofline.traffic), then two option sweeps if time remains.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu --no-header -x > gpurun_out/t19_suite.log 2>&1; echo "suite exit $?"; tail -n 4 gpurun_out/t19_suite.log
echo "=== bench default"
timeout 200 python bench.py > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo "exit $?"; tail -n 5 gpurun_out/bench19.err; python -c "
import json;d=json.loads(open('gpurun_out/bench19.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'], d['clocks']); r=d['roofline']; print(r['kernel'], r['frac'], r['traffic']); print({k:(v['ms'],v['launches'],v['frac_of_hbm_peak']) for k,v in r['families'].items()})"
for v in "4 128" "4 64" "3 128"; do set -- $v
  echo "=== sweep microbatches=$1 gemm_bn=$2"
  timeout 90 python bench.py --sweep --steps 1 --warmup 3 --microbatches $1 --engine-opt gemm_bn=$2 > gpurun_out/bench19_mb$1_bn$2.json 2> gpurun_out/bench19_mb$1_bn$2.err
  python -c "
import json,sys;d=json.loads(open('gpurun_out/bench19_mb$1_bn$2.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'])" || tail -n 3 gpurun_out/bench19_mb$1_bn$2.err
done
