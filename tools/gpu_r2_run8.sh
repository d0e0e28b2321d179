#!/bin/bash
# round 2, run 8: CTA-pair GEMM ring variants; bench with the measured-best defaults (whole new code path), short
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=40 run r2h_gemm_probe 300 python tools/gemm_probe.py
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"; tail -n 12 gpurun_out/r2h_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step','gpt_tokens_per_s','fp32_value')}, d['e2e']['value'], d.get('ragged',{}).get('value'), d.get('cfg3_ttfa'), d.get('cpu_baseline',{}).get('value'));print({k:(v['ms'],v['launches']) for k,v in d['roofline']['families'].items()})"
