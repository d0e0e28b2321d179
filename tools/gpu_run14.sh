#!/bin/bash
# run 14: capped attention grid x micro-batches, forced GEMM tile width; suite first
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -x > gpurun_out/t14_suite.log 2>&1; echo "suite exit $?"; tail -n 5 gpurun_out/t14_suite.log
if grep -q "failed\|rror" gpurun_out/t14_suite.log; then echo "suite failed"; exit 0; fi
b() { name=$1; shift; timeout 400 python bench.py --gpus 1 --steps 1 --warmup 2 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "exit $? ($name)"; python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]);print('$name', {k:round(d[k],1) for k in ('value','ms_per_step','gpt_tokens_per_s')}, round(d['e2e']['value'],1))"; grep "device arm" gpurun_out/$name.err; }
b bench14_base
b bench14_attn4 --engine-opt attn_ctas_per_sm=4
b bench14_attn5 --engine-opt attn_ctas_per_sm=5
b bench14_attn3_mb3 --engine-opt attn_ctas_per_sm=3 --microbatches 3
b bench14_bn64 --engine-opt gemm_bn=64
b bench14_attn4_bn64 --engine-opt attn_ctas_per_sm=4 --engine-opt gemm_bn=64
