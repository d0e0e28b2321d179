#!/bin/bash
# run 18: vectorised split-K reduce + LayerNorm; ncu --set full of the decode-shaped GEMMs (M = 163)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -x > gpurun_out/t18_suite.log 2>&1; echo "suite exit $?"; tail -n 4 gpurun_out/t18_suite.log
if grep -q "failed\|rror" gpurun_out/t18_suite.log; then echo "suite failed"; exit 0; fi
echo "=== bench default"
timeout 700 python bench.py > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "exit $?"; tail -n 4 gpurun_out/bench18.err; python -c "
import json;d=json.loads(open('gpurun_out/bench18.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'], d['clocks']); r=d['roofline']; print({k:(v['ms'],v['launches']) for k,v in r['families'].items()})"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_bf16_tc_kernel<\(int\)32>' -s 900 -c 4 \
    -o gpurun_out/prof18_gemm_decode -f python tools/profile_kernels.py 163 6 8 > gpurun_out/prof18_gemm.log 2>&1; echo "gemm exit $?"; tail -n 3 gpurun_out/prof18_gemm.log
