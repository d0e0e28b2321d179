#!/bin/bash
# run 11: fused decode chain kernel — parity tests, bench with and without it on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 16 gpurun_out/$name.log; }
run t11_chain   300 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "chain or bf16_full or microbatch" --no-header -s -x
if grep -q "failed\|error" gpurun_out/t11_chain.log; then echo "chain tests failed: skipping bench"; exit 0; fi
run t11_suite   600 python -m pytest tests -q -m gpu --no-header -x
echo "=== bench chain=1"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "exit $?"; tail -n 6 gpurun_out/bench11.err; python -c "
import json;d=json.loads(open('gpurun_out/bench11.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'], d['clocks']);print({k:(v['ms'],v['launches']) for k,v in d['roofline']['families'].items()})"
echo "=== bench chain=0"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --decode-chain 0 > gpurun_out/bench11_nochain.json 2> gpurun_out/bench11_nochain.err; echo "exit $?"; tail -n 4 gpurun_out/bench11_nochain.err; python -c "
import json;d=json.loads(open('gpurun_out/bench11_nochain.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'])"
