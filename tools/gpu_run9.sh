#!/bin/bash
# run 9: sampler fast path + micro-batched decode branches; whole GPU suite with durations; bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 16 gpurun_out/$name.log; }
run t9_sampler  200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "sampler" --no-header
run t9_mb       400 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "microbatch or sampler_full or bf16_full" --no-header -s
run t9_step163  400 python tools/step_probe.py 163 64
run t9_step41   300 python tools/step_probe.py 41 96
echo "=== bench full"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "exit $?"; tail -n 9 gpurun_out/bench9.err; python -c "
import json;d=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e'], d['clocks']);print({k:(v['ms'],v['launches']) for k,v in d['roofline']['families'].items()})"
echo "=== whole GPU suite"
run t9_suite   1500 python -m pytest tests -q -m gpu --durations=30 --no-header
