#!/bin/bash
# run 12: chain kernel + lean conv epilogue — suite, vocoder probe, bench with and without the chain on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 12 gpurun_out/$name.log; }
run t12_suite   600 python -m pytest tests -q -m gpu --no-header -x
if grep -q "failed\|rror" gpurun_out/t12_suite.log; then echo "suite failed: skipping bench"; exit 0; fi
run t12_vprobe  400 python tools/voc_probe.py 605
echo "=== bench chain=1"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "exit $?"; tail -n 6 gpurun_out/bench12.err; python -c "
import json;d=json.loads(open('gpurun_out/bench12.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'], d['clocks']);print({k:(v['ms'],v['launches']) for k,v in d['roofline']['families'].items()})"
echo "=== bench chain=0"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --decode-chain 0 > gpurun_out/bench12_nochain.json 2> gpurun_out/bench12_nochain.err; echo "exit $?"; tail -n 4 gpurun_out/bench12_nochain.err; python -c "
import json;d=json.loads(open('gpurun_out/bench12_nochain.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'])"
