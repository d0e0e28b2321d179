#!/bin/bash
# round 2, run 13: decode attention with 1 / 2 / 4 warps per item (interleaved repeats); bench with the tuned host allocator
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
run r2m_attn 300 python -m pytest tests/test_gpu_gpt.py -q --no-header -k "capped_attention or microbatch or full_size_decode"
TAIL=12 run r2m_probe 900 python tools/stagger_probe.py 163
timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --no-extras > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench exit $?"; tail -n 6 gpurun_out/r2m_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1]);print({k:d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))"
