#!/bin/bash
# round 2, run 2: windowed / ragged / own-stream vocoder — new tests, whole suite, first overlap sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 14 gpurun_out/$name.log; }
run r2b_windows 420 python -m pytest tests/test_gpu_windows.py -q --no-header -x
if grep -q "failed\|rror" gpurun_out/r2b_windows.log; then echo "window tests failed: stopping"; exit 0; fi
run r2b_regime  600 python -m pytest tests/test_gpu_bench_regime.py -q --no-header -s
run r2b_suite   600 python -m pytest tests -q -m gpu --no-header --deselect tests/test_gpu_bench_regime.py --deselect tests/test_gpu_windows.py
b() { name=$1; shift; timeout 400 python bench.py --gpus 1 --steps 2 --warmup 2 --sweep "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "exit $? ($name)"; python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]);print('$name', {k:round(d[k],1) for k in ('value','ms_per_step','gpt_tokens_per_s')})"; grep "device arm" gpurun_out/$name.err; }
b r2b_bench_seg0
b r2b_bench_seg96_all   --engine-opt voc_segment=96
b r2b_bench_seg96_sm48  --engine-opt voc_segment=96 --engine-opt voc_sms=48
b r2b_bench_seg96_sm32  --engine-opt voc_segment=96 --engine-opt voc_sms=32
b r2b_bench_seg128_sm64 --engine-opt voc_segment=128 --engine-opt voc_sms=64
b r2b_bench_seg64_sm48  --engine-opt voc_segment=64 --engine-opt voc_sms=48
