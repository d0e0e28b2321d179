#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 14 gpurun_out/$name.log; }
run t7_voc      300 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -k "tc_full or max_length or tc_small" --no-header -s
run t7_small    300 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_kernels.py tests/test_gpu_api.py -q -m gpu -k "small or tcgen05 or gemm or api or speech or streaming or prepared" --no-header -x
run t7_step     300 python tools/step_probe.py 41 96
run t7_full     400 python -m pytest tests/test_gpu_gpt.py -q -m gpu -k "bf16_full or prefill_full" --no-header -s
echo "=== bench full"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; tail -n 9 gpurun_out/bench_full.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e'], d['clocks']);print({k:(v['ms'],v['launches']) for k,v in d['roofline']['families'].items()})"
echo "=== reference arm"
timeout 400 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; tail -c 600 gpurun_out/bench_ref.json
