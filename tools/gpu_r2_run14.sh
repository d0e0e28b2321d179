#!/bin/bash
# round 2, run 14: fp16-operand mode at bench scale (speed vs bf16 on the same box) + its window tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
run r2n_win 400 python -m pytest tests/test_gpu_windows.py -q --no-header -k "fp16"
for prec in fp16 bf16 fp16; do
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --sweep --precision $prec > gpurun_out/r2n_bench_$prec.json 2> gpurun_out/r2n_bench_$prec.err; echo "bench $prec exit $?"
python -c "
import json;d=json.loads(open('gpurun_out/r2n_bench_$prec.json').read().strip().splitlines()[-1]);print('$prec', {k:round(d.get(k),1) for k in ('value','ms_per_step','gpt_tokens_per_s')}, d['dtype'])"
done
