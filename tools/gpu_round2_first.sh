#!/bin/bash
# First GPU call of the next round (~6 min on one B200): validate what round 1 could only write after its GPU budget was
# spent, then measure it.  Outputs under gpurun_out/; copy what is to be judged into profiles/ as r02_*.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 240 python -m pytest tests -q -m gpu --no-header -x > gpurun_out/r2_suite.log 2>&1; echo "suite exit $?"; tail -n 3 gpurun_out/r2_suite.log
echo "=== experimental: first audio early (xtts_sampling.early_tokens)"
XTTS_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_zz_gpu_early_audio.py -q --no-header > gpurun_out/r2_early.log 2>&1; echo "early exit $?"; tail -n 6 gpurun_out/r2_early.log
echo "=== bench default (e2e after chunk-by-chunk assembly + speaker host cache)"
timeout 240 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "exit $?"; tail -n 4 gpurun_out/r2_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1]);print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d.get('cpu_baseline',{}).get('value'))"
echo "=== cfg3 stream, 64 requests: time to first audio without / with early emit (58 tokens)"
timeout 150 python tools/bench_stream.py 64 256 0  > gpurun_out/r2_stream_plain.json 2> gpurun_out/r2_stream_plain.err; tail -c 600 gpurun_out/r2_stream_plain.json
timeout 150 python tools/bench_stream.py 64 256 58 > gpurun_out/r2_stream_early.json 2> gpurun_out/r2_stream_early.err; tail -c 600 gpurun_out/r2_stream_early.json
