#!/bin/bash
# run 15: validation of the defaults (suite, smoke, bench both arms), ncu launch list of a reduced bench command,
# cfg3 streaming bench, cfg4 book on one GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header > gpurun_out/t15_suite.log 2>&1; echo "suite exit $?"; tail -n 4 gpurun_out/t15_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/t15_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/t15_smoke.log
echo "=== bench default"
timeout 700 python bench.py > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "exit $?"; tail -n 5 gpurun_out/bench15.err; python -c "
import json;d=json.loads(open('gpurun_out/bench15.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e']['value'], d['clocks']); r=d['roofline']; print({k:v for k,v in r.items() if k!='families'}); print({k:(v['ms'],v['launches'],v['frac_of_hbm_peak'],v['frac_of_tensor_peak']) for k,v in r['families'].items()}); print(d['cpu_baseline'])"
echo "=== bench reference arm"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench15_ref.json 2> gpurun_out/bench15_ref.err; echo "exit $?"; cat gpurun_out/bench15_ref.json | cut -c1-600
echo "=== ncu launch list of a reduced bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench15_launches.csv \
    python bench.py --steps 1 --warmup 1 --requests 4 --max-tokens 48 > gpurun_out/bench15_ncu.json 2> gpurun_out/bench15_ncu.err; echo "exit $?"; wc -l gpurun_out/bench15_launches.csv
echo "=== cfg3 streaming"
timeout 600 python tools/bench_stream.py 256 256 > gpurun_out/bench15_stream.json 2> gpurun_out/bench15_stream.err; echo "exit $?"; cat gpurun_out/bench15_stream.json
echo "=== cfg4 book, 1 GPU"
timeout 600 python tools/bench_book.py 500000 5000 > gpurun_out/bench15_book.json 2> gpurun_out/bench15_book.err; echo "exit $?"; tail -n 2 gpurun_out/bench15_book.err; cat gpurun_out/bench15_book.json
