#!/bin/bash
# run 10: conv epilogue warpgroups A/B, ncu source-level capture of stage-3 / stage-1 convs at batch 8, suite, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 16 gpurun_out/$name.log; }
run t10_voc     300 python -m pytest tests/test_gpu_vocoder.py -q -m gpu --no-header
run t10_vprobe  400 python tools/voc_probe.py 605
run t10_suite   600 python -m pytest tests -q -m gpu --no-header
echo "=== ncu stage-3 convs (eg=2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 59 -c 2 -o gpurun_out/prof10_conv_s3 -f \
    python tools/profile_vocoder_batch.py 605 2 > gpurun_out/prof10_s3.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/prof10_s3.log
echo "=== ncu stage-1 convs (eg=2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 21 -c 2 -o gpurun_out/prof10_conv_s1 -f \
    python tools/profile_vocoder_batch.py 605 2 > gpurun_out/prof10_s1.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/prof10_s1.log
echo "=== bench full"
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "exit $?"; tail -n 6 gpurun_out/bench10.err; python -c "
import json;d=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpt_tokens_per_s','gpu_launches')}, d['e2e'], d['clocks']);print({k:(v['ms'],v['launches']) for k,v in d['roofline']['families'].items()})"
