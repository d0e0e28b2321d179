#!/bin/bash
# run 13: micro-batch sweep, chain v2 (separate barrier lines + backoff), ncu evidence for the new conv epilogue and the chain kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
b() { name=$1; shift; timeout 400 python bench.py --gpus 1 --steps 1 --warmup 2 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "exit $? ($name)"; python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]);print('$name', {k:round(d[k],1) for k in ('value','ms_per_step','gpt_tokens_per_s')}, round(d['e2e']['value'],1))"; grep "device arm" gpurun_out/$name.err; }
b bench13_mb3 --decode-chain 0 --microbatches 3
b bench13_mb4 --decode-chain 0 --microbatches 4
b bench13_mb2 --decode-chain 0 --microbatches 2
b bench13_chain --decode-chain 1
echo "=== vocoder launch list (batch 8)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv1d_tc|conv_post|interp|atoms_zero' --csv \
    --log-file gpurun_out/voc13_launches.csv python tools/profile_vocoder_batch.py 605 > gpurun_out/voc13_launches.log 2>&1; echo "exit $?"
echo "=== ncu stage-3 convs"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 59 -c 2 -o gpurun_out/prof13_conv_s3 -f \
    python tools/profile_vocoder_batch.py 605 2 > gpurun_out/prof13_s3.log 2>&1; echo "exit $?"
echo "=== ncu chain kernel"
XTTS_OPTS=decode_chain=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_chain -s 40 -c 1 -o gpurun_out/prof13_chain -f \
    python tools/profile_kernels.py 163 4 8 > gpurun_out/prof13_chain.log 2>&1; echo "exit $?"; tail -n 2 gpurun_out/prof13_chain.log
