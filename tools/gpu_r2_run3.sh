#!/bin/bash
# round 2, run 3: bench-regime parity tests, real timeline of the decode step, vocoder probe
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
run r2c_regime  900 python -m pytest tests/test_gpu_bench_regime.py -q --no-header -s
TAIL=40 run r2c_trace2  300 python tools/trace_step.py 163 415 2
TAIL=40 run r2c_trace1  300 python tools/trace_step.py 163 415 1
run r2c_vprobe  400 python tools/voc_probe.py 605
