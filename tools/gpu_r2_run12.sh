#!/bin/bash
# round 2, run 12: fp16-operand mode tests, whole GPU suite, ncu launch lists + a full capture of the transposed-conv launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAIL:-14} gpurun_out/$name.log; }
TAIL=20 run r2l_fp16 400 python -m pytest tests/test_gpu_gpt.py -q --no-header -s -k "fp16"
run r2l_suite 900 python -m pytest tests -q -m gpu --no-header
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches.csv python tools/profile_kernels.py 163 5 605 > gpurun_out/r2l_launches.log 2>&1; echo "launch list exit $?"; wc -l gpurun_out/r2l_launches.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv1d_tc|conv_post|interp|atoms_zero|gather' --csv --log-file gpurun_out/r2l_voc_launches.csv python tools/profile_vocoder_batch.py 605 > gpurun_out/r2l_voc_launches.log 2>&1; echo "voc launch list exit $?"; wc -l gpurun_out/r2l_voc_launches.csv
# conv launch #20 (0-based) of a vocoder batch in a batch = the stage-2 transposed conv (the 1851 us outlier of round 1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 20 -c 2 -o gpurun_out/r2l_prof_convT -f python tools/profile_vocoder_batch.py 605 > gpurun_out/r2l_ncu_conv.log 2>&1; echo "conv capture exit $?"
