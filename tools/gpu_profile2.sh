#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv1d_tc|conv_post|interp|atoms_zero' --csv \
    --log-file gpurun_out/voc_launches.csv python tools/profile_vocoder_batch.py 605 > gpurun_out/voc_launches.log 2>&1
echo "voc launch list exit $?"; wc -l gpurun_out/voc_launches.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_kernels.py 32 6 605 > gpurun_out/launches.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 30 -c 4 -o gpurun_out/prof_conv1d_tc \
    python tools/profile_vocoder_batch.py 605 > gpurun_out/prof_conv.log 2>&1
echo "conv capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 3000 -c 2 -o gpurun_out/prof_attn_decode \
    python tools/profile_kernels.py 160 150 8 > gpurun_out/prof_attn.log 2>&1
echo "attn capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 1000 -c 4 -o gpurun_out/prof_gemm_tc \
    python tools/profile_kernels.py 160 12 8 > gpurun_out/prof_gemm.log 2>&1
echo "gemm capture exit $?"
timeout 300 python -m pytest tests -q -m gpu --durations=12 -k "full" --no-header > gpurun_out/full_durations.log 2>&1; tail -n 20 gpurun_out/full_durations.log
echo "=== cfg3 streaming bench"
timeout 600 python tools/bench_stream.py 256 256 > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; echo "exit $?"; tail -n 3 gpurun_out/bench_stream.err; cat gpurun_out/bench_stream.json
timeout 300 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_api.py -q -m gpu -k "small or api or speech or streaming or prepared" --no-header -x > gpurun_out/t_small_after_priority.log 2>&1; tail -n 3 gpurun_out/t_small_after_priority.log
