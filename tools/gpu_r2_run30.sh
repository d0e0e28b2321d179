#!/bin/bash
# round 2, run 30: ncu launch list of the bench command itself (per-launch durations; shares per kernel family, not absolutes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 --launch-count 12000 --csv --log-file gpurun_out/r2ad_bench_launches.csv python bench.py --gpus 1 --steps 1 --warmup 1 --no-extras --sweep > gpurun_out/r2ad_bench_under_ncu.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/r2ad_bench_under_ncu.log | cut -c1-300
wc -l gpurun_out/r2ad_bench_launches.csv
python - <<'PY'
import csv, collections
fam = collections.OrderedDict()
rows = [r for r in csv.reader(open('gpurun_out/r2ad_bench_launches.csv')) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
tot = 0.0
for r in rows:
    name = r[4]
    for key in ('gemm_bf16_2cta', 'gemm_bf16_tc', 'attn_decode', 'attn_generic', 'residual_reduce_ln', 'layernorm', 'conv1d_tc', 'conv_post', 'interp', 'sample', 'head_norms', 'build_decode_rows', 'kv_write'):
        if key in name: break
    else: key = 'other'
    d = fam.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += float(r[14]) / 1e3; tot += float(r[14]) / 1e3
print('launches', len(rows), 'kernel time', round(tot / 1e3, 2), 'ms')
for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:22s} {n:6d} launches {us / 1e3:9.3f} ms {us / n:8.1f} us/launch {100 * us / tot:5.1f} %')
PY
