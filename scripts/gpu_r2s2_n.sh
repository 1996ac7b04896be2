#!/bin/bash
# launch list of the bench command on the final tree (kernel share of a step; numbers under ncu are never bench values)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s2n_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-configs --no-strong --no-gpu-baseline --no-cpu-baseline > gpurun_out/r2s2n_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2s2n_launches.csv
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2s2n_launches.csv')) if len(r) > 14 and r[0].isdigit()]
acc = collections.OrderedDict()
for r in rows:
    k = r[4][:70]; acc.setdefault(k, [0, 0.0]); acc[k][0] += 1; acc[k][1] += float(r[14]) / 1e6
for k, (n, ms) in acc.items(): print("%4d launches %10.3f ms  %s" % (n, ms, k))
PY
