#!/bin/bash
# round 2, session 2, call D (1 GPU): the whole GPU suite and the default bench line on the session's tree, then the launch
# list of the same bench command under ncu (kernel share of the step; numbers printed under ncu are never bench values)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s2d_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2s2d_tests.log | head
timeout 900 python bench.py > gpurun_out/r2s2d_bench.json 2> gpurun_out/r2s2d_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2s2d_bench.json; tail -3 gpurun_out/r2s2d_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s2d_bench.json").read().strip().splitlines()[-1])
print("e2e", json.dumps(d["e2e"])[:700])
print("roofline", json.dumps(d["roofline"])[:900])
for k, v in d.get("configs", {}).items():
    print(k, v.get("value"), v.get("kernel_ms"), (v.get("roofline") or {}).get("frac"))
print("gpu ref", json.dumps(d.get("gpu_reference_baseline"))[:300])
print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
print("strong", json.dumps(d.get("strong"))[:900])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s2d_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-configs --no-strong --no-gpu-baseline --no-cpu-baseline > gpurun_out/r2s2d_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2s2d_launches.csv
