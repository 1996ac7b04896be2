#!/bin/bash
# round 2, session 2, call F (1 GPU): the tree with the reversed fetch order -- GPU suite, headline kernel re-captured
# (ncu --set full, steady-state launch) for profiles/roofline_traffic.json, C1 A/B of the fetch order, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s2f_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2s2f_tests.log | head
bash scripts/gpu_ncu.sh r2s2f_c1 > /dev/null 2>&1
python scripts/make_roofline_traffic.py gpurun_out/r2s2f_c1.raw.csv 1e7 "ncu --set full --clock-control none --import-source on, scripts/gpu_ncu.sh r2s2f_c1: scripts/quick_bench.py 1e7 1, third launch (round 2, session 2 tree)" > gpurun_out/r2s2f_traffic.log 2>&1
cp profiles/roofline_traffic.json gpurun_out/r2s2f_roofline_traffic.json
tail -2 gpurun_out/r2s2f_traffic.log | cut -c1-600
rm -f gpurun_out/r2s2f_c1.ncu-rep
timeout 900 python bench.py > gpurun_out/r2s2f_bench.json 2> gpurun_out/r2s2f_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2s2f_bench.json; tail -3 gpurun_out/r2s2f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s2f_bench.json").read().strip().splitlines()[-1])
print("e2e", d["e2e"]["value"], d["e2e"]["serial_value"])
print("roofline", json.dumps(d["roofline"])[:700])
for k, v in d.get("configs", {}).items():
    print(k, v.get("value"), v.get("kernel_ms"))
print("strong", json.dumps(d.get("strong"))[:500])
PY
