#!/bin/bash
# round 2, session 2, final single-GPU verification: whole GPU suite, smoke(), default bench line, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s2z_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2s2z_tests.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2s2z_bench.json 2> gpurun_out/r2s2z_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2s2z_bench.json; tail -3 gpurun_out/r2s2z_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s2z_bench.json").read().strip().splitlines()[-1])
print("e2e", d["e2e"]["value"], d["e2e"]["serial_value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("roofline frac", d["roofline"]["frac"], d["roofline"]["traffic_source"], d["roofline"]["issue_slots"]["frac"])
for k, v in d.get("configs", {}).items():
    print(k, v.get("value"), v.get("kernel_ms"))
print("gpu ref", (d.get("gpu_reference_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
print("strong", d["strong"]["time_to_criterion_s"], d["strong"]["passes"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s2z_bench_reference.json 2> gpurun_out/r2s2z_bench_reference.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/r2s2z_bench_reference.json
