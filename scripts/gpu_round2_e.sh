#!/bin/bash
# GPU call (1 GPU): full suite on the final Dij code, config benches, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2e_tests.log | head
timeout 600 python scripts/config_bench.py c4 c4big rs rs1 c2 c3 > gpurun_out/r2e_configs.log 2>&1; cut -c1-60,100-330 gpurun_out/r2e_configs.log
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "stack_overflows")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic_source"])
    print("cpu", d.get("cpu_baseline")); print("gpu ref", {k: d["gpu_reference_baseline"].get(k) for k in ("value", "threads", "kernel_s", "this_over_reference_cuda", "note")})
    print("strong", {k: d["strong"].get(k) for k in ("time_to_criterion_s", "passes", "histories", "transport_s", "stat_s", "uncertainty_percent", "note")})
    for k, v in d["configs"].items():
        print(k, {a: v.get(a) for a in ("value", "kernel_ms", "note")}, (v.get("roofline") or {}).get("frac"))
except Exception as ex:
    print("parse failed", ex); print(open("gpurun_out/r2e_bench.err").read()[-2000:])
PY
du -sh gpurun_out
