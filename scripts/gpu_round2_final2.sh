#!/bin/bash
# last single-GPU verification of round 2: full GPU suite and the default bench line on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2g_tests.log | head
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2g_bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print(json.dumps(d["strong"])[:900])
PY
