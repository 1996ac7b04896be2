#!/bin/bash
# e2e leg after the host-side changes (hu_to_material small enough to run beside a transport kernel, beam-source buffers pooled)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 10 --no-configs --no-strong --no-gpu-baseline --no-cpu-baseline > gpurun_out/r2s2m_e2e.json 2> gpurun_out/r2s2m_e2e.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s2m_e2e.json").read().strip().splitlines()[-1])
print("value %.5g e2e %.5g serial %.5g ratio %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial_value"], d["e2e"]["value"] / d["value"]))
PY
timeout 600 python -m pytest tests/test_gpu_tps.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -2
