#!/bin/bash
# round 2, session 2, multi-GPU call: bash scripts/gpu_r2s2_multi.sh <n_gpus> [bench sizes...]   (gpurun --gpus N)
# the product's multi-GPU tests, then bench.py at the given sizes (weak headline, double-buffered e2e, strong C3 record)
N=${1:-2}; shift
SIZES=${@:-$N}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2s2m${N}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2s2m${N}_tests.log
for n in $SIZES; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 \
      > gpurun_out/r2s2m_bench_${n}gpu.json 2> gpurun_out/r2s2m_bench_${n}gpu.err; echo "bench $n rc=$?"; cut -c1-200 gpurun_out/r2s2m_bench_${n}gpu.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2s2m_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("  value %.4e e2e %.4e (serial %.4e) checksum %s / %s" % (d["value"], d["e2e"]["value"], d["e2e"]["serial_value"], d["config"]["dose_checksum"], d["e2e"]["last_step_dose_checksum"]))
    print("  strong:", {k: d["strong"].get(k) for k in ("time_to_criterion_s", "passes", "histories", "transport_s", "stat_s", "final_reduce_s", "overhead_share", "stat_phases_s_rank0", "uncertainty_percent", "collective", "note")})
except Exception as ex:
    print("  parse failed", ex); print(open("gpurun_out/r2s2m_bench_${n}gpu.err").read()[-2500:])
PY
done
