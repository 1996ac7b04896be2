#!/bin/bash
# Multi-GPU call of round 2: bash scripts/gpu_round2_multi.sh <n_gpus>   (gpurun --gpus N)
# product multi-GPU tests, bench at N (weak + strong records), C4 spot-sharded, C5 scenarios
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ -n "$FULL_SUITE" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m${N}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2m${N}_tests.log | head
  timeout 300 python scripts/config_bench.py c4 c4big rs 2>&1 | cut -c1-50,100-330
else
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tps.py -m gpu -q > gpurun_out/r2m${N}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2m${N}_tests.log
fi
for n in $( [ "$N" = 8 ] && echo "8 4 2" || echo "$N" ); do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 \
      > gpurun_out/r2m_bench_${n}gpu.json 2> gpurun_out/r2m_bench_${n}gpu.err; echo "bench $n rc=$?"; cut -c1-200 gpurun_out/r2m_bench_${n}gpu.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("  value %.4e e2e %.4e strong: %s" % (d["value"], d["e2e"]["value"], {k: d["strong"].get(k) for k in ("time_to_criterion_s", "passes", "histories", "transport_s", "stat_s", "final_reduce_s", "overhead_share", "stat_phases_s_rank0", "uncertainty_percent", "collective", "note")}))
except Exception as ex:
    print("  parse failed", ex); print(open("gpurun_out/r2m_bench_${n}gpu.err").read()[-1500:])
PY
done
timeout 600 python scripts/c4_multi.py $N > gpurun_out/r2m${N}_c4.json 2> gpurun_out/r2m${N}_c4.err; cut -c1-600 gpurun_out/r2m${N}_c4.json; tail -3 gpurun_out/r2m${N}_c4.err
timeout 900 python scripts/c5_scenarios.py $N > gpurun_out/r2m${N}_c5.json 2> gpurun_out/r2m${N}_c5.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m${N}_c5.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "wall_s", "histories", "histories_per_s_wall", "kernel_rate_per_gpu_median")})
    print([ (s["XShift"], s["YShift"], s["ZShift"], s["DensityScaling"], [round(x, 2) for x in s["centroid_shift_voxels"]]) for s in d["scenarios"]][:9])
except Exception as ex:
    print("c5 parse failed", ex); print(open("gpurun_out/r2m${N}_c5.err").read()[-1500:])
PY
du -sh gpurun_out
