#!/bin/bash
# Final single-GPU evidence of round 2: full GPU suite, the bench line, the ncu launch list of the same command, L2 window experiment
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2f_tests.log | head
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2f_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_reference_arm.err; cut -c1-300 gpurun_out/r2f_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2f_launches_bench.csv
for m in 0 1 2; do echo "== l2_persist $m"; MQI_L2_PERSIST=$m python scripts/quick_bench.py 1e7 1 | tail -1; done
python __graft_entry__.py smoke 2>&1 | tail -1
du -sh gpurun_out
