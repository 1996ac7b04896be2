#!/bin/bash
# Final single-GPU evidence of round 2: full GPU suite, the bench line, the ncu launch list of the same command, L2 window experiment
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# the ncu capture the bench line quotes (DRAM bytes, warp instructions), tied to this library's kernel by its SASS hash
bash scripts/gpu_ncu.sh r2f_c1 > /dev/null 2>&1
python scripts/make_roofline_traffic.py gpurun_out/r2f_c1.raw.csv 1e7 "ncu --set full --clock-control none --import-source on, scripts/gpu_ncu.sh r2f_c1: scripts/quick_bench.py 1e7 1, third launch (round 2, final tree)" | cut -c1-300
cp profiles/roofline_traffic.json gpurun_out/r2f_roofline_traffic.json; rm -f gpurun_out/r2f_c1.ncu-rep gpurun_out/r2f_c1.sass.csv
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2f_tests.log | head
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2f_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_reference_arm.err; cut -c1-300 gpurun_out/r2f_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2f_launches_bench.csv
for m in 0 1 2; do echo "== l2_persist $m"; MQI_L2_PERSIST=$m python scripts/quick_bench.py 1e7 1 | tail -1; done
python __graft_entry__.py smoke 2>&1 | tail -1
du -sh gpurun_out
