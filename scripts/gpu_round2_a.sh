#!/bin/bash
# GPU call 2 of round 2: tests on the new kernels, config benches, ncu captures, then the 1e9-history C1 goldens
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2a_tests.log
timeout 600 python scripts/config_bench.py c2 c3 c4 c4big rs rs1 rs0 > gpurun_out/r2a_configs.log 2>&1; cat gpurun_out/r2a_configs.log | cut -c1-400
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.json
SKIP=0 bash scripts/gpu_ncu_cmd.sh r2a_c4 python scripts/c4_bench.py 80000001 1000 10000 1
SKIP=1 bash scripts/gpu_ncu_cmd.sh r2a_rs python scripts/multi_bench.py 4000000 2
python scripts/make_c3_case.py /tmp/c3 1 && SKIP=0 bash scripts/gpu_ncu_cmd.sh r2a_c3 moquimc_b200/bin/tps_env /tmp/c3/c3.in
SKIP=1 bash scripts/gpu_ncu_cmd.sh r2a_c2 python scripts/config_bench.py c2
rm -f gpurun_out/*.ncu-rep.tmp
timeout 1500 python oracle/gen_golden_gpu.py c1 --budget-s 700 --c1-histories 1e9 --c1-histories-release 6e8 > gpurun_out/gold2.log 2>&1; tail -5 gpurun_out/gold2.log | cut -c1-600
ls -la gpurun_out/golden_gpu; du -sh gpurun_out
