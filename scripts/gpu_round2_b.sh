#!/bin/bash
# GPU call 3 of round 2: full GPU suite (new fixtures, RED inserts, batched hand-overs), config benches, variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|C1 |C2 sweep|^  [ 0-9]{3}  |Error|assert" gpurun_out/r2b_tests.log | head -60
timeout 600 python scripts/config_bench.py c4 c4big rs rs1 c2 c3 > gpurun_out/r2b_configs.log 2>&1; cut -c1-330 gpurun_out/r2b_configs.log
for v in m640 m512 adv1; do echo "== $v"; MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py rs rs1 2>&1 | cut -c1-60,130-330; done
for v in park3 park6; do echo "== $v"; MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 2>&1 | cut -c1-60,100-330; done
echo "== rspx"; MQI_B200_LIB=moquimc_b200/variants/libmqi_rspx.so python scripts/quick_bench.py 1e7 1 | tail -1; python scripts/quick_bench.py 1e7 1 | tail -1
MQI_COUNT_STEPS=0 SKIP=0 bash scripts/gpu_ncu_cmd.sh r2b_c4 python scripts/c4_bench.py 80000001 1000 10000 1
SKIP=1 bash scripts/gpu_ncu_cmd.sh r2b_rs python scripts/multi_bench.py 4000000 2
du -sh gpurun_out
