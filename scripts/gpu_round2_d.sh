#!/bin/bash
# GPU call 5 of round 2 (1 GPU): kernel-difference bisection, Dij with immediate insertion + wide probing, variants, c3like fixture, suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/packed
timeout 300 python scripts/debug_dij_wc2.py 2>&1 | tail -14
timeout 900 python scripts/config_bench.py c4 c4big > gpurun_out/r2d_configs.log 2>&1; cut -c1-330 gpurun_out/r2d_configs.log
for v in pw1 pw2 park4; do echo "== $v"; MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 2>&1 | cut -c1-40,100-330; done
timeout 600 python oracle/gen_golden_gpu.py c3like > gpurun_out/gold4.log 2>&1; grep -E "^==|FAILED|harness failed" gpurun_out/gold4.log
python oracle/pack_golden_gpu.py > gpurun_out/r2d_pack.log 2>&1; cp tests/golden/c3like_head_release.npz gpurun_out/packed/; rm -rf gpurun_out/golden_gpu
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2d_tests.log | head
MQI_COUNT_STEPS=0 SKIP=0 bash scripts/gpu_ncu_cmd.sh r2d_c4 python scripts/c4_bench.py 80000001 1000 10000 1
bash scripts/gpu_ncu.sh r2d_c1 > /dev/null 2>&1; rm -f gpurun_out/r2d_c1.ncu-rep gpurun_out/r2d_c1.sass.csv; ls -la gpurun_out/r2d_c1.*
du -sh gpurun_out
