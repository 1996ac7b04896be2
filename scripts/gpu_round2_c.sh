#!/bin/bash
# GPU call 4 of round 2: Dij write-combining debug, fixtures regenerated with independent sampler streams, full suite, benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/packed
timeout 120 python scripts/debug_dij_wc.py 2>&1 | tail -12
timeout 900 python scripts/config_bench.py c4 c4big rs > gpurun_out/r2c_configs.log 2>&1; cut -c1-330 gpurun_out/r2c_configs.log
for v in m640 fp4 adv10; do echo "== $v"; MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py $( [ $v = fp4 ] && echo c4 || echo rs ) 2>&1 | cut -c1-40,120-330; done
timeout 2400 python oracle/gen_golden_gpu.py c2sweep dose2 c3like c1 --budget-s 700 --c1-histories 1e9 --c1-histories-release 3e8 > gpurun_out/gold3.log 2>&1; grep -E "^==|FAILED" gpurun_out/gold3.log
python oracle/pack_golden_gpu.py > gpurun_out/r2c_pack.log 2>&1; tail -8 gpurun_out/r2c_pack.log
cp tests/golden/c1_water200_*_3d.npz tests/golden/c2_sweep_release.npz tests/golden/c2_slabs150_dose2.npz tests/golden/c3like_head_release.npz gpurun_out/packed/
rm -rf gpurun_out/golden_gpu
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|C1 |per-voxel|depth slabs|C2 sweep|^  [ 0-9]{3}  " gpurun_out/r2c_tests.log | head -70
du -sh gpurun_out
