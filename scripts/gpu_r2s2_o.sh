#!/bin/bash
# what the correctly rounded voxel-exit divisions cost: C1 with rcp.approx * n instead (measurement only), and the same-stream parity test on it
mkdir -p gpurun_out
for v in new apxdiv new apxdiv; do
  echo "== $v"; MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so python scripts/quick_bench.py 1e7 1 | tail -1 | cut -c1-110
done
MQI_B200_LIB=$PWD/moquimc_b200/variants/libmqi_apxdiv.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
for v in new apxdiv; do echo "== $v c3"; MQI_B200_LIB=$PWD/moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c2 2>&1 | tail -1 | cut -c1-40,170-260; done
