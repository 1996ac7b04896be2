#!/bin/bash
# A/B of kernel variants (scripts/build_variants.py) on the GPU box: gpurun -- 'bash scripts/gpu_ab.sh v12 kin2 ...'
for v in "$@"; do
  echo "== $v"
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so python scripts/quick_bench.py 1e7 1 | tail -2
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so python scripts/quick_bench.py 1e7 0 | tail -1
done
