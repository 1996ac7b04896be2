#!/bin/bash
# round 2, session 2, call B: ncu --set full of the current SET_DIJ kernel on the C4 geometry (1 000 spots, load 0.73),
# CSV pages only; plus the Dij kernel at 896 / 1024 threads per CTA
mkdir -p gpurun_out
export MQI_COUNT_STEPS=0
bash scripts/gpu_ncu_c4.sh r2s2b_c4
rm -f gpurun_out/r2s2b_c4.ncu-rep
for v in cur896 cur1024; do
  echo "== $v"
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 2>&1 | tail -1 | cut -c1-400
done
