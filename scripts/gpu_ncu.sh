#!/bin/bash
# ncu --set full capture of one steady-state transport launch (1e7 histories, debug physics) + CSV pages
# usage: bash scripts/gpu_ncu.sh <tag> [variant-lib]
tag=$1; lib=${2:-moquimc_b200/libmqi_b200.so}
MQI_B200_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel --launch-skip 2 -c 1 \
  -f -o gpurun_out/$tag python scripts/quick_bench.py 1e7 1 > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/$tag.sass.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$tag.src.csv 2>/dev/null
ls -la gpurun_out/$tag.*
