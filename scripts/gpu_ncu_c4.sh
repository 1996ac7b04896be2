#!/bin/bash
# ncu --set full capture of the general kernel with the Dij scorer (C4 geometry, 1 000 spots x 1e4 histories,
# table sized for the C4 load factor 0.73): bash scripts/gpu_ncu_c4.sh <tag>
tag=$1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -c 1 \
  -f -o gpurun_out/$tag python scripts/c4_bench.py 80000001 1000 10000 1 > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$tag.src.csv 2>/dev/null
tail -2 gpurun_out/$tag.log
ls -la gpurun_out/$tag.*
