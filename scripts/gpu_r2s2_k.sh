#!/bin/bash
# ncu --set full of the range shifter + aperture world with the hand-over queue (transport_kernel<release, SET_DOSE, MULTI>)
mkdir -p gpurun_out
tag=r2s2k_rs
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:transport_kernelILi0ELi1ELb1ELb0 --launch-skip 1 -c 1 \
  -f -o gpurun_out/$tag python scripts/config_bench.py rs > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$tag.src.csv 2>/dev/null
rm -f gpurun_out/$tag.ncu-rep
tail -3 gpurun_out/$tag.log | cut -c1-300
ls -la gpurun_out/$tag.*
