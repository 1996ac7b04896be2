#!/bin/bash
# ncu --set full capture of one transport launch of an arbitrary command + CSV pages:
#   SKIP=<launches to skip> bash scripts/gpu_ncu_cmd.sh <tag> <command ...>
tag=$1; shift
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transport_kernel --launch-skip ${SKIP:-1} -c 1 \
  -f -o gpurun_out/$tag "$@" > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$tag.src.csv 2>/dev/null
[ -n "$KEEP_REP" ] || rm -f gpurun_out/$tag.ncu-rep   # the CSV pages are what travels back (gpurun merges at most 64 MiB)
tail -3 gpurun_out/$tag.log
ls -la gpurun_out/$tag.*
