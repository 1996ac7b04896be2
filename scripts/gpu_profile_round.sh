#!/bin/bash
# Round evidence in one GPU call: bench line, ncu launch list of the same command, ncu --set full of one steady-state launch.
# usage: bash scripts/gpu_profile_round.sh <tag>
tag=$1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
bash scripts/gpu_ncu.sh ${tag}_full
