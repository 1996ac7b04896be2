#!/bin/bash
# round 2, session 2, call I: exact fetch near the end of a launch (option tail_percent): shard tests, C1, C3 whole pass and 1/8 pass
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for t in 0 50 100 200 400; do
  echo "== C1 tail_percent $t"; MQI_TAIL_PERCENT=$t python scripts/quick_bench.py 1e7 1 | tail -2 | cut -c1-120
done
python scripts/tail_probe.py 0 50 100 200 400 2>&1 | tail -12
