#!/bin/bash
# CTA size of the multi-node kernels with the hand-over queue
mkdir -p gpurun_out
: > gpurun_out/r2s2l.log
for v in m512 m576 m704; do
  echo "== $v" >> gpurun_out/r2s2l.log; MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py rs rs1 2>&1 | tail -2 >> gpurun_out/r2s2l.log
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2s2l.log'):
    if ln.startswith('=='): print(ln.strip())
    else:
        n, j = ln.split(' ', 1); d = json.loads(j); print("   %s %.4g (%.1f ms)" % (n, d['value'], d['kernel_ms']))
PY
