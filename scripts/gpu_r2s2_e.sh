#!/bin/bash
# round 2, session 2, call E: the CT configurations (index volume + dose grids larger than L2) with a persisting L2 window over the
# material volume (option l2_persist = 1 through MQI_L2_PERSIST) and with an L1 prefetch of the next voxel's material index
mkdir -p gpurun_out
O=gpurun_out/r2s2e.log
: > $O
for v in new pfn2; do for l in 0 1; do
  echo "== $v l2_persist=$l" >> $O
  MQI_L2_PERSIST=$l MQI_B200_LIB=$PWD/moquimc_b200/variants/libmqi_$v.so timeout 600 python scripts/config_bench.py c3 c4 c4big 2>&1 | tail -3 >> $O
done; done
echo "== pfn1 C1" >> $O
MQI_B200_LIB=moquimc_b200/variants/libmqi_pfn1.so python scripts/quick_bench.py 1e7 1 | tail -1 >> $O
python - <<'PY'
import json
for ln in open('gpurun_out/r2s2e.log'):
    if ln.startswith('=='): print(ln.strip())
    else:
        try:
            n, j = ln.split(' ', 1); d = json.loads(j); print("   %s %.4g (%.1f ms)" % (n, d['value'], d['kernel_ms']))
        except Exception: print("   ", ln.strip()[:200])
PY
