#!/bin/bash
# carried-over probing of the Dij insert (one probe step per turn after the second) on C4; parity tests on the variant
mkdir -p gpurun_out
O=gpurun_out/r2s2p.log
: > $O
for v in new carry carry640; do
  echo "== $v" >> $O
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 c4big 2>&1 | tail -2 >> $O
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2s2p.log'):
    if ln.startswith('=='): print(ln.strip())
    else:
        try:
            n, j = ln.split(' ', 1); d = json.loads(j); print("   %s %.4g (%.1f ms) nnz %d full %d" % (n, d['value'], d['kernel_ms'], d['nnz'], d['table_full']))
        except Exception: print(ln.strip()[:300])
PY
MQI_B200_LIB=$PWD/moquimc_b200/variants/libmqi_carry.so timeout 600 python -m pytest tests -m gpu -x -q -k "dij or Dij or sparse" 2>&1 | tail -2
