#!/bin/bash
# round 2, session 2, call G: adaptive probe width of the Dij insert (first step / later steps) on C4, parity tests on the candidate
mkdir -p gpurun_out
O=gpurun_out/r2s2g.log
: > $O
for v in new pw2_3 pw2_4 pw2_6 pw2_8 pw1_4 pw3_6; do
  echo "== $v" >> $O
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 2>&1 | tail -1 >> $O
done
for v in new pw2_4; do
  echo "== $v c4big" >> $O
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4big 2>&1 | tail -1 >> $O
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2s2g.log'):
    if ln.startswith('=='): print(ln.strip(), end='  ')
    else:
        try:
            n, j = ln.split(' ', 1); d = json.loads(j); print("%s %.4g (%.1f ms) nnz %d full %d" % (n, d['value'], d['kernel_ms'], d['nnz'], d['table_full']))
        except Exception: print(ln.strip()[:300])
PY
MQI_B200_LIB=$PWD/moquimc_b200/variants/libmqi_pw2_4.so timeout 600 python -m pytest tests -m gpu -x -q -k "dij or Dij or sparse" 2>&1 | tail -2
