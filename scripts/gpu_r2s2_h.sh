#!/bin/bash
# round 2, session 2, call H: first probe step one slot, later steps 2 ... 8 slots, on C4 (both table sizes)
mkdir -p gpurun_out
O=gpurun_out/r2s2h.log
: > $O
for v in pw1_2 pw1_3 pw1_4 pw1_5 pw1_6 pw1_8; do
  echo "== $v" >> $O
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 2>&1 | tail -1 >> $O
done
for v in pw1_4 pw1_6; do
  echo "== $v c4big" >> $O
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4big 2>&1 | tail -1 >> $O
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2s2h.log'):
    if ln.startswith('=='): print(ln.strip(), end='  ')
    else:
        try:
            n, j = ln.split(' ', 1); d = json.loads(j); print("%s %.4g (%.1f ms) nnz %d full %d" % (n, d['value'], d['kernel_ms'], d['nnz'], d['table_full']))
        except Exception: print(ln.strip()[:300])
PY
