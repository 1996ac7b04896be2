#!/bin/bash
# round 2, session 2, call C: the hand-over queue of the multi-node kernels and the warp-cooperative Dij insert:
# parity tests on the default build, then throughput of the variants
mkdir -p gpurun_out
O=gpurun_out/r2s2c
timeout 900 python -m pytest tests -m gpu -x -q -k "dij or Dij or sparse or beamline or roi or tps or multi or node or scorer" > $O.tests.log 2>&1
echo "tests rc=$?"; tail -4 $O.tests.log
: > $O.perf.log
for v in old new coop_b2 coop640 coop640_b2 coop_t4b2 coop_t4b3 coop_t16; do
  echo "== $v" >> $O.perf.log
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 2>&1 | tail -1 >> $O.perf.log
done
for v in old new; do
  echo "== $v c4big" >> $O.perf.log
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4big 2>&1 | tail -1 >> $O.perf.log
done
for v in old new advmin16 advmin24 adv768; do
  echo "== $v rs" >> $O.perf.log
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py rs rs1 2>&1 | tail -2 >> $O.perf.log
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2s2c.perf.log'):
    if ln.startswith('=='): print(ln.strip(), end='  ')
    else:
        try:
            n, j = ln.split(' ', 1); d = json.loads(j); print("%s %.4g (%.1f ms)" % (n, d['value'], d['kernel_ms']), d.get('nnz', ''), d.get('table_full', ''))
        except Exception: print(ln.strip()[:300])
PY
