#!/bin/bash
# round 2, session 2, call A: Dij insert variants (prefetch at voxel entry / deferred resolve / CTA size) on C4,
# the Dij parity tests on the deferred variant, and a short bench run for the double-buffered e2e leg
mkdir -p gpurun_out
O=gpurun_out/r2s2a
: > $O.c4.log
for v in cur pf1 pf2 df dfp cur640 pf1_640 df640 dfp640 dfp2_640; do
  echo "== $v" >> $O.c4.log
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4 >> $O.c4.log 2>&1
done
for v in cur dfp640 dfp; do
  echo "== $v c4big" >> $O.c4.log
  MQI_B200_LIB=moquimc_b200/variants/libmqi_$v.so timeout 300 python scripts/config_bench.py c4big >> $O.c4.log 2>&1
done
cat $O.c4.log
MQI_B200_LIB=$PWD/moquimc_b200/variants/libmqi_dfp640.so timeout 600 python -m pytest tests -m gpu -x -q -k "dij or Dij or sparse" > $O.dijtests.log 2>&1
echo "dij tests (dfp640) rc=$?"; tail -3 $O.dijtests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-strong --no-gpu-baseline --no-cpu-baseline > $O.bench.json 2> $O.bench.err
echo "bench rc=$?"; cat $O.bench.json; tail -5 $O.bench.err
