#!/bin/bash
# how often does the reference-side drop-in die, and where?  (debug variant, 200 000 histories as in the test; then small runs)
mkdir -p /tmp/dj; cd /tmp/dj
python - <<'PY'
import numpy as np
hu = np.zeros((200, 100, 100), dtype=np.int16); hu[140:160] = 800; hu.tofile('/tmp/dj/phantom.raw')
PY
R=$GRAFT_REPO_ROOT
fail=0
for i in $(seq 1 30); do
  mkdir -p out$i
  MALLOC_CHECK_=3 $R/oracle/_ref/ref_dropin_debug --lxyz 100 100 200 --pxyz 0 0 -100 --nxyz 100 100 200 --spot_energy 150 0 --spot_position 0 0 0.5 --spot_size 20 20 \
     --histories 200000 --phantom_path /tmp/dj/phantom.raw --output_prefix /tmp/dj/out$i --random_seed $((4321 + i)) --gpu_id 0 > log$i.txt 2>&1
  rc=$?
  if [ $rc -ne 0 ]; then fail=$((fail+1)); echo "run $i rc=$rc"; tail -3 log$i.txt; fi
done
echo "failures: $fail of 30"
