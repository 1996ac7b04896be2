"""Config C4 with the spots sharded over the GPUs of the box (one process, one engine per device, no reduction: the rows
of the Dij matrix are disjoint): python scripts/c4_multi.py [n_gpus] [slots per device]
Prints one JSON line: kernel time (max over devices), aggregate histories/s, entries per device; the first device's rows
are compared with a single-device run of the same spots."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import capi, configs as K, parallel as P

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else capi.device_count()
cap = int(float(sys.argv[2])) if len(sys.argv) > 2 else 393_216_001
n_spots, per = 5000, 10_000
eng = []
for d in range(n_gpus):
    e, s, total = K.c4_setup(d, cap, n_spots, per)
    eng.append((e, s))
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 4
shards = [P.spot_shard_blocks([per] * n_spots, d, n_gpus, blocks) for d in range(n_gpus)]
for (e, s), sh in zip(eng, shards):      # warm-up
    e.run(77, sh[0][0], min(sh[0][1], 100_000), per_spot=True)
    e.clear_scorers()
t0 = time.time()
dev_ms = [0.0] * n_gpus
full = [0] * n_gpus
for k in range(max(len(sh) for sh in shards)):      # one block per device at a time, the devices run concurrently
    live = [d for d in range(n_gpus) if k < len(shards[d])]
    for d in live:
        eng[d][0].run_async(77, shards[d][k][0], shards[d][k][1], per_spot=True)
    for d in live:
        st = eng[d][0].run_stats()
        dev_ms[d] += st.kernel_ms
        full[d] += int(st.dij_table_full)
wall = time.time() - t0
kms = max(dev_ms)
nnz = [e.get_sparse_count(s) for e, s in eng]
# rows of the first 8 spots of device 0 against a single-device run of exactly those spots
k1, k2, v = eng[0][0].get_sparse(eng[0][1])
sel = k2 < 8
a = {(int(x), int(y)): z for x, y, z in zip(k1[sel], k2[sel], v[sel])}
e1, s1, _ = K.c4_setup(0, 8_000_001, n_spots, per)
e1.run(77, 0, 8 * per, per_spot=True)
b1, b2, bv = e1.get_sparse(s1)
b = {(int(x), int(y)): z for x, y, z in zip(b1, b2, bv)}
same = a.keys() == b.keys() and bool(np.allclose([a[k] for k in sorted(a)], [b[k] for k in sorted(a)], rtol=1e-9))
print(json.dumps({"config": "C4: 5000 spots x 10000 histories, spots sharded over %d GPU(s), %d slots per device" % (n_gpus, cap | 1),
                  "n_gpus": n_gpus, "kernel_ms_max": kms, "histories": n_spots * per, "histories_per_s": n_spots * per / (kms * 1e-3),
                  "wall_s": wall, "nnz_per_device": nnz, "kernel_ms_per_device": dev_ms, "blocks_per_device": blocks,
                  "table_full": full, "rows_equal_single_device": same}))
