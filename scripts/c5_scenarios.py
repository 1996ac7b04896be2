"""Config C5 on all GPUs of the box: the 21 robust-evaluation scenarios of the C3 plan (nominal and +-3 mm setup shifts
on one axis at a time, each at density scaling 1, 0.965 and 1.035) as independent tps_env runs, round-robin over the
devices (replicas only: no collective).  Prints one JSON line: wall time of the whole evaluation, per-scenario kernel
rates, and the dose-centroid shifts that show every scenario did what its input says.
    python scripts/c5_scenarios.py [n_gpus] [ParticlesPerHistory]"""
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import capi, configs as K, parallel as P, synthetic as S

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else capi.device_count()
pph = float(sys.argv[2]) if len(sys.argv) > 2 else 2000.0
root = tempfile.mkdtemp(prefix="mqi_c5_")
K.c3_case(root)
scen = P.robust_scenarios()
res = [None] * len(scen)


def worker(gpu):
    for i in P.scenario_shard(len(scen), gpu, n_gpus):
        od = os.path.join(root, "o_%02d" % i)
        inp = os.path.join(root, "s_%02d.in" % i)
        S.write_input(inp, root, od, ParticlesPerHistory=pph, GPUID=gpu, **scen[i])
        t0 = time.time()
        out = K.run_tps(inp)
        r = K.parse_tps(out)
        d = np.fromfile(os.path.join(od, "G000_0_Dose.raw"), dtype=np.float64).reshape(200, 512, 512)
        w = d.sum(axis=(0, 1)), d.sum(axis=(0, 2)), d.sum(axis=(1, 2))
        cen = [float((np.arange(a.size) * a).sum() / a.sum()) for a in w]
        res[i] = dict(scen[i], gpu=gpu, wall_s=time.time() - t0, kernel_ms=r["kernel_ms"], rate=r["value"], histories=r["histories"],
                      centroid_xyz=cen, dose_sum=float(d.sum()))
        os.remove(os.path.join(od, "G000_0_Dose.raw"))


t0 = time.time()
th = [threading.Thread(target=worker, args=(g,)) for g in range(n_gpus)]
[t.start() for t in th]
[t.join() for t in th]
wall = time.time() - t0
nom = res[0]["centroid_xyz"]
for r in res:
    r["centroid_shift_voxels"] = [a - b for a, b in zip(r["centroid_xyz"], nom)]
hist = sum(r["histories"] for r in res)
print(json.dumps({"config": "C5: 21 scenarios of the C3 plan (512x512x200 CT, 2025 spots, Dose), ParticlesPerHistory %g" % pph,
                  "n_gpus": n_gpus, "wall_s": wall, "histories": hist, "histories_per_s_wall": hist / wall,
                  "kernel_rate_per_gpu_median": float(np.median([r["rate"] for r in res])), "scenarios": res}))
