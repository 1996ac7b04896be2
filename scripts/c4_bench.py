"""C4 (Dij, 5 000 spots x 1e4 histories, 256 x 256 x 150 CT) throughput probe: python scripts/c4_bench.py [capacity] [spots] [per] [write_combine]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moquimc_b200 import capi, synthetic as S

cap = int(float(sys.argv[1])) if len(sys.argv) > 1 else 393_216_001
n_spots = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
per = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
n, sp = (256, 256, 150), (1.5, 1.5, 2.0)
hu, origin = S.head_ct(n, sp, seed=4)
xe = (np.float32(origin[0] - sp[0] / 2) + np.arange(n[0] + 1, dtype=np.float32) * np.float32(sp[0])).astype(np.float32)
ye = (np.float32(origin[1] - sp[1] / 2) + np.arange(n[1] + 1, dtype=np.float32) * np.float32(sp[1])).astype(np.float32)
ze = (np.float32(origin[2] - sp[2] / 2) + np.arange(n[2] + 1, dtype=np.float32) * np.float32(sp[2])).astype(np.float32)
rng = np.random.default_rng(5)
g = np.arange(-30.0, 30.0 + 1e-6, 4.0)
pos = [(x, z) for x in g for z in g if x * x + z * z <= 30.0 ** 2 + 1e-6]
R = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32)
bl = []
for e_mev in np.linspace(80.0, 160.0, 20):
    for i in rng.choice(len(pos), size=n_spots // 20, replace=True):
        x, z = pos[i]
        bl.append(capi.make_beamlet(float(e_mev), [x, z, 250.0, 0, 0, -1], [3.0, 3.0, 0.0, 0.003, 0.003, 0.0], uniform=False,
                                    sigma_energy=0.6, rot=R))
e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
e.set_grid_hu(xe, ye, ze, hu)
s = e.add_scorer(capi.SCORER_DIJ, "Dij", capacity=cap | 1)
e.set_beamlets(bl, [per] * len(bl))
e.set_option("count_steps", int(os.environ.get("MQI_COUNT_STEPS", "1")))   # 0: the SET_DIJ kernel (the counter runs the general one)
wc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
e.set_option("dij_write_combine", wc)
for rep in range(2):
    e.clear_scorers()
    st = e.run(seed=77, first=0, count=len(bl) * per, per_spot=True)
    nnz = e.get_sparse_count(s) if hasattr(e, "get_sparse_count") else -1
    print("write_combine %d nnz %d " % (wc, nnz), end="")
    print("capacity %d: %d histories kernel %.1f ms -> %.3e hist/s, steps/hist %.1f, table full %d"
          % (cap, st.histories, st.kernel_ms, st.histories / (st.kernel_ms * 1e-3), st.steps / st.histories, st.dij_table_full), flush=True)
