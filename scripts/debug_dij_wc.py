"""Debug aid: key sets of a write-combined and a directly inserted Dij scorer in one run, against a run without write-combining."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from moquimc_b200 import capi
from test_gpu_edge_cases import engine

def run(wc, kinds=(capi.SCORER_DIJ, capi.SCORER_DOSE, capi.SCORER_DIJ)):
    e, ids = engine(kinds=kinds, capacity=1_000_003)
    e.set_option("dij_write_combine", wc)
    bl = [capi.make_beamlet(90.0 + 15.0 * i, [-12.0 + 8.0 * i, 3.0, 0.5, 0, 0, -1], [3, 3, 0, 0.002, 0.002, 0], uniform=False) for i in range(4)]
    e.set_beamlets(bl, [3000] * 4)
    st = e.run(seed=5, first=0, count=12000, per_spot=True)
    out = []
    for s, k in zip(ids, kinds):
        if k != capi.SCORER_DIJ:
            continue
        k1, k2, v = e.get_sparse(s)
        out.append({(int(a), int(b)): c for a, b, c in zip(k1, k2, v)})
    return out, st

(a0, a2), st = run(1)
(b0, b2), _ = run(0)
print("table_full", st.dij_table_full, "sizes", len(a0), len(a2), len(b0), len(b2))
for name, x in (("a0", a0), ("a2", a2), ("b2", b2)):
    only_x, only_b = set(x) - set(b0), set(b0) - set(x)
    print(name, "only in it", len(only_x), sorted(only_x)[:8], "missing", len(only_b), sorted(only_b)[:8])
    common = sorted(set(x) & set(b0))
    d = np.array([x[k] for k in common]) / np.array([b0[k] for k in common]) - 1
    print("   value rel diff max", np.abs(d).max(), "sum ratio", sum(x.values()) / sum(b0.values()))
(c0,), st = run(1, kinds=(capi.SCORER_DIJ,))
print("single Dij (SET_DIJ) sizes", len(c0), "only", len(set(c0) - set(b0)), "missing", len(set(b0) - set(c0)), "sum ratio", sum(c0.values()) / sum(b0.values()))
