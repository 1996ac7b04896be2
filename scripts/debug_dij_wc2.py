"""Debug aid: find the histories whose dose differs between kernel instantiations (bisection over the history range) and
compare them with the CPU restatement on the same Philox stream."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from moquimc_b200 import capi
import oracle_lib as O
from test_gpu_edge_cases import engine, edges, NX, NY, NZ

bl = [capi.make_beamlet(90.0 + 15.0 * i, [-12.0 + 8.0 * i, 3.0, 0.5, 0, 0, -1], [3, 3, 0, 0.002, 0.002, 0], uniform=False) for i in range(4)]


def mk(kinds, wc):
    e, ids = engine(kinds=kinds, capacity=1_000_003)
    e.set_option("dij_write_combine", wc)
    e.set_beamlets(bl, [3000] * 4)
    return e, ids


K3 = (capi.SCORER_DIJ, capi.SCORER_DOSE, capi.SCORER_DIJ)
A, ia = mk(K3, 1)
B, ib = mk(K3, 0)
C, ic = mk((capi.SCORER_DOSE,), 0)


def dose(e, sid, a, n):
    e.clear_scorers()
    e.run(seed=5, first=a, count=n, per_spot=True)
    return e.get_dense(sid).copy()


def differs(a, n):
    return not np.array_equal(dose(A, ia[1], a, n), dose(B, ib[1], a, n))


print("full range differs:", differs(0, 12000), " A deterministic:", np.array_equal(dose(A, ia[1], 0, 12000), dose(A, ia[1], 0, 12000)))
bad = []
stack = [(0, 12000)]
while stack and len(bad) < 4:
    a, n = stack.pop()
    if not differs(a, n):
        continue
    if n == 1:
        bad.append(a)
        continue
    stack.append((a + n // 2, n - n // 2))
    stack.append((a, n // 2))
print("histories that differ:", bad)
xe, ye, ze = edges()
rho = np.full(NX * NY * NZ, O.hu_to_density(np.array([0]))[0], dtype=np.float32)
g, keep = O.make_grid(xe, ye, ze, rho)
ob = [O.make_beamlet(90.0 + 15.0 * i, [-12.0 + 8.0 * i, 3.0, 0.5, 0, 0, -1], [3, 3, 0, 0.002, 0.002, 0], uniform=False) for i in range(4)]
for h in bad:
    da, db, dc = dose(A, ia[1], h, 1), dose(B, ib[1], h, 1), dose(C, ic[0], h, 1)
    (od,), st = O.transport(g, O.VARIANT_RELEASE, ob, [3000] * 4, seed=5, h0=h, n=1, kinds=[O.SCORER_DOSE])
    od = od.reshape(da.shape)
    f = lambda x: (float(x.sum()), int((x > 0).sum()))
    print("history", h, "A(wc generic)", f(da), "B(generic)", f(db), "C(simple)", f(dc), "oracle", f(od))
    print("   A==B", np.array_equal(da, db), "B==C", np.array_equal(db, dc), " max|A-oracle|/max", np.abs(da - od).max() / od.max(), " max|B-oracle|/max", np.abs(db - od).max() / od.max())
    za = np.nonzero(da.sum(axis=(1, 2)))[0]; zb = np.nonzero(db.sum(axis=(1, 2)))[0]; zo = np.nonzero(od.sum(axis=(1, 2)))[0]
    print("   depth slabs touched A %d..%d B %d..%d oracle %d..%d" % (za.min(), za.max(), zb.min(), zb.max(), zo.min(), zo.max()))
