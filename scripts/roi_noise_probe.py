"""Is the xy-projection deviation of tests/test_gpu_beamline_roi.py::test_gpu_roi_against_reference_golden noise or bias?
Prints max |a - r| / max r of the 5x5-rebinned projection for several seeds and history counts (noise falls like
1/sqrt(n) down to the reference's own 1.2e6-history noise floor)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import test_oracle_beamline_roi as T
from test_gpu_beamline_roi import grid_edges, NX, NY, NZ
from moquimc_b200 import capi

g = np.load(os.path.join(ROOT, "tests", "golden", "f4_roi_release.npz"))
meta = json.loads(str(g["meta"]))
mt = T.G.f4_mask_total()
ref_xy = g["water_dE_total_xy"]
r = T.rebin2(ref_xy, 5)
for n in (1_000_000, 4_000_000, 16_000_000):
    for seed in (123, 7, 99, 2024):
        e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
        xe, ye, ze = grid_edges()
        e.set_grid_hu(xe, ye, ze, np.zeros((NZ, NY, NX), dtype=np.int16))
        s = e.add_scorer(capi.SCORER_DOSE, "Dose")
        e.set_scorer_roi(s, mt)
        e.set_beamlets([capi.make_beamlet(meta["energy"], [0, 0, 0.5, 0, 0, -1], [meta["spot_size"]] * 2 + [0, 0, 0, 0], uniform=True)], [n])
        e.run(seed=seed, first=0, count=n)
        d = e.get_dense(s) / n
        a = T.rebin2(d.sum(axis=0), 5)
        dev = (a - r) / r.max()
        print("n %8d seed %5d: max |dev| %.4f  rms %.4f  mean %.5f  sum ratio %.5f" % (n, seed, np.abs(dev).max(), np.sqrt((dev[r > 0.5 * r.max()] ** 2).mean()), dev[r > 0.5 * r.max()].mean(), d.sum() / float(g["water_dE_total_total"])), flush=True)
