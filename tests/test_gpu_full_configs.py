"""BASELINE.json configs C3-C5 at their full sizes on one GPU (run on the B200 box with -m gpu), checked
through size-independent properties: the statistical stopping criterion is met, Dij rows add up to the
dense dose and are reproducible spot by spot (what makes sharding spots over GPUs reduction-free), setup
shifts move the dose by the shift and density scaling moves the range the right way.  Throughputs of
these runs are printed (pytest -s) and quoted in DESIGN.md.

  C3  synthetic head-and-neck CT 512 x 512 x 200, ~2 000-spot PBS plan, Dose, 1 % statistical stopping
  C4  Dij: 5 000 spots x UnitWeights 1e4 histories on a 256 x 256 x 150 CT, sparse hash scoring
  C5  robust scenarios of the C3 plan: +-3 mm setup shift, +-3.5 % density scaling
"""
import os
import re
import subprocess
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moquimc_b200 import capi, synthetic as S  # noqa: E402

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "moquimc_b200", "bin", "tps_env")
N3, SP3 = (512, 512, 200), (1.0, 1.0, 2.5)


def run_tps(inp):
    t = time.time()
    r = subprocess.run([EXE, inp], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout, time.time() - t


def kernel_rate(out):
    m = re.findall(r"Transport kernels ([0-9.eE+-]+) ms on (\d+) GPU\(s\): ([0-9.eE+-]+) histories/s", out)
    return float(m[-1][2]) if m else float("nan")


@pytest.fixture(scope="module")
def c3_root(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("c3"))
    S.make_case(root, n=N3, spacing=SP3, n_layers=25, pitch=5.0, half_width=25.0, ParticlesPerHistory=400.0)
    return root


def test_c3_head_and_neck_plan_with_one_percent_stopping(c3_root):
    od = os.path.join(c3_root, "o_c3")
    inp = os.path.join(c3_root, "c3.in")
    S.write_input(inp, c3_root, od, ParticlesPerHistory=400.0, StoppingStatistics="true", StoppingCriteria=1.0,
                  StatThreshold=0.5, MaxStatPasses=400)
    out, wall = run_tps(inp)
    runs = [(int(a), float(b)) for a, b in re.findall(r"Run (\d+): current uncertainty ([0-9.eE+-]+) %", out)]
    tracked = int(re.findall(r"Number of particles tracked (\d+)", out)[-1])
    n_spots = int(re.findall(r"beamlets (\d+)", out)[-1])
    assert n_spots >= 1900
    assert runs and runs[-1][1] <= 1.0 and all(u > 1.0 for _, u in runs[:-1])
    unc = np.array([u for _, u in runs])
    if len(unc) > 3:   # fresh streams per pass: the uncertainty falls like 1 / sqrt(passes)
        assert abs(unc[-1] / unc[0] * np.sqrt(len(unc)) - 1.0) < 0.25
    d = np.fromfile(os.path.join(od, "G000_0_Dose.raw"), dtype=np.float64)
    assert d.size == N3[0] * N3[1] * N3[2] and d.max() > 0 and np.isfinite(d).all()
    vol = d.reshape(N3[2], N3[1], N3[0])
    # the 50 mm wide field around the isocentre axis: nothing far outside it
    lat = vol.sum(axis=(0, 1))
    xc = np.arange(N3[0]) - (N3[0] - 1) / 2.0
    assert lat[np.abs(xc) > 60].sum() < 0.02 * lat.sum()
    print("\nC3: %d spots, %d passes, %d histories, final uncertainty %.3f %%, transport %.3e histories/s, wall %.1f s"
          % (n_spots, len(runs), tracked, runs[-1][1], kernel_rate(out), wall))


def test_c5_robust_scenarios_shift_and_density(c3_root):
    """Five of the 21 scenarios (the rest are the same code with other numbers; they are independent replicas)."""
    res = {}
    for name, kw in (("nominal", {}), ("xp3", {"XShift": 3.0}), ("xm3", {"XShift": -3.0}),
                     ("dense", {"DensityScaling": 1.035}), ("light", {"DensityScaling": 0.965})):
        od = os.path.join(c3_root, "o_c5_" + name)
        inp = os.path.join(c3_root, "c5_%s.in" % name)
        S.write_input(inp, c3_root, od, ParticlesPerHistory=2000.0, **kw)
        out, wall = run_tps(inp)
        d = np.fromfile(os.path.join(od, "G000_0_Dose.raw"), dtype=np.float64).reshape(N3[2], N3[1], N3[0])
        res[name] = (d.sum(axis=(0, 1)), d.sum(axis=(0, 2)), d.sum(), kernel_rate(out))
    xi, yi = np.arange(N3[0], dtype=np.float64), np.arange(N3[1], dtype=np.float64)

    def cx(name):
        return (res[name][0] * xi).sum() / res[name][0].sum()

    def cy(name):
        return (res[name][1] * yi).sum() / res[name][1].sum()

    # XShift moves the CT edges by +-3 mm under a fixed beam: the dose moves by -+3 voxels of 1 mm in the CT's index space
    assert abs((cx("xp3") - cx("nominal")) + 3.0) < 0.3
    assert abs((cx("xm3") - cx("nominal")) - 3.0) < 0.3
    # the beam enters from -y (gantry 0): denser tissue stops it earlier (smaller y), lighter tissue later
    assert cy("dense") < cy("nominal") < cy("light")
    shift = cy("light") - cy("dense")
    assert 0.5 < shift < 15.0, shift
    print("\nC5: dose centroid y  dense %.2f  nominal %.2f  light %.2f voxels; transport %.3e histories/s"
          % (cy("dense"), cy("nominal"), cy("light"), res["nominal"][3]))


def test_c4_dij_five_thousand_spots_full_size():
    n = (256, 256, 150)
    sp = (1.5, 1.5, 2.0)
    hu, origin = S.head_ct(n, sp, seed=4)
    xe = (np.float32(origin[0] - sp[0] / 2) + np.arange(n[0] + 1, dtype=np.float32) * np.float32(sp[0])).astype(np.float32)
    ye = (np.float32(origin[1] - sp[1] / 2) + np.arange(n[1] + 1, dtype=np.float32) * np.float32(sp[1])).astype(np.float32)
    ze = (np.float32(origin[2] - sp[2] / 2) + np.arange(n[2] + 1, dtype=np.float32) * np.float32(sp[2])).astype(np.float32)
    # 5 000 spots: 20 energy layers x 250 positions, beam along +y from outside the head
    rng = np.random.default_rng(5)
    g = np.arange(-30.0, 30.0 + 1e-6, 4.0)
    pos = [(x, z) for x in g for z in g if x * x + z * z <= 30.0 ** 2 + 1e-6]
    energies = np.linspace(80.0, 160.0, 20)
    # beam along +y: phsp "z'" is the normalised component, so rotate the beam frame (x, y, z) -> world (x, -z, y)
    R = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32)
    bl = []
    for e_mev in energies:
        idx = rng.choice(len(pos), size=250, replace=True)
        for i in idx:
            x, z = pos[i]
            bl.append(capi.make_beamlet(float(e_mev), [x, z, 250.0, 0, 0, -1], [3.0, 3.0, 0.0, 0.003, 0.003, 0.0], uniform=False,
                                        sigma_energy=0.6, rot=R))
    n_spots, per = len(bl), 10_000
    assert n_spots == 5000
    e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
    e.set_grid_hu(xe, ye, ze, hu)
    cap = 393_216_000 | 1          # the reference's fixed table size (mqi_tps_env.hpp:922), 6.3 GB here
    s_dij = e.add_scorer(capi.SCORER_DIJ, "Dij", capacity=cap)
    s_dose = e.add_scorer(capi.SCORER_DOSE, "Dose")
    e.set_beamlets(bl, [per] * n_spots)
    t = time.time()
    st = e.run(seed=77, first=0, count=n_spots * per, per_spot=True)
    wall = time.time() - t
    assert st.histories == n_spots * per and st.dij_table_full == 0
    k1, k2, v = e.get_sparse(s_dij)
    dense = e.get_dense(s_dose).ravel()
    assert dense.sum() > 0
    nnz = len(v)
    assert (np.bincount(k2, minlength=n_spots) > 0).all() and k2.max() == n_spots - 1   # every spot has a row
    acc = np.bincount(k1, weights=v, minlength=dense.size)
    np.testing.assert_allclose(acc, dense, rtol=1e-9, atol=dense.max() * 1e-13)   # rows add up to the dense dose
    # spot rows do not depend on the other spots (counter-based streams): the first 64 spots alone give the same rows,
    # which is why sharding spots over GPUs needs no reduction
    sel = k2 < 64
    ref = {(int(a), int(b)): c for a, b, c in zip(k1[sel], k2[sel], v[sel])}
    e.clear_scorers()
    e.run(seed=77, first=0, count=64 * per, per_spot=True)
    a1, a2, av = e.get_sparse(s_dij)
    got = {(int(a), int(b)): c for a, b, c in zip(a1, a2, av)}
    assert got.keys() == ref.keys()
    kk = sorted(ref)
    np.testing.assert_allclose([got[k] for k in kk], [ref[k] for k in kk], rtol=1e-9)
    print("\nC4: %d spots x %d histories, nnz %d (load %.3f of %d slots), kernel %.1f ms = %.3e histories/s, wall %.1f s"
          % (n_spots, per, nnz, nnz / cap, cap, st.kernel_ms, st.histories / (st.kernel_ms * 1e-3), wall))
