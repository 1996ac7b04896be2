"""GPU tests (run on the B200 box with -m gpu) of the scorer kinds and the source model that the C2 / C3 tests do not
reach, against the reference's own CPU runs: EnergyDeposition, track-averaged LET and the gaussian pbs beamlet
(tests/golden/c2_slabs150_{edep,lett}_release.npz, g1_gauss150_release.npz; generators in oracle/gen_golden.py,
the CPU restatement is pinned to the same files in tests/test_oracle_dose.py)."""
import os

import numpy as np
import pytest

import dose_metrics as M
from moquimc_b200 import capi
from test_gpu_parity import c1_beamlet, c1_engine
from test_oracle_dose import _quantile_sigma

pytestmark = pytest.mark.gpu


def slab_hu():
    hu = np.zeros((350, 200, 200), dtype=np.int16)
    hu[350 - 70:350 - 50] = 1000
    hu[350 - 100:350 - 70] = -741
    return hu


def depth_profiles(e, n_scorers, n_total, n_batches, seed):
    per = n_total // n_batches
    idd = {k: [] for k in range(n_scorers)}
    for b in range(n_batches):
        e.clear_scorers()
        e.run(seed, b * per, per)
        for k in range(n_scorers):
            idd[k].append((e.get_dense(k) / per).sum(axis=(1, 2)))
    return {k: np.mean(v, axis=0) for k, v in idd.items()}


def test_energy_deposition_against_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "c2_slabs150_edep_release.npz"))
    e = c1_engine(capi.PHYSICS_RELEASE, hu=slab_hu(), scorers=(capi.SCORER_EDEP,))
    n = 2_000_000
    e.set_beamlets([c1_beamlet(150.0, 10.0)], [n])
    idd = depth_profiles(e, 1, n, 4, seed=2718)[0]
    ref_idd, ref_tot = gold["Edep_idd"], float(gold["Edep_total"])
    assert abs(idd.sum() / ref_tot - 1.0) < 2e-3            # MeV deposited locally per primary history (143.9 of 150)
    assert abs(M.r80_mm(idd) - M.r80_mm(ref_idd)) < 0.1
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99


def test_track_averaged_let_against_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "c2_slabs150_lett_release.npz"))
    e = c1_engine(capi.PHYSICS_RELEASE, hu=slab_hu(), scorers=(capi.SCORER_LETT_NUMER, capi.SCORER_LETT_DENOM))
    n = 2_000_000
    e.set_beamlets([c1_beamlet(150.0, 10.0)], [n])
    p = depth_profiles(e, 2, n, 4, seed=1618)
    gn, gd = gold["LETt_numer_idd"], gold["LETt_denom_idd"]
    assert abs(p[1].sum() / float(gold["LETt_denom_total"]) - 1.0) < 2e-3    # track length per primary history [mm]
    assert abs(p[0].sum() / float(gold["LETt_numer_total"]) - 1.0) < 4e-3
    assert abs(M.r80_mm(p[1]) - M.r80_mm(gd)) < 0.15
    r10 = lambda a: a.reshape(35, 10).sum(axis=1)   # noqa: E731
    m = r10(gd) > 0.2 * r10(gd).max()
    assert np.abs((r10(p[0])[m] / r10(p[1])[m]) / (r10(gn)[m] / r10(gd)[m]) - 1.0).max() < 0.02


def test_gaussian_pencil_beam_against_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "g1_gauss150_release.npz"))
    e = c1_engine(capi.PHYSICS_RELEASE, scorers=(capi.SCORER_DOSE,))
    n = 2_000_000
    b = capi.make_beamlet(150.0, [0, 0, 0.5, 0, 0, -1], [4.0, 3.0, 0, 0.004, 0.003, 0], uniform=False, sigma_energy=1.5)
    e.set_beamlets([b], [n])
    e.run(seed=4, first=0, count=n)
    d = e.get_dense(0) / n
    idd, ref_idd = d.sum(axis=(1, 2)), gold["water_dE_total_idd"]
    assert abs(d.sum() / float(gold["water_dE_total_total"]) - 1.0) < 3e-3
    assert abs(M.r80_mm(idd) - M.r80_mm(ref_idd)) < 0.1
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99
    x = (np.arange(200) + 0.5) * 0.5 - 50.0
    xz, yz = d.sum(axis=1), d.sum(axis=2)
    for k0, k1 in ((340, 350), (290, 300), (240, 250), (205, 215)):   # 0-10, 50-60, 100-110, 135-145 mm depth
        for mine, ref in ((xz, gold["water_dE_total_xz"]), (yz, gold["water_dE_total_yz"])):
            a, r = _quantile_sigma(mine[k0:k1].sum(axis=0), x), _quantile_sigma(ref[k0:k1].sum(axis=0), x)
            assert abs(a / r - 1.0) < 0.02, (k0, a, r)


def test_debug_variant_in_heterogeneous_media_against_reference_golden(golden_dir):
    """-D__PHYSICS_DEBUG__ through bone and lung (tests/golden/c2_slabs150_debug.npz: the reference's own CPU
    phantom_env): the part of the delta-electron dose the reference loses to rsp(rho, 0) = inf (quirk B16) is lost
    here too, slab by slab."""
    gold = np.load(os.path.join(golden_dir, "c2_slabs150_debug.npz"))
    e = c1_engine(capi.PHYSICS_DEBUG, hu=slab_hu(), scorers=(capi.SCORER_DOSE,))
    n = 2_000_000
    e.set_beamlets([c1_beamlet(150.0, 10.0)], [n])
    idd = depth_profiles(e, 1, n, 4, seed=577)[0]
    ref_idd = gold["water_dE_total_idd"]
    assert abs(idd.sum() / float(gold["water_dE_total_total"]) - 1.0) < 3e-3
    assert abs(M.r80_mm(idd) - M.r80_mm(ref_idd)) < 0.1
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99
    for lo, hi in ((0, 50), (50, 70), (70, 100), (100, 160)):   # water, bone, lung, water (depth d mm <-> k = 349 - floor(d))
        a, r = idd[350 - hi:350 - lo].sum(), ref_idd[350 - hi:350 - lo].sum()
        assert abs(a / r - 1.0) < 0.005, (lo, hi, a, r)
