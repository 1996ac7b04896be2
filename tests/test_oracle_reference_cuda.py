"""CPU tests that pin the restatement (oracle/mqi_oracle.c) and the numpy restatement of the stopping criterion to
fixtures produced by the reference's own CUDA path on a B200 (oracle/gen_golden_gpu.py; tests/golden/README.md).
The same fixtures are what tests/test_gpu_reference_cuda.py holds the CUDA path to."""
import json
import os

import numpy as np
import pytest

import dose_metrics as M
import oracle_lib as O
from moquimc_b200 import synthetic as S


def load(golden_dir, name):
    path = os.path.join(golden_dir, name)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % name)
    g = np.load(path)
    return g, json.loads(str(g["meta"]))


def criterion(s1, s2, n, threshold):
    """calculate_standard_deviation (kernel_functions/mqi_variables.hpp:20-48) + the host part of calculate_stat
    (mqi_tps_env.hpp:1409-1425) in the reference's precision: double sums, float sigma and mean per voxel."""
    sd = np.zeros(s1.size, dtype=np.float32)
    mean = np.zeros(s1.size, dtype=np.float32)
    occ = s1 > 0       # key1 != empty_pair: the slot was touched
    m = s1[occ] / n
    sd[occ] = np.sqrt(((s2[occ] / n) - m * m) / (n - 1)).astype(np.float32)
    mean[occ] = m.astype(np.float32)
    dmax = float(mean.max())
    sel = mean.astype(np.float64) > dmax * threshold
    value = float((sd[sel] / mean[sel]).astype(np.float64).sum() / sel.sum())
    return sd, mean, value, int(sel.sum()), dmax


@pytest.mark.parametrize("tag", ["a", "b"])
def test_stopping_criterion_restatement_matches_the_reference_kernel(golden_dir, tag):
    g, meta = load(golden_dir, "a15_stat_release.npz")
    n, thr = int(g[tag + "_n"]), float(g[tag + "_threshold"])
    sd, mean, value, count, dmax = criterion(g[tag + "_sum"], g[tag + "_sumsq"], n, thr)
    ref_sd, ref_mean = g[tag + "_sd"], g[tag + "_mean"]
    np.testing.assert_array_equal(mean, ref_mean)                        # double / int rounded to float: exact
    ok = ref_sd > 0
    np.testing.assert_allclose(sd[ok], ref_sd[ok], rtol=3e-7)            # sqrtf under --use_fast_math: one ulp
    assert np.array_equal(sd > 0, ok)
    assert count == int(g[tag + "_count"])
    np.testing.assert_allclose(dmax, float(g[tag + "_dose_max"]), rtol=1e-7)
    np.testing.assert_allclose(value, float(g[tag + "_value"]), rtol=3e-7)
    # the stat pair is the dose and the sum of its squared step contributions: 0 < sum d^2 <= (sum d)^2
    s1, s2 = g[tag + "_sum"], g[tag + "_sumsq"]
    assert ((s2 > 0) == (s1 > 0)).all() and (s2 <= s1 * s1 * (1 + 1e-12)).all()


def slab_rho():
    hu = np.zeros((350, 200, 200), dtype=np.int16)
    hu[350 - 70:350 - 50] = 1000
    hu[350 - 100:350 - 70] = -741
    lut = O.hu_to_density(np.arange(-1000, 2996))
    return lut[hu.astype(np.int64) + 1000].astype(np.float32)


@pytest.mark.parametrize("variant", ["release", "debug"])
def test_oracle_dose_square_against_reference_cuda(golden_dir, variant):
    g, meta = load(golden_dir, "c2_slabs150_dose2.npz")
    grid, keep = O.make_grid(O.uniform_edges(-50, 50, 200), O.uniform_edges(-50, 50, 200), O.uniform_edges(-350, 0, 350), slab_rho())
    n = 24000
    b = O.make_beamlet(150.0, [0, 0, 0.5, 0, 0, -1], [10.0, 10.0, 0, 0, 0, 0], uniform=True)
    var = O.VARIANT_DEBUG if variant == "debug" else O.VARIANT_RELEASE
    (d, d2), st = O.transport(grid, var, [b], [n], seed=8, h0=0, n=n, kinds=[O.SCORER_DOSE, O.SCORER_DOSE_SQ])
    pre = variant + "_"
    assert abs(d.sum() / n / float(g[pre + "Dose_total"]) - 1.0) < 4e-3
    # sum of squared step doses per history: the reference's eight runs of 5e5 histories give its standard error;
    # this run has 24 000 histories, i.e. sqrt(5e5 * 8 / 24000) times that error
    ref, ref_se = float(g[pre + "Dose2_total"]), float(g[pre + "Dose2_total_se"])
    sig = ref_se * np.sqrt(1.0 + meta["histories"] / n)
    assert abs(d2.sum() / n - ref) < 4.0 * sig, (d2.sum() / n, ref, sig)
    idd2 = d2.reshape(350, 200, 200).sum(axis=(1, 2)) / n
    r10 = lambda a: a.reshape(35, 10).sum(axis=1)   # noqa: E731
    m = r10(g[pre + "Dose2_idd"]) > 0.1 * r10(g[pre + "Dose2_idd"]).max()
    s10 = np.sqrt(r10(g[pre + "Dose2_idd_se"] ** 2)) * np.sqrt(1.0 + meta["histories"] / n)
    assert (np.abs(r10(idd2) - r10(g[pre + "Dose2_idd"]))[m] < np.maximum(5.0 * s10[m], 0.05 * r10(g[pre + "Dose2_idd"])[m])).all()


def test_oracle_on_the_c3like_head_case_against_reference_cuda(golden_dir):
    """Heterogeneous CT (skull shell, air cavities), 20 oblique gaussian pbs beamlets of 90 ... 128 MeV: the
    restatement against the dense dose and the Dij row totals of the reference's CUDA run."""
    g, meta = load(golden_dir, "c3like_head_release.npz")
    nx, ny, nz = meta["nxyz"]
    lx, ly, lz = meta["lxyz"]
    hu, _ = S.head_ct(tuple(meta["nxyz"]), tuple(meta["spacing"]), seed=meta["hu_seed"])
    lut = O.hu_to_density(np.arange(-1000, 2996))
    rho = lut[hu.astype(np.int64) + 1000].astype(np.float32)
    grid, keep = O.make_grid(O.uniform_edges(-lx / 2, lx / 2, nx), O.uniform_edges(-ly / 2, ly / 2, ny),
                             O.uniform_edges(-lz / 2, lz / 2, nz), rho)
    gx, gy, pitch = meta["grid"]
    sx, sy, sxp, syp, se = meta["gauss"]
    bl = []
    for i in range(int(gx * gy)):
        ox = (i % gx - 0.5 * (gx - 1)) * pitch
        oy = (i // gx - 0.5 * (gy - 1)) * pitch
        bl.append(O.make_beamlet(meta["e0"] + i * meta["de"], [ox, oy, meta["spot_z"], 0, 0, -1], [sx, sy, 0, sxp, syp, 0],
                                 uniform=False, sigma_energy=se, rot=np.array(meta["rot"], dtype=np.float32).reshape(3, 3)))
    per = 3000
    ns = len(bl)
    (tab,), st = O.transport(grid, O.VARIANT_RELEASE, bl, [per] * ns, seed=3, h0=0, n=per * ns, kinds=[O.SCORER_DIJ],
                             per_spot=True, dij_capacity=8_000_003)
    rows = np.zeros((ns, nz * ny * nx))
    np.add.at(rows, (tab["key2"], tab["key1"]), tab["value"] / per)
    ref_tot = g["dij_row_total"].astype(np.float64)
    tot = rows.sum(axis=1)
    assert abs(tot.sum() / ref_tot.sum() - 1.0) < 4e-3
    assert np.abs(tot / ref_tot - 1.0).max() < 0.03
    dense = rows.sum(axis=0).reshape(nz, ny, nx) / ns
    ref = g["dose_q"].astype(np.float64) * (meta["dose_max"] / meta["dose_levels"])
    assert abs(dense.sum() / ref.sum() - 1.0) < 4e-3
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for ax, sp in ((zz, 2.5), (yy, 1.0), (xx, 1.0)):
        assert abs((ax * dense).sum() / dense.sum() - (ax * ref).sum() / ref.sum()) * sp < 0.25
    idd, ref_idd = rows.reshape(ns, nz, -1).sum(axis=2), g["dij_row_idd"].astype(np.float64)
    for i in (0, 9, 19):
        assert M.gamma_1d(ref_idd[i], idd[i], 2.5, dd=0.03, dta_mm=2.5)[0] >= 0.95, i
