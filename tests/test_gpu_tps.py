"""tps_env (C++ front end over the C ABI) on a GPU: output files, scaling factors, statistical
stopping, the sparse Dij matrix and multi-device sharding.  The transport physics itself is covered by
test_gpu_parity.py; here the front end is checked against the same engine driven directly through
ctypes with the beam source `tps_env --dry-run` reports (identical counter-based streams => identical
dose up to fp64 summation order)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moquimc_b200 import capi, synthetic as S  # noqa: E402

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "moquimc_b200", "bin", "tps_env")
N = (64, 64, 40)
SP = (4.0, 4.0, 6.0)


def run_tps(inp, dry=False):
    r = subprocess.run([EXE] + (["--dry-run"] if dry else []) + [inp], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if dry:
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("DRYRUN ")][-1]
        return json.loads(line[len("DRYRUN "):])
    return r.stdout


def engine_for(root, src, scorers, per_spot=False, seed=12345, capacity=0):
    """the same run through ctypes: CT from the .mha the case wrote, beamlets from the dry run"""
    hu, origin = S.head_ct(N, SP, 1)
    g = src["grid"]
    xe = (np.float32(g["xe"][0]) + np.arange(N[0] + 1, dtype=np.float32) * np.float32(SP[0])).astype(np.float32)
    ye = (np.float32(g["ye"][0]) + np.arange(N[1] + 1, dtype=np.float32) * np.float32(SP[1])).astype(np.float32)
    ze = np.empty(N[2] + 1, dtype=np.float32)
    ze[0] = g["ze"][0]
    for i in range(1, N[2] + 1):
        ze[i] = ze[i - 1] + np.float32(SP[2])
    e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
    e.set_grid_hu(xe, ye, ze, hu)
    ids = [e.add_scorer(k, n, capacity) for k, n in scorers]
    beam = src["beams"][0]
    bl = [capi.make_beamlet(s["energy"], s["mean"], s["sigma"], uniform=False, sigma_energy=s["sigma_energy"],
                            rot=s["rot"], trans=s["trans"]) for s in beam["spots"]]
    for b in bl:
        b.energy_normal = 1
    hist = [s["histories"] for s in beam["spots"]]
    e.set_beamlets(bl, hist)
    st = e.run(seed, 0, sum(hist), per_spot=per_spot)
    assert st.histories == sum(hist)
    return e, ids, hist


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    """the synthetic case (ct.mha, machine.txt, plan.txt) every test of this module reads: written here, so that any
    test can run on its own (-k)"""
    r = str(tmp_path_factory.mktemp("tpsgpu"))
    S.make_case(r, n=N, spacing=SP, ParticlesPerHistory=2000.0, OutputDir=os.path.join(r, "o_dose"))
    return r


def test_per_beam_dose_files_and_scaling(root):
    inp = S.make_case(root, n=N, spacing=SP, ParticlesPerHistory=2000.0, OutputDir=os.path.join(root, "o_dose"))
    src = run_tps(inp, dry=True)
    out = run_tps(inp)
    total = sum(s["histories"] for s in src["beams"][0]["spots"])
    assert "Number of particles tracked %d" % total in out and "Time taken by MC engine" in out
    d = np.fromfile(os.path.join(root, "o_dose", "G000_0_Dose.raw"), dtype=np.float64)
    assert d.size == N[0] * N[1] * N[2] and d.max() > 0
    e, ids, _ = engine_for(root, src, [(capi.SCORER_DOSE, "Dose")])
    ref = e.get_dense(ids[0]).ravel() * (2000.0 * np.float32(1.1) * 30)   # ParticlesPerHistory * RBE * NumberOfFraction
    np.testing.assert_allclose(d, ref, rtol=1e-6, atol=ref.max() * 1e-12)
    # the beam enters from -y (gantry 0, iec2dicom): most dose on the anterior side of the isocentre plane
    vol = d.reshape(N[2], N[1], N[0])
    assert vol[:, : N[1] // 2, :].sum() > vol[:, N[1] // 2:, :].sum()


def test_letd_edep_scorers_and_mhd_output(root):
    od = os.path.join(root, "o_let")
    inp = os.path.join(root, "let.in")
    S.write_input(inp, root, od, Scorer="Dose,LETd,EnergyDeposition", ParticlesPerHistory=4000.0, OutputFormat="mhd")
    run_tps(inp)
    names = ["Dose", "LETd_numer", "LETd_denom", "EnergyDeposition"]
    vols = {}
    for n in names:
        hdr = open(os.path.join(od, "G000_0_%s.mhd" % n)).read()
        assert "DimSize = 64 64 40" in hdr and "ElementType = MET_DOUBLE" in hdr and "ElementDataFile = G000_0_%s.raw" % n in hdr
        vols[n] = np.fromfile(os.path.join(od, "G000_0_%s.raw" % n), dtype=np.float64)
    m = vols["LETd_denom"] > 0
    letd = vols["LETd_numer"][m] / vols["LETd_denom"][m]
    assert 0.2 < np.average(letd, weights=vols["LETd_denom"][m]) < 10.0   # keV/um-ish magnitudes, MeV/mm/(g/cm3)
    assert vols["EnergyDeposition"].sum() > vols["LETd_denom"].sum() * 0.99   # dE + local dE >= dE
    # without the B2 quirk the Dose of a 3-scorer run is the Dose of a 1-scorer run on the same streams
    inp1 = os.path.join(root, "let1.in")
    S.write_input(inp1, root, os.path.join(root, "o_let1"), Scorer="Dose", ParticlesPerHistory=4000.0)
    run_tps(inp1)
    d1 = np.fromfile(os.path.join(root, "o_let1", "G000_0_Dose.raw"), dtype=np.float64)
    np.testing.assert_allclose(vols["Dose"], d1, rtol=1e-9, atol=d1.max() * 1e-12)


def test_statistical_stopping(root):
    od = os.path.join(root, "o_stat")
    inp = os.path.join(root, "stat.in")
    S.write_input(inp, root, od, ParticlesPerHistory=4000.0, StoppingStatistics="true", StoppingCriteria=4.0, StatThreshold=0.5,
                  SaveStoppingStatistics="true")
    out = run_tps(inp)
    runs = [(int(a), float(b)) for a, b in re.findall(r"Run (\d+): current uncertainty ([0-9.eE+-]+) %", out)]
    assert runs and runs[-1][1] <= 4.0 and all(u > 4.0 for _, u in runs[:-1])
    unc = [u for _, u in runs]
    assert all(a >= b * 0.9 for a, b in zip(unc, unc[1:]))           # fresh streams per pass: falls ~ 1/sqrt(passes)
    d = np.fromfile(os.path.join(od, "G000_0_Dose.raw"), dtype=np.float64)
    s = np.fromfile(os.path.join(od, "G000_0_Dose_stat.raw"), dtype=np.float64)
    # calculate_average_results: the saved Dose is the accumulated one scaled back to one pass
    np.testing.assert_allclose(d * len(runs), s, rtol=1e-9, atol=s.max() * 1e-12)
    inp1 = os.path.join(root, "stat1.in")
    S.write_input(inp1, root, os.path.join(root, "o_stat1"), ParticlesPerHistory=4000.0)
    run_tps(inp1)
    d1 = np.fromfile(os.path.join(root, "o_stat1", "G000_0_Dose.raw"), dtype=np.float64)
    assert abs(d.sum() / d1.sum() - 1.0) < 0.02


def test_dij_npz_matches_engine(root):
    import scipy.sparse as sp
    od = os.path.join(root, "o_dij")
    inp = os.path.join(root, "dij.in")
    S.write_input(inp, root, od, Scorer="Dij", UnitWeights=300, OutputFormat="npz")
    src = run_tps(inp, dry=True)
    run_tps(inp)
    m = sp.load_npz(os.path.join(od, "G000_0_Dij.npz"))
    n_spots = len(src["beams"][0]["spots"])
    assert m.shape == (n_spots, N[0] * N[1] * N[2]) and m.format == "csr"
    assert (np.diff(m.indptr) > 0).all()                      # every spot deposits somewhere
    e, ids, hist = engine_for(root, src, [(capi.SCORER_DIJ, "Dij")], per_spot=True, capacity=4_000_037)
    assert set(hist) == {300}
    k1, k2, v = e.get_sparse(ids[0], scale=float(np.float32(1.0) * np.float32(1.1) * 30))   # UnitWeights: ParticlesPerHistory := 1
    ref = sp.csr_matrix((v, (k2, k1)), shape=m.shape)
    ref.sum_duplicates()
    m2 = m.copy()
    m2.sum_duplicates()
    m2.sort_indices()
    ref.sort_indices()
    assert m2.nnz == ref.nnz == m.nnz
    np.testing.assert_array_equal(m2.indices, ref.indices)
    np.testing.assert_allclose(m2.data, ref.data, rtol=1e-9)


def test_batches_and_density_scaling_change_nothing_or_everything(root):
    # MaxHistoriesPerBatch only splits the launches: same histories, same streams
    a, b = os.path.join(root, "o_b0"), os.path.join(root, "o_b1")
    i0, i1 = os.path.join(root, "b0.in"), os.path.join(root, "b1.in")
    S.write_input(i0, root, a, ParticlesPerHistory=4000.0)
    S.write_input(i1, root, b, ParticlesPerHistory=4000.0, MaxHistoriesPerBatch=1000)
    run_tps(i0)
    assert "batches expected" in run_tps(i1)
    d0 = np.fromfile(os.path.join(a, "G000_0_Dose.raw"), dtype=np.float64)
    d1 = np.fromfile(os.path.join(b, "G000_0_Dose.raw"), dtype=np.float64)
    np.testing.assert_allclose(d0, d1, rtol=1e-9, atol=d0.max() * 1e-12)
    # robust scenario: DensityScaling multiplies every voxel density (mqi_tps_env.hpp:768) -- a different dose
    # on the same streams (the range effect itself is checked on a water phantom in
    # test_gpu_parity.py::test_density_scaling_shortens_the_range)
    c = os.path.join(root, "o_b2")
    i2 = os.path.join(root, "b2.in")
    S.write_input(i2, root, c, ParticlesPerHistory=4000.0, DensityScaling=1.035)
    run_tps(i2)
    d2 = np.fromfile(os.path.join(c, "G000_0_Dose.raw"), dtype=np.float64)
    assert abs(d2.sum() / d0.sum() - 1.0) < 0.05 and np.abs(d2 - d0).max() > 0.01 * d0.max()


@pytest.mark.skipif(capi.device_count() < 2, reason="needs two GPUs")
def test_two_devices_give_the_single_device_dose(root):
    a, b = os.path.join(root, "o_g1"), os.path.join(root, "o_g2")
    i0, i1 = os.path.join(root, "g1.in"), os.path.join(root, "g2.in")
    S.write_input(i0, root, a, ParticlesPerHistory=2000.0)
    S.write_input(i1, root, b, ParticlesPerHistory=2000.0, GPUID="0,1")
    run_tps(i0)
    run_tps(i1)
    d0 = np.fromfile(os.path.join(a, "G000_0_Dose.raw"), dtype=np.float64)
    d1 = np.fromfile(os.path.join(b, "G000_0_Dose.raw"), dtype=np.float64)
    np.testing.assert_allclose(d0, d1, rtol=1e-9, atol=d0.max() * 1e-12)


def test_beamline_children_and_scoring_mask_through_the_front_end(tmp_path):
    """Range shifter (by ID) + aperture block from the text plan, ScoringMask + StatROI masks from .mha files:
    the front end must build the same world as the engine driven through ctypes with the geometry the dry
    run reports (identical streams => identical dose), name the outputs after the patient's child index and
    leave every voxel outside the mask at zero."""
    root = str(tmp_path)
    poly = [(-18, -12), (14, -12), (14, 16), (-18, 16)]
    extra = {"rangeshifter_ids": ["RS1"], "blocks": [poly], "block_thickness": 20.0, "block_tray_distance": 140.0}
    mask = np.zeros((N[2], N[1], N[0]), dtype=np.uint8)
    mask[8:34, 6:50, 14:52] = 1
    S.write_mask_mha(os.path.join(root, "mask.mha"), mask)
    od = os.path.join(root, "o_bl")
    inp = S.make_case(root, n=N, spacing=SP, n_layers=3, beam_extra=extra, ParticlesPerHistory=1500.0, OutputDir=od,
                      ScoringMask="true", Mask=os.path.join(root, "mask.mha"), Scorer="Dose,EnergyDeposition")
    src = run_tps(inp, dry=True)
    out = run_tps(inp)
    assert "RANGE SHIFTER added" in out and "APERTURE added" in out
    nodes = src["beams"][0]["beamline"]
    assert len(nodes) == 2 and src["scoring_roi_size"] == int(mask.sum())
    d = np.fromfile(os.path.join(od, "G000_2_Dose.raw"), dtype=np.float64).reshape(N[2], N[1], N[0])   # patient = child 2
    ed = np.fromfile(os.path.join(od, "G000_2_EnergyDeposition.raw"), dtype=np.float64).reshape(N[2], N[1], N[0])
    assert d[mask == 0].sum() == 0.0 and ed[mask == 0].sum() == 0.0 and d[mask == 1].sum() > 0
    # the same world through ctypes
    hu, origin = S.head_ct(N, SP, 1)
    g = src["grid"]
    xe = (np.float32(g["xe"][0]) + np.arange(N[0] + 1, dtype=np.float32) * np.float32(SP[0])).astype(np.float32)
    ye = (np.float32(g["ye"][0]) + np.arange(N[1] + 1, dtype=np.float32) * np.float32(SP[1])).astype(np.float32)
    ze = np.empty(N[2] + 1, dtype=np.float32)
    ze[0] = g["ze"][0]
    for i in range(1, N[2] + 1):
        ze[i] = ze[i - 1] + np.float32(SP[2])
    beam = src["beams"][0]
    rot, trans = beam["spots"][0]["rot"], beam["spots"][0]["trans"]
    e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
    e.set_grid_hu(xe, ye, ze, hu)
    rs, ap = nodes
    e.add_beamline_node(np.float32(rs["xe"]), np.float32(rs["ye"]), np.float32(rs["ze"]), np.float32([rs["rho0"]]), rot=rot, trans=trans)
    axe = (np.float32(ap["xe"][0]) + np.arange(ap["n"][0] + 1, dtype=np.float32)).astype(np.float32)
    aye = (np.float32(ap["ye"][0]) + np.arange(ap["n"][1] + 1, dtype=np.float32)).astype(np.float32)
    aze = (np.float32(ap["ze"][0]) + np.arange(ap["n"][2] + 1, dtype=np.float32)).astype(np.float32)
    xc, yc = axe[:-1] + np.float32(0.5), aye[:-1] + np.float32(0.5)
    open_xy = (xc[None, :] >= -18) & (xc[None, :] < 14) & (yc[:, None] >= -12) & (yc[:, None] < 16)
    assert int(open_xy.sum()) * ap["n"][2] == ap["open_voxels"]
    arho = np.broadcast_to(np.where(open_xy, np.float32(1e-8), np.float32(100.0)), (ap["n"][2],) + open_xy.shape).astype(np.float32)
    e.add_beamline_node(axe, aye, aze, arho.copy(), rot=rot, trans=trans)
    s_d = e.add_scorer(capi.SCORER_DOSE, "Dose")
    s_e = e.add_scorer(capi.SCORER_EDEP, "EnergyDeposition")
    e.set_scorer_roi(s_d, mask)
    e.set_scorer_roi(s_e, mask)
    bl = [capi.make_beamlet(s["energy"], s["mean"], s["sigma"], uniform=False, sigma_energy=s["sigma_energy"],
                            rot=s["rot"], trans=s["trans"]) for s in beam["spots"]]
    for b in bl:
        b.energy_normal = 1
    hist = [s["histories"] for s in beam["spots"]]
    e.set_beamlets(bl, hist)
    e.run(12345, 0, sum(hist))
    ref = e.get_dense(s_d) * (1500.0 * np.float32(1.1) * 30)
    assert ref.sum() > 0
    np.testing.assert_allclose(d, ref, rtol=1e-6, atol=ref.max() * 1e-12)
    # the range shifter costs range, the aperture removes the spots outside the opening: less dose than the open beam
    inp0 = S.make_case(os.path.join(root, "open"), n=N, spacing=SP, n_layers=3, ParticlesPerHistory=1500.0,
                       OutputDir=os.path.join(root, "o_open"), Scorer="Dose")
    run_tps(inp0)
    d0 = np.fromfile(os.path.join(root, "o_open", "G000_0_Dose.raw"), dtype=np.float64)
    assert 0.05 * d0.sum() < d.sum() < 0.9 * d0.sum()


def test_stat_roi_mask_drives_the_stopping_criterion(tmp_path):
    root = str(tmp_path)
    mask = np.zeros((N[2], N[1], N[0]), dtype=np.uint8)
    mask[12:28, 20:44, 20:44] = 1
    S.write_mask_mha(os.path.join(root, "roi.mha"), mask)
    od = os.path.join(root, "o_stat")
    inp = S.make_case(root, n=N, spacing=SP, ParticlesPerHistory=3000.0, OutputDir=od, StoppingStatistics="true",
                      SaveStoppingStatistics="true", StoppingCriteria="6.0", StatThreshold="0.5", MaxStatPasses=60,
                      StatROIMaskFilename=os.path.join(root, "roi.mha"))
    out = run_tps(inp)
    m = re.findall(r"Run (\d+): current uncertainty ([0-9.eE+-]+) %", out)
    assert m and float(m[-1][1]) <= 6.0
    ds = np.fromfile(os.path.join(od, "G000_0_Dose_stat.raw"), dtype=np.float64).reshape(N[2], N[1], N[0])
    dd = np.fromfile(os.path.join(od, "G000_0_Dose.raw"), dtype=np.float64).reshape(N[2], N[1], N[0])
    assert ds[mask == 0].sum() == 0.0 and ds[mask == 1].sum() > 0      # the stat scorers see only their roi
    assert dd[mask == 0].sum() > 0                                      # the Dose scorer keeps the DIRECT roi
