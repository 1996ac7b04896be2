"""GPU parity tests of SURVEY.md section 8(f) rows 3-4 (run on the B200 box with -m gpu):

  * beamline children of the world (range shifter slab, voxelised aperture) in front of the scored grid:
    the c_ind loop of transport_particles_patient (mqi_transport.hpp:162-240) over nodes built like
    create_rangeshifter / create_voxelized_aperture (mqi_tps_env.hpp:1605-1736);
  * CONTOUR regions of interest from mask volumes (mask_reader::mask_to_roi, mqi_file_handler.hpp:176-217;
    roi_t::idx, mqi_roi.hpp:48-58,127-137).

Everything goes through the C ABI; the oracle is the checker, on identical Philox streams.
"""
import numpy as np
import pytest

import dose_metrics as M
import oracle_lib as O
from moquimc_b200 import capi

pytestmark = pytest.mark.gpu

NX, NY, NZ = 100, 100, 200   # 1 x 1 x 1 mm voxels, z in [-200, 0]


def grid_edges():
    return capi.uniform_edges(-50, 50, NX), capi.uniform_edges(-50, 50, NY), capi.uniform_edges(-200, 0, NZ)


def water_rho():
    return np.full(NX * NY * NZ, O.hu_to_density(np.array([0]))[0], dtype=np.float32)


def rot_y(deg):
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)


def beamline_nodes(rot=None, trans=None):
    """[(xe, ye, ze, rho)]: a 40 mm range shifter slab (one voxel: grid3d(..., 2 edges per axis), density
    1.19 g/cm^3 like RangeshifterDensity) at z = 100..140 and a 20 mm thick aperture voxelised at 1 mm
    with a 24 x 16 mm opening (1e-8 open, 100 closed) at z = 40..60, both in the beam frame."""
    rs = (np.array([-150, 150], np.float32), np.array([-150, 150], np.float32), np.array([100, 140], np.float32),
          np.array([1.19e-3], np.float32))
    axe = capi.uniform_edges(-40, 40, 80)
    aye = capi.uniform_edges(-40, 40, 80)
    aze = capi.uniform_edges(40, 60, 20)
    xc = 0.5 * (axe[1:] + axe[:-1])
    yc = 0.5 * (aye[1:] + aye[:-1])
    open_xy = (np.abs(yc)[:, None] < 8.0) & (np.abs(xc)[None, :] < 12.0)
    arho = np.where(open_xy, np.float32(1e-8), np.float32(100.0)).astype(np.float32)
    arho = np.broadcast_to(arho, (20, 80, 80)).copy()
    return [rs, (axe, aye, aze, arho.ravel())]


def run_gpu(nodes, beamlet, n, seed, rot=None, trans=None, physics=capi.PHYSICS_RELEASE, kinds=(capi.SCORER_DOSE,)):
    e = capi.Engine(0, physics=physics)
    xe, ye, ze = grid_edges()
    e.set_grid_hu(xe, ye, ze, np.zeros((NZ, NY, NX), dtype=np.int16))
    for (bx, by, bz, brho) in nodes:
        e.add_beamline_node(bx, by, bz, brho, rot=rot, trans=trans)
    for k in kinds:
        e.add_scorer(k, "s%d" % k)
    e.set_beamlets([beamlet], [n])
    e.set_option("count_steps", 1)
    st = e.run(seed=seed, first=0, count=n)
    assert st.histories == n
    return e, st


def run_oracle(nodes, beamlet, n, seed, rot=None, trans=None, variant=O.VARIANT_RELEASE, kinds=(O.SCORER_DOSE,), roi=None):
    xe, ye, ze = grid_edges()
    keep = []
    gl = []
    for (bx, by, bz, brho) in nodes:
        g, k = O.make_grid(bx, by, bz, brho, rot=rot, trans=trans)
        gl.append(g)
        keep.append(k)
    g, k = O.make_grid(xe, ye, ze, water_rho())
    gl.append(g)
    keep.append(k)
    outs, st = O.transport(gl, variant, [beamlet], [n], seed=seed, h0=0, n=n, kinds=list(kinds), roi_members=roi)
    return [o.reshape(NZ, NY, NX) for o in outs], st


def test_beamline_nodes_match_oracle_same_streams():
    """Range shifter + aperture in front of a water phantom, beam along -z, identity node frames."""
    n = 12000
    nodes = beamline_nodes()
    src = dict(energy=150.0, mean=[0, 0, 180.0, 0, 0, -1], sigma=[20, 20, 0, 0, 0, 0], uniform=True)
    e, st = run_gpu(nodes, capi.make_beamlet(src["energy"], src["mean"], src["sigma"], uniform=True), n, seed=11)
    d = e.get_dense(0)
    (od,), ost = run_oracle(nodes, O.make_beamlet(src["energy"], src["mean"], src["sigma"], uniform=True), n, seed=11)
    assert od.sum() > 0
    assert abs(d.sum() / od.sum() - 1.0) < 5e-3
    gi, oi = d.sum(axis=(1, 2)), od.sum(axis=(1, 2))
    assert np.abs(gi - oi).max() / oi.max() < 0.03
    assert abs(M.r80_mm(gi) - M.r80_mm(oi)) < 0.15
    assert abs(st.steps / ost.steps - 1.0) < 5e-3   # steps in all three nodes
    # the aperture shapes the field: 24 x 16 mm opening (plus scatter) out of a 40 x 40 mm spot
    lat_x, lat_y = d.sum(axis=(0, 1)), d.sum(axis=(0, 2))
    xc = np.arange(NX) - 49.5
    assert lat_x[np.abs(xc) > 25].sum() < 0.03 * lat_x.sum()   # wide-angle nuclear secondaries only
    assert lat_y[np.abs(xc) > 20].sum() < 0.03 * lat_y.sum()
    assert lat_x[np.abs(xc) < 10].min() > 0.5 * lat_x.max()
    # the range shifter pulls the range back by about 40 mm x 1.19 x rsp
    e0, _ = run_gpu([], capi.make_beamlet(src["energy"], src["mean"], src["sigma"], uniform=True), 4000, seed=11)
    r_open = M.r80_mm(e0.get_dense(0).sum(axis=(1, 2)))
    r_rs = M.r80_mm(gi)
    assert 40.0 < r_open - r_rs < 55.0, (r_open, r_rs)


def test_rotated_beamline_matches_oracle_and_unrotated_geometry():
    """The beamline frame rotated by a gantry angle about y and shifted to an isocentre: node rot / trans
    (rotation_matrix_fwd, translation_vector) are exercised, and the secondaries' Rfwd (x - T) + T map."""
    n = 10000
    iso = np.array([3.0, -2.0, -100.0], dtype=np.float32)
    R = rot_y(90.0)   # beam frame -z  ->  world -x
    nodes = beamline_nodes()
    # beamlet in the beam frame, mapped into the world by the same transform as the beamline objects
    kw = dict(uniform=True, rot=R, trans=tuple(iso))
    bg = capi.make_beamlet(160.0, [0, 0, 180.0, 0, 0, -1], [20, 20, 0, 0, 0, 0], **kw)
    bo = O.make_beamlet(160.0, [0, 0, 180.0, 0, 0, -1], [20, 20, 0, 0, 0, 0], **kw)
    e, st = run_gpu(nodes, bg, n, seed=5, rot=R, trans=iso, physics=capi.PHYSICS_DEBUG)
    d = e.get_dense(0)
    (od,), ost = run_oracle(nodes, bo, n, seed=5, rot=R, trans=iso, variant=O.VARIANT_DEBUG)
    assert od.sum() > 0
    assert abs(d.sum() / od.sum() - 1.0) < 5e-3
    # the beam now runs along -x: depth profile along x
    gi, oi = d.sum(axis=(0, 1)), od.sum(axis=(0, 1))
    assert np.abs(gi - oi).max() / oi.max() < 0.03
    # field centred on the isocentre in y (opening 16 mm) and z (opening 24 mm: beam x -> world -z)
    lat_y = d.sum(axis=(0, 2))
    yc = np.arange(NY) - 49.5
    assert abs((lat_y * yc).sum() / lat_y.sum() - iso[1]) < 0.5
    lat_z = d.sum(axis=(1, 2))
    zc = np.arange(NZ) - 199.5
    assert abs((lat_z * zc).sum() / lat_z.sum() - iso[2]) < 0.7
    osteps = ost.steps - ost.delta_events   # debug: the oracle counts the zero-energy delta daughters
    assert abs(st.steps / osteps - 1.0) < 1e-2


def test_closed_aperture_stops_everything_and_vacuum_node_is_transparent():
    xe = np.array([-60, 60], np.float32)
    ze = np.array([40, 60], np.float32)
    closed = [(xe, xe, ze, np.array([100.0], np.float32))]
    b = capi.make_beamlet(120.0, [0, 0, 100.0, 0, 0, -1], [10, 10, 0, 0, 0, 0], uniform=True)
    e, st = run_gpu(closed, b, 3000, seed=1)
    assert e.get_dense(0).sum() == 0.0
    vac = [(xe, xe, ze, np.array([1e-8], np.float32))]
    e1, _ = run_gpu(vac, b, 3000, seed=1)
    e0, _ = run_gpu([], b, 3000, seed=1)
    # same histories, same streams; passing through the vacuum node re-rounds the entry point (fp32), so a
    # few histories differ in a branch: integral quantities agree far below the statistical error
    d1, d0 = e1.get_dense(0), e0.get_dense(0)
    assert d0.sum() > 0 and abs(d1.sum() / d0.sum() - 1.0) < 2e-3
    e1.clear_beamline()
    e1.clear_scorers()
    e1.run(seed=1, first=0, count=3000)
    np.testing.assert_allclose(e1.get_dense(0), d0, rtol=1e-9, atol=1e-24)


def make_mask_total():
    """Sum of two overlapping 0/1 masks: a box around the beam axis plus a slab; where they overlap the sum
    is 2, which neither opens nor closes a run (reference quirk kept)."""
    m1 = np.zeros((NZ, NY, NX), dtype=np.uint8)
    m1[60:190, 30:70, 35:65] = 1
    m2 = np.zeros((NZ, NY, NX), dtype=np.uint8)
    m2[100:120, 40:60, 35:50] = 1
    return m1 + m2


def test_mask_to_roi_runs_and_device_roi_size():
    mt = make_mask_total()
    start, stride, member = O.mask_to_roi(mt)
    assert len(start) == len(stride) > 0
    assert member.sum() == stride.sum()
    # inside [start, start+stride) of every run, nowhere else
    chk = np.zeros(mt.size, dtype=np.uint8)
    for s, t in zip(start[:50], stride[:50]):
        assert mt.ravel()[s] == 1 and (s == 0 or not member[s - 1])
        chk[s:s + t] = 1
    assert np.array_equal(chk[:start[50]], member[:start[50]])
    # a row that begins inside the overlap (sum 2) opens its run only at the first voxel with sum 1
    row, mrow = mt[110, 50], member.reshape(NZ, NY, NX)[110, 50]
    assert row[35] == 2 and row[50] == 1 and not mrow[35:50].any() and mrow[50:65].all()
    e = capi.Engine(0)
    xe, ye, ze = grid_edges()
    e.set_grid_hu(xe, ye, ze, np.zeros((NZ, NY, NX), dtype=np.int16))
    s = e.add_scorer(capi.SCORER_DOSE, "Dose")
    assert e.set_scorer_roi(s, mt) == int(member.sum())
    assert e.set_scorer_roi(s, None) == NX * NY * NZ
    with pytest.raises(capi.MqiError):
        e.set_scorer_roi(s, mt.ravel()[:-1])


def test_roi_scoring_matches_oracle_and_masks_the_direct_dose():
    n = 8000
    mt = make_mask_total()
    _, _, member = O.mask_to_roi(mt)
    member3 = member.reshape(NZ, NY, NX).astype(bool)
    b = capi.make_beamlet(150.0, [0, 0, 0.5, 0, 0, -1], [25, 25, 0, 0, 0, 0], uniform=True)
    kinds = (capi.SCORER_DOSE, capi.SCORER_EDEP, capi.SCORER_DOSE_SQ)
    e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
    xe, ye, ze = grid_edges()
    e.set_grid_hu(xe, ye, ze, np.zeros((NZ, NY, NX), dtype=np.int16))
    for k in kinds:
        e.add_scorer(k, "s%d" % k)
    e.set_scorer_roi(0, mt)          # Dose: masked
    e.set_scorer_roi(2, mt)          # DoseSquare: masked (the stat scorers carry their own roi)
    e.set_beamlets([b], [n])
    e.run(seed=77, first=0, count=n)
    d_roi, edep, dsq = e.get_dense(0), e.get_dense(1), e.get_dense(2)
    assert d_roi[~member3].sum() == 0.0 and dsq[~member3].sum() == 0.0
    assert d_roi[member3].sum() > 0
    assert edep[~member3].sum() > 0   # the unmasked scorer of the same run still scores everywhere
    # same histories without the roi: the masked dose is the direct dose inside the roi
    e.clear_scorers()
    e.set_scorer_roi(0, None)
    e.run(seed=77, first=0, count=n)
    d_all = e.get_dense(0)
    np.testing.assert_allclose(d_roi[member3], d_all[member3], rtol=1e-9, atol=1e-24)
    # oracle with the same roi on the same streams
    ob = O.make_beamlet(150.0, [0, 0, 0.5, 0, 0, -1], [25, 25, 0, 0, 0, 0], uniform=True)
    (od, oe, _), _ = run_oracle([], ob, n, seed=77, kinds=(O.SCORER_DOSE, O.SCORER_EDEP, O.SCORER_DOSE_SQ),
                                roi=[member, None, member])
    assert od[~member3].sum() == 0.0
    assert abs(d_roi.sum() / od.sum() - 1.0) < 5e-3
    gi, oi = d_roi.sum(axis=(1, 2)), od.sum(axis=(1, 2))
    assert np.abs(gi - oi).max() / oi.max() < 0.03
    assert abs(edep.sum() / oe.sum() - 1.0) < 5e-3
    # stopping criterion over the roi: voxels outside hold zeros and drop out of the count
    s, cnt, mx = e.stat_partial(0, 2, n, 0.5)
    assert cnt > 0


# ------------------------------------------------------------------------------------------------
# against fixtures written by the reference's own CPU code (oracle/gen_golden.py f3 / f4)
# ------------------------------------------------------------------------------------------------
def test_gpu_beamline_against_reference_golden(golden_dir):
    import json
    import os
    import test_oracle_beamline_roi as T
    G = T.G
    g = np.load(os.path.join(golden_dir, "f3_beamline_release.npz"))
    meta = json.loads(str(g["meta"]))
    fr = np.array(G.f3_frame(), dtype=np.float32)
    zlo, zhi, half, dens = G.F3_RS
    e2 = np.array([-half, half], dtype=np.float32)
    rs = (e2, e2, np.array([zlo, zhi], dtype=np.float32), np.array([np.float32(dens * 1e-3)]))
    zlo, zhi, half, ohx, ohy = G.F3_AP
    nxy, nz = int(np.ceil(2 * half)), int(np.ceil(zhi - zlo))
    axe = (np.float32(-half) + np.arange(nxy + 1, dtype=np.float32)).astype(np.float32)
    aze = (np.float32(zlo) + np.arange(nz + 1, dtype=np.float32)).astype(np.float32)
    xc = axe[:-1] + np.float32(0.5)
    open_xy = (np.abs(xc)[None, :] < ohx) & (np.abs(xc)[:, None] < ohy)
    arho = np.broadcast_to(np.where(open_xy, np.float32(1e-8), np.float32(100.0)), (nz, nxy, nxy)).astype(np.float32)
    n = 1_500_000
    b = capi.make_beamlet(meta["energy"], [0, 0, 180.0, 0, 0, -1], [meta["spot_size"]] * 2 + [0, 0, 0, 0], uniform=True)
    e, st = run_gpu([rs, (axe, axe, aze, arho.ravel())], b, n, seed=99, rot=fr[:9], trans=fr[9:])
    d = e.get_dense(0) / n
    ref_idd, ref_xy, ref_tot = g["water_dE_total_idd"], g["water_dE_total_xy"], float(g["water_dE_total_total"])
    assert abs(d.sum() / ref_tot - 1.0) < 3 * float(g["water_dE_total_total_se"]) / ref_tot + 2e-3
    idd = d.sum(axis=(1, 2))
    assert abs(M.r80_mm(idd) - M.r80_mm(ref_idd)) < 0.1
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99
    # lateral field shape behind the rotated, shifted aperture: 2-D gamma 1 % / 1 mm on the xy projection
    rate, _, _ = M.gamma_2d(ref_xy, d.sum(axis=0), (1.0, 1.0), dd=0.02)
    assert rate >= 0.97, rate   # the reference projection itself carries about 1.5 % noise per pixel (1.2e6 histories)


def test_gpu_roi_against_reference_golden(golden_dir):
    import json
    import os
    import test_oracle_beamline_roi as T
    g = np.load(os.path.join(golden_dir, "f4_roi_release.npz"))
    meta = json.loads(str(g["meta"]))
    mt = T.G.f4_mask_total()
    _, _, member = O.mask_to_roi(mt)
    m3 = member.reshape(NZ, NY, NX).astype(bool)
    n = 4_000_000
    e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
    xe, ye, ze = grid_edges()
    e.set_grid_hu(xe, ye, ze, np.zeros((NZ, NY, NX), dtype=np.int16))
    s = e.add_scorer(capi.SCORER_DOSE, "Dose")
    assert e.set_scorer_roi(s, mt) == int(member.sum())
    e.set_beamlets([capi.make_beamlet(meta["energy"], [0, 0, 0.5, 0, 0, -1], [meta["spot_size"]] * 2 + [0, 0, 0, 0], uniform=True)], [n])
    e.run(seed=123, first=0, count=n)
    d = e.get_dense(s) / n
    assert d[~m3].sum() == 0.0 and np.all(d.sum(axis=0)[m3.any(axis=0)] > 0)
    ref_idd, ref_xy, ref_tot = g["water_dE_total_idd"], g["water_dE_total_xy"], float(g["water_dE_total_total"])
    assert abs(d.sum() / ref_tot - 1.0) < 3 * float(g["water_dE_total_total_se"]) / ref_tot + 2e-3
    idd = d.sum(axis=(1, 2))
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99
    # lateral projection on 5x5-rebinned pixels.  The reference run itself (1.2e6 histories) carries 0.9 % noise
    # per pixel, which is the floor of this comparison whatever n is (scripts/roi_noise_probe.py: rms 0.9-1.0 %,
    # maximum over the ~100 pixels 1.8-2.6 % for n >= 4e6 with either Philox round count, and 2.1-3.9 % at n = 1e6)
    a, r = T.rebin2(d.sum(axis=0), 5), T.rebin2(ref_xy, 5)
    dev = (a - r) / r.max()
    assert np.sqrt((dev[r > 0.5 * r.max()] ** 2).mean()) < 0.0125
    assert np.abs(dev).max() < 0.035
