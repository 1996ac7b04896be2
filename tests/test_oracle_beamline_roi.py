"""CPU tests: the oracle's multi-node transport and CONTOUR roi against fixtures produced by the
reference's own CPU code (oracle/ref_harness.cpp: beamline children built like create_rangeshifter /
create_voxelized_aperture, roi from mask_reader::mask_to_roi; generator oracle/gen_golden.py f3 / f4)."""
import json
import os
import sys

import numpy as np

import dose_metrics as M
import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gen_golden as G   # noqa: E402  (fixture geometry only; nothing from /root/reference is touched)

NX, NY, NZ = 100, 100, 200


def edges():
    return O.uniform_edges(-50, 50, NX), O.uniform_edges(-50, 50, NY), O.uniform_edges(-200, 0, NZ)


def water():
    return np.full(NX * NY * NZ, O.hu_to_density(np.array([0]))[0], dtype=np.float32)


def rebin2(a, f):
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(a.shape[0] // f, f, a.shape[1] // f, f).sum(axis=(1, 3))


def f3_nodes():
    """The harness's beamline children: grid3d(-half, half, 2, ...) slab and 1 mm voxelised aperture."""
    fr = np.array(G.f3_frame(), dtype=np.float32)
    rot, trans = fr[:9], fr[9:]
    zlo, zhi, half, dens = G.F3_RS
    e2 = np.array([-half, half], dtype=np.float32)
    rs, k0 = O.make_grid(e2, e2, np.array([zlo, zhi], dtype=np.float32), np.array([np.float32(dens * 1e-3)]), rot=rot, trans=trans)
    zlo, zhi, half, ohx, ohy = G.F3_AP
    nxy, nz = int(np.ceil(2 * half)), int(np.ceil(zhi - zlo))
    xe = (np.float32(-half) + np.arange(nxy + 1, dtype=np.float32)).astype(np.float32)
    ze = (np.float32(zlo) + np.arange(nz + 1, dtype=np.float32)).astype(np.float32)
    xc = xe[:-1] + np.float32(0.5)
    open_xy = (np.abs(xc)[None, :] < ohx) & (np.abs(xc)[:, None] < ohy)
    rho = np.broadcast_to(np.where(open_xy, np.float32(1e-8), np.float32(100.0)), (nz, nxy, nxy)).astype(np.float32)
    ap, k1 = O.make_grid(xe, xe, ze, rho.copy(), rot=rot, trans=trans)
    return [rs, ap], (k0, k1)


def test_mask_to_roi_matches_the_reference_roi_size(golden_dir):
    mt = G.f4_mask_total()
    start, stride, member = O.mask_to_roi(mt)
    m3 = member.reshape(NZ, NY, NX)
    # rows that begin inside the overlap of the two masks (sum 2) open their run at the first voxel with sum 1
    assert mt[110, 50, 35] == 2 and not m3[110, 50, 35:50].any() and m3[110, 50, 50:65].all()
    assert m3[60, 30, 35:65].all() and not m3[59].any()
    assert int(stride.sum()) == int(member.sum()) == 130 * 40 * 30 - 20 * 20 * 15
    assert np.all(np.diff(start.astype(np.int64)) > 0)


def test_oracle_roi_dose_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "f4_roi_release.npz"))
    meta = json.loads(str(g["meta"]))
    _, _, member = O.mask_to_roi(G.f4_mask_total())
    m3 = member.reshape(NZ, NY, NX).astype(bool)
    xe, ye, ze = edges()
    grid, keep = O.make_grid(xe, ye, ze, water())
    n = 60000
    b = O.make_beamlet(meta["energy"], [0, 0, 0.5, 0, 0, -1], [meta["spot_size"], meta["spot_size"], 0, 0, 0, 0], uniform=True)
    (d,), _ = O.transport(grid, O.VARIANT_RELEASE, [b], [n], seed=3, h0=0, n=n, kinds=[O.SCORER_DOSE], roi_members=[member])
    d = d.reshape(NZ, NY, NX) / n
    assert d[~m3].sum() == 0.0
    ref_xy, ref_idd = g["water_dE_total_xy"], g["water_dE_total_idd"]
    # the reference scores exactly the same voxel set: its projections vanish where the roi has no voxel
    assert ref_xy[~m3.any(axis=0)].sum() == 0.0 and ref_idd[~m3.any(axis=(1, 2))].sum() == 0.0
    assert np.all(ref_xy[m3.any(axis=0)] > 0)
    # ... including the rows shortened by the overlap quirk: slices 100-119 hold less than their neighbours
    assert ref_idd[100:120].mean() < 0.97 * 0.5 * (ref_idd[95:100].mean() + ref_idd[120:125].mean())
    assert abs(d.sum() / float(g["water_dE_total_total"]) - 1.0) < 0.01
    idd = d.sum(axis=(1, 2))
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99
    assert np.abs(idd - ref_idd).max() / ref_idd.max() < 0.03
    xy = d.sum(axis=0)
    # 5 x 5 mm blocks (about 600 histories each at 6e4 histories: 4 % noise, maximum over 50 blocks)
    a, r = rebin2(xy, 5), rebin2(ref_xy, 5)
    assert np.abs(a - r).max() / r.max() < 0.15


def test_oracle_beamline_dose_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "f3_beamline_release.npz"))
    meta = json.loads(str(g["meta"]))
    nodes, keep = f3_nodes()
    xe, ye, ze = edges()
    grid, k2 = O.make_grid(xe, ye, ze, water())
    n = 60000
    b = O.make_beamlet(meta["energy"], [0, 0, 180.0, 0, 0, -1], [meta["spot_size"], meta["spot_size"], 0, 0, 0, 0], uniform=True)
    (d,), st = O.transport(nodes + [grid], O.VARIANT_RELEASE, [b], [n], seed=8, h0=0, n=n, kinds=[O.SCORER_DOSE])
    d = d.reshape(NZ, NY, NX) / n
    ref_idd, ref_xy = g["water_dE_total_idd"], g["water_dE_total_xy"]
    assert abs(d.sum() / float(g["water_dE_total_total"]) - 1.0) < 0.015
    idd = d.sum(axis=(1, 2))
    assert abs(M.r80_mm(idd) - M.r80_mm(ref_idd)) < 0.15
    assert M.gamma_1d(ref_idd, idd, 1.0)[0] >= 0.99
    # the rotated, shifted aperture opening: same orientation and centre as the reference's
    xy = d.sum(axis=0)
    yy, xx = np.mgrid[0:NY, 0:NX] - 49.5

    def moments(a):
        w = a / a.sum()
        cx, cy = (w * xx).sum(), (w * yy).sum()
        sxx, syy, sxy = (w * (xx - cx) ** 2).sum(), (w * (yy - cy) ** 2).sum(), (w * (xx - cx) * (yy - cy)).sum()
        return cx, cy, 0.5 * np.degrees(np.arctan2(2 * sxy, sxx - syy)), sxx, syy

    mo, mr = moments(xy), moments(ref_xy)
    assert abs(mo[0] - mr[0]) < 0.3 and abs(mo[1] - mr[1]) < 0.3
    assert abs(mo[2] - mr[2]) < 3.0 and abs(mr[2] - G.F3_ANGLE_Z) < 5.0, (mo, mr)
    assert abs(mo[3] / mr[3] - 1) < 0.04 and abs(mo[4] / mr[4] - 1) < 0.04
    a, r = rebin2(xy, 4), rebin2(ref_xy, 4)
    assert np.abs(a - r).max() / r.max() < 0.15
