"""Host side of the treatment-planning front end (moquimc_b200/bin/tps_env, csrc/mqi_tps_host.hpp),
checked on CPU through `tps_env --dry-run`, which parses the input-parameter file, the .mha CT, the
generic PBS beam-model file and the text plan and prints the resulting beam source as JSON.

The expected values are a numpy restatement of the reference's formulas (file:line under
/root/reference/moqui) -- the reference's own tps_env cannot be built here (GDCM missing, SURVEY.md
section 8c), so this part of the boundary is pinned by restatement only:
  pbs::characterize_beamlet   base/mqi_treatment_machine_pbs.hpp:238-276 (fp32, intpl :94-96)
  pbs::characterize_history   :134-143
  beam_starting_position      base/mqi_treatment_machine_ion.hpp:324-333
  coordinate_transform        base/mqi_coordinate_transform.hpp:52-58, create_coordinate_transform tmi:42-70
  CT edges                    base/environments/mqi_tps_env.hpp:468-531
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moquimc_b200 import synthetic as S  # noqa: E402

EXE = os.path.join(ROOT, "moquimc_b200", "bin", "tps_env")
f32 = np.float32


def dry_run(inp, expect_fail=False):
    r = subprocess.run([EXE, "--dry-run", inp], capture_output=True, text=True, timeout=120)
    if expect_fail:
        assert r.returncode != 0, r.stdout
        return r.stderr
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DRYRUN ")][-1]
    return json.loads(line[len("DRYRUN "):])


def intpl(x, x0, x1, y0, y1):
    x, x0, x1, y0, y1 = map(f32, (x, x0, x1, y0, y1))
    return y0 if x1 == x0 else f32(y0 + f32(f32(f32(x - x0) * f32(y1 - y0)) / f32(x1 - x0)))


def expected_spot(rows, spot, sid, sad, pph):
    """rows: beam-model [spot] table; spot = (e, x, y, meterset)"""
    e, x, y, w = map(f32, spot)
    keys = [f32(r[0]) for r in rows]
    up = next(i for i, k in enumerate(keys) if k >= e)          # std::map::lower_bound
    up = max(up, 1)
    dn = up - 1
    D, U = [f32(v) for v in rows[dn]], [f32(v) for v in rows[up]]
    mid_e = intpl(e, D[0], U[0], D[1], U[1])                      # x axis = nominal energies
    mids = [intpl(e, D[1], U[1], D[c], U[c]) for c in (2, 3, 4, 5, 6)]   # x axis = down.E / up.E (as the reference)
    ratio = intpl(e, D[0], U[0], D[7], U[7])
    z = f32(sid)
    bx = f32(f32(x * f32(f32(sad[0]) - z)) / f32(sad[0]))
    by = f32(f32(y * f32(f32(sad[1]) - z)) / f32(sad[1]))
    d = np.array([x - bx, y - by, f32(0) - z], dtype=f32)
    d = d / f32(np.sqrt(f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))
    return {"energy": mid_e, "sigma_energy": mids[0], "mean": [bx, by, z, d[0], d[1], d[2]],
            "sigma": [mids[1], mids[2], 0, mids[3], mids[4], 0], "histories": int(f32(f32(w * ratio) / f32(pph)))}


def rot_matrix(collimator, gantry, couch):
    """iec2dicom(90 deg about x) * couch(about z, sign flipped) * gantry(about y) * collimator(about z),
    each mat3x3(a, b, c) = identity rotated about x, y, z (mqi_matrix.hpp:69-76, 214-273)"""
    def rx(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])

    def ry(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    def rz(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    d = np.pi / 180.0
    return rx(90 * d) @ rz(-couch * d) @ ry(gantry * d) @ rz(collimator * d)


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("tps"))
    inp = S.make_case(root, beams=2, ParticlesPerHistory=2.5e4, XShift=1.5, YShift=-2.0, ZShift=0.25)
    return root, inp


def test_beam_source_matches_reference_formulas(case):
    root, inp = case
    out = dry_run(inp)
    rows = S.beam_model_rows()
    assert [b["name"] for b in out["beams"]] == ["G000", "G090"]
    for bi, beam in enumerate(out["beams"]):
        spots = S.spot_list(n_layers=4, pitch=10.0, half_width=20.0, seed=1 + bi)
        assert beam["sid"] == 300.0 and len(beam["spots"]) == len(spots)
        R = rot_matrix(0.0, 90.0 * bi, 0.0)
        for got, sp in zip(beam["spots"], spots):
            sp = tuple(float("%.6g" % v) for v in sp[:3]) + (float("%.8g" % sp[3]),)   # as written to the plan file
            exp = expected_spot(rows, sp, 300.0, (2000.0, 1800.0), 2.5e4)
            assert got["histories"] == exp["histories"]
            np.testing.assert_allclose(got["energy"], exp["energy"], rtol=3e-7)
            np.testing.assert_allclose(got["sigma_energy"], exp["sigma_energy"], rtol=3e-6)
            np.testing.assert_allclose(got["mean"], exp["mean"], rtol=3e-6, atol=1e-6)
            np.testing.assert_allclose(got["sigma"], exp["sigma"], rtol=3e-6)
            np.testing.assert_allclose(np.array(got["rot"]).reshape(3, 3), R, atol=2e-7)
            assert got["trans"] == [0.0, 0.0, 0.0]


def test_ct_edges_follow_the_reference_rule(case):
    root, inp = case
    g = dry_run(inp)["grid"]
    assert g["n"] == [64, 64, 40]
    # origin = centre of voxel 0; first edge = centre - spacing / 2 + shift; last = first + n * spacing (fp32)
    x0 = f32(f32(-(64 - 1) / 2.0 * 4.0) - f32(4.0) / 2.0) + f32(1.5)
    y0 = f32(f32(-(64 - 1) / 2.0 * 4.0) - f32(4.0) / 2.0) + f32(-2.0)
    z0 = f32(f32(-(40 - 1) / 2.0 * 6.0) - f32(6.0) / 2.0) + f32(0.25)
    np.testing.assert_allclose(g["xe"], [x0, x0 + 64 * 4.0], rtol=1e-6)
    np.testing.assert_allclose(g["ye"], [y0, y0 + 64 * 4.0], rtol=1e-6)
    np.testing.assert_allclose(g["ze"], [z0, z0 + 40 * 6.0], rtol=1e-6)


def write_variant(root, name, **kw):
    p = os.path.join(root, name)
    S.write_input(p, root, os.path.join(root, "out_" + name), **kw)
    return p


def test_parser_semantics(case):
    root, _ = case
    # keys are case-insensitive, text after '#' is dropped, vectors split on commas
    p = os.path.join(root, "mixed.in")
    with open(p, "w") as f:
        f.write("# comment line\n\ngpuid 0\nRANDOMSEED 77 # trailing comment\nParentDir %s\nDicomDir .\nCTVolumeName ct.mha\n"
                "PlanFile plan.txt\nscorer Dose , LETd\nMachine pbs:machine.txt\nOutputDir %s\nOverwriteResults TRUE\n"
                "ParticlesPerHistory 1e5\nBeamNumbers 2\n" % (root, os.path.join(root, "o1")))
    out = dry_run(p)
    assert out["seed"] == 77 and out["n_fractions"] == 30 and [b["name"] for b in out["beams"]] == ["G090"]
    # Dij forces per-spot simulation and cannot be combined (mqi_tps_env.hpp:213-220)
    assert dry_run(write_variant(root, "dij.in", Scorer="Dij", UnitWeights=100))["sim_type"] == 1
    assert "Dij cannot be scored" in dry_run(write_variant(root, "dij2.in", Scorer="Dose,Dij"), expect_fail=True)
    assert "Unrecognized scorer name" in dry_run(write_variant(root, "bad.in", Scorer="Fluence"), expect_fail=True)
    assert "Valid machine is not available" in dry_run(write_variant(root, "m.in", Machine="gtr1:x"), expect_fail=True)
    assert "threshold cannot be zero" in dry_run(write_variant(root, "st.in", StoppingStatistics="true", StoppingCriteria=1.0),
                                                 expect_fail=True)
    assert "needs GDCM" in dry_run(write_variant(root, "rs.in", ReadStructure="true"), expect_fail=True)
    # UnitWeights overrides the histories of every spot in a per-spot run (:1488-1491)
    out = dry_run(write_variant(root, "uw.in", Scorer="Dij", UnitWeights=321))
    assert {s["histories"] for s in out["beams"][0]["spots"]} == {321}


def test_output_directory_guard(case):
    root, _ = case
    od = os.path.join(root, "exists")
    os.makedirs(od, exist_ok=True)
    p = os.path.join(root, "guard.in")
    S.write_input(p, root, od, OverwriteResults="false")
    assert "Output directory exists" in dry_run(p, expect_fail=True)


def test_npz_writer_is_a_scipy_csr(tmp_path):
    import scipy.sparse as sp
    path = str(tmp_path / "dij.npz")
    subprocess.run([EXE, "--npz-selftest", path], check=True, timeout=60)
    z = np.load(path)
    assert sorted(z.files) == ["data", "format", "indices", "indptr", "shape"]
    assert z["indices"].dtype == np.uint32 and z["indptr"].dtype == np.uint32 and z["shape"].dtype == np.uint32
    assert z["data"].dtype == np.float64 and z["format"].tobytes() == b"csr"
    assert z["shape"].tolist() == [3, 10] and z["indptr"].tolist() == [0, 2, 3, 6]
    # columns inside a row keep the table-slot order of the triplets, like the reference's single scan
    assert z["indices"].tolist() == [2, 0, 2, 7, 9, 5]
    m = sp.load_npz(path)
    dense = np.zeros((3, 10))
    for v, s, x in zip([7, 2, 9, 2, 0, 5], [2, 0, 2, 1, 0, 2], [0.5, 1.5, 2.5, 3.5, 4.5, 5.5]):
        dense[s, v] += x
    np.testing.assert_array_equal(m.toarray(), dense)


# ------------------------------------------------------------------------------------------------
# beamline children and mask regions of interest, host side
# ------------------------------------------------------------------------------------------------
def test_beamline_geometry_follows_the_reference_rules(tmp_path):
    """characterize_rangeshifter (pbs:279-331), characterize_aperture (:375-397), create_beamline sorting
    (tmi:238-276), create_voxelized_aperture / is_inside (mqi_tps_env.hpp:1651-1768)."""
    big = [(-30, -20), (30, -20), (30, 20), (-30, 20)]
    small = [(-10, -5), (20, -5), (20, 15), (-10, 15)]
    extra = {"rangeshifter_ids": ["RS1"], "blocks": [big, small], "block_thickness": 20.0, "block_tray_distance": 140.0}
    inp = S.make_case(str(tmp_path), beam_extra=extra)
    j = dry_run(inp)
    rs, ap = j["beams"][0]["beamline"]          # sorted by z, upstream first
    # model: rangeshifter(mm) 300 300, gap 10, "RS1" 40 mm; plan: snout 250
    assert rs["n"] == [1, 1, 1] and rs["xe"] == [-150, 150] and rs["ye"] == [-150, 150]
    assert rs["pos_z"] == 250 - (40 * 0.5 + 10) and rs["ze"] == [200, 240]
    assert abs(rs["rho0"] - 1.19e-3) < 1e-9     # RangeshifterDensity default 1.19 g/cm^3
    # aperture(mm) 300 300, thickness 20 at tray distance 140: centre 150, 1 mm voxels
    assert ap["n"] == [300, 300, 20] and ap["pos_z"] == 150 and ap["ze"] == [140, 160]
    assert ap["rho0"] == 100.0
    # with several blocks only the LAST polygon counts (is_inside overwrites `inside`): 30 x 20 mm opening
    assert ap["open_voxels"] == 30 * 20 * 20
    assert ap["open_centroid"] == [5.0, 5.0]
    # range shifter by water-equivalent thickness when the model lists no IDs is covered by the formula only:
    # thickness = WET / 1.15, position = distance - thickness
    model = os.path.join(str(tmp_path), "machine.txt")
    txt = open(model).read().replace('"RS1" 40.0\n', "")
    open(model, "w").write(txt)
    extra2 = {"rangeshifter_wet": (46.0, 230.0)}
    S.write_plan(os.path.join(str(tmp_path), "plan.txt"),
                 [{"name": "G000", "spots": S.spot_list(n_layers=2, pitch=10.0, half_width=10.0), **extra2}])
    j2 = dry_run(inp)
    (rs2,) = j2["beams"][0]["beamline"]
    lz = f32(46.0) / f32(1.15)
    assert abs(rs2["ze"][1] - rs2["ze"][0] - float(lz)) < 1e-4 and abs(rs2["pos_z"] - float(f32(230.0) - lz)) < 1e-4


def test_mask_keys_and_roi_sizes(tmp_path):
    root = str(tmp_path)
    n = (64, 64, 40)
    m1 = np.zeros((n[2], n[1], n[0]), dtype=np.uint8)
    m1[10:30, 20:40, 20:44] = 1
    m2 = np.zeros_like(m1)
    m2[12:14, 25:30, 20:30] = 1                 # overlaps m1 from the start of its rows
    S.write_mask_mha(os.path.join(root, "m1.mha"), m1)
    S.write_mask_mha(os.path.join(root, "m2.mha"), m2)
    inp = S.make_case(root, n=n, ScoringMask="true", Mask="%s,%s" % (os.path.join(root, "m1.mha"), os.path.join(root, "m2.mha")),
                      StoppingStatistics="true", StoppingCriteria="2.0", StatThreshold="0.0",
                      StatROIMaskFilename=os.path.join(root, "m1.mha"))
    j = dry_run(inp)
    # rows where the sum starts at 2 open their run only where it drops to 1: 10 voxels per such row are lost
    assert j["scoring_roi_size"] == 20 * 20 * 24 - 2 * 5 * 10
    assert j["stat_roi_size"] == 20 * 20 * 24
    # StatThreshold 0 without a stat roi is an error, ScoringMask without Mask too, RTSTRUCT options are refused
    err = dry_run(S.make_case(root, n=n, StoppingStatistics="true", StoppingCriteria="2.0", StatThreshold="0.0"), expect_fail=True)
    assert "threshold" in err
    err = dry_run(S.make_case(root, n=n, ScoringMask="true"), expect_fail=True)
    assert "Mask filename is missing" in err
    err = dry_run(S.make_case(root, n=n, ReadStructure="true"), expect_fail=True)
    assert "StructureFile" in err
    # a mask with other dimensions is refused
    S.write_mask_mha(os.path.join(root, "bad.mha"), np.zeros((4, 4, 4), dtype=np.uint8))
    err = dry_run(S.make_case(root, n=n, ScoringMask="true", Mask=os.path.join(root, "bad.mha")), expect_fail=True)
    assert "dimensions" in err


def rasterize_numpy(contours, xe, ye, ze, dx, dy):
    """fill_contour + sol1_1 (mqi_tps_env.hpp:1769-1826) restated with numpy in fp32."""
    nx, ny, nz = len(xe) - 1, len(ye) - 1, len(ze) - 1
    vol = np.zeros((nz, ny, nx), dtype=np.uint8)
    px = (xe[:nx - 1] + f32(dx) * f32(0.5)).astype(f32)   # the reference promotes dx * 0.5 to double; identical here
    py = (ye[:ny - 1] + f32(dy) * f32(0.5)).astype(f32)
    for c in contours:
        c = np.asarray(c, dtype=f32)
        z = c[0, 2]
        k = next((i for i in range(nz - 1) if ze[i] < z < ze[i + 1]), -1)
        if k < 0:
            continue
        inside = np.zeros((ny - 1, nx - 1), dtype=bool)
        n = len(c)
        for i in range(n):
            j = (i - 1) % n
            x0, y0, x1, y1 = c[i, 0], c[i, 1], c[j, 0], c[j, 1]
            cond_y = ((y0 <= py) & (py < y1)) | ((y1 <= py) & (py < y0))
            with np.errstate(divide="ignore", invalid="ignore"):
                xi = ((x1 - x0) * (py - y0) / (y1 - y0) + x0).astype(f32)
            hit = cond_y[:, None] & (px[None, :] < xi[:, None])
            inside ^= hit
        vol[k, :ny - 1, :nx - 1] |= inside.astype(np.uint8)
    return vol


def test_structure_contours_are_rasterised_like_fill_contour(tmp_path):
    root = str(tmp_path)
    n, sp = (64, 64, 40), (4.0, 4.0, 6.0)
    zc = (np.arange(n[2]) - (n[2] - 1) / 2.0) * sp[2]
    body = S.ellipse_contours(70.0, 90.0, zc[3:37], centre=(4.0, -6.0))
    ptv = S.ellipse_contours(22.0, 18.0, zc[14:26], n_points=24, centre=(-8.0, 10.0))
    # a contour in the LAST slab is never found (the slice search stops at nz - 1): kept
    body_plus = body + S.ellipse_contours(70.0, 90.0, [zc[-1]], centre=(4.0, -6.0))
    S.write_structures(os.path.join(root, "structures.txt"), {"External": body_plus, "PTV 1": ptv})
    inp = S.make_case(root, n=n, spacing=sp, ReadStructure="true", BodyContourName="external", StructureFile="structures.txt",
                      StoppingStatistics="true", StoppingCriteria="3.0", StatROIStructFromRT="true", StatROI="PTV 1")
    env = dict(os.environ, MQI_DRYRUN_DUMP_MASKS=root)
    r = subprocess.run([EXE, "--dry-run", inp], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stderr
    j = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("DRYRUN ")][-1][len("DRYRUN "):])
    g = j["grid"]
    xe = (f32(g["xe"][0]) + np.arange(n[0] + 1, dtype=f32) * f32(sp[0])).astype(f32)
    ye = (f32(g["ye"][0]) + np.arange(n[1] + 1, dtype=f32) * f32(sp[1])).astype(f32)
    ze = np.empty(n[2] + 1, dtype=f32)
    ze[0] = g["ze"][0]
    for i in range(1, n[2] + 1):
        ze[i] = ze[i - 1] + f32(sp[2])
    for name, contours, key in (("scoring_mask.raw", body_plus, "scoring_roi_size"), ("stat_mask.raw", ptv, "stat_roi_size")):
        got = np.fromfile(os.path.join(root, name), dtype=np.uint8).reshape(n[2], n[1], n[0])
        want = rasterize_numpy(contours, xe, ye, ze, sp[0], sp[1])
        assert want.sum() > 0 and np.array_equal(got, want), name
        assert j[key] == int(want.sum())      # contours are 0/1: every run of the mask is inside the roi
        assert got[-1].sum() == 0 and got[:, -1, :].sum() == 0 and got[:, :, -1].sum() == 0
    # unknown structure names and a missing structure file are errors
    err = dry_run(S.make_case(root, n=n, spacing=sp, ReadStructure="true", BodyContourName="Skin", StructureFile="structures.txt"), expect_fail=True)
    assert "Skin" in err
    err = dry_run(S.make_case(root, n=n, spacing=sp, ReadStructure="true"), expect_fail=True)
    assert "StructureFile" in err


def roi_python(mask):
    """mask_reader::mask_to_roi + roi_t::get_contour_idx restated in plain Python (mqi_file_handler.hpp:176-217,
    mqi_roi.hpp:88-100), with the documented fix: a run still open at the end of the volume is closed there."""
    start, stride, acc = [], [], []
    open_, s0 = False, 0
    for i, v in enumerate(mask):
        if v == 1 and not open_:
            open_, s0 = True, i
        if v == 0 and open_:
            open_ = False
            start.append(s0)
            stride.append(i - s0)
            acc.append((acc[-1] if acc else 0) + i - s0)
    if open_:
        start.append(s0)
        stride.append(len(mask) - s0)
        acc.append((acc[-1] if acc else 0) + len(mask) - s0)

    def idx(v):
        c = sum(1 for s in start if s <= v) - 1     # lower_bound_cpp(v) - 1
        if c < 0:
            return -1
        d = v - start[c]
        if d < stride[c]:
            return d + (acc[c - 1] if c > 0 else 0)
        return -1
    return start, stride, acc, idx


@pytest.mark.parametrize("case", ["random", "zeros", "ones", "open_end", "overlap", "single"])
def test_run_length_roi_matches_the_reference_rule(tmp_path, case):
    rng = np.random.default_rng(7)
    n = 400
    if case == "random":
        m = (rng.random(n) < 0.5).astype(np.uint8)
    elif case == "zeros":
        m = np.zeros(n, np.uint8)
    elif case == "ones":
        m = np.ones(n, np.uint8)
    elif case == "open_end":
        m = np.zeros(n, np.uint8); m[350:] = 1; m[10:20] = 1
    elif case == "overlap":   # sums of two masks: 2 neither opens nor closes a run
        m = (rng.random(n) < 0.4).astype(np.uint8) + (rng.random(n) < 0.4).astype(np.uint8)
    else:
        m = np.zeros(n, np.uint8); m[0] = 1
    path = os.path.join(str(tmp_path), "m.raw")
    m.tofile(path)
    r = subprocess.run([EXE, "--roi-selftest", path], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    start, stride, acc, idx = roi_python(m.tolist())
    runs = [tuple(int(x) for x in ln.split()[1:]) for ln in lines if ln.startswith("run ")]
    assert runs == list(zip(start, stride, acc))
    assert lines[0] == "runs %d size %d" % (len(start), acc[-1] if acc else 0)
    got = {int(ln.split()[1]): int(ln.split()[2]) for ln in lines if ln.startswith("idx ")}
    assert got == {v: idx(v) for v in range(0, n, 7)}
    assert int([ln for ln in lines if ln.startswith("bits ")][0].split()[1]) == sum(stride)
    # the oracle's restatement agrees (it is the checker of the device bitmask in the GPU suite)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    s2, t2, member = O.mask_to_roi(m)
    assert list(s2) == start and list(t2) == stride
    exp = np.zeros(n, np.uint8)
    for s, t in zip(start, stride):
        exp[s:s + t] = 1
    assert np.array_equal(member, exp)


def test_beam_frame_matches_the_reference_coordinate_transform(tmp_path, golden_dir):
    """The beam frame of every beam -- collimator, gantry, couch (negated) and iec2dicom = 90 degrees composed by
    coordinate_transform (mqi_coordinate_transform.hpp:52-58, create_coordinate_transform tmi:42-70) -- against
    matrices the reference header itself produced (oracle/ref_kat.cpp section 8 -> tests/golden/kat_release.npz).
    The numpy restatement used by the other tests of this module is held to the same vectors."""
    k = np.load(os.path.join(golden_dir, "kat_release.npz"))
    ang, rot, moved = k["ct_ang"].reshape(-1, 4), k["ct_rot"].reshape(-1, 3, 3), k["ct_moved"].reshape(-1, 3)
    sel = [i for i in range(len(ang)) if ang[i, 3] == 90.0]
    assert len(sel) == 96
    for i in sel:   # the restatement: rot_matrix takes the couch angle as the plan states it
        np.testing.assert_allclose(rot_matrix(ang[i, 0], ang[i, 1], -ang[i, 2]), rot[i], atol=3e-7)
    # the C++ host code, through the text plan (a subset keeps the plan small: every gantry angle, mixed collimator / couch)
    pick = [i for i in sel if (i // 2) % 5 == 0][:20]
    root = str(tmp_path)
    os.makedirs(root, exist_ok=True)
    hu, origin = S.head_ct((32, 32, 20), (4.0, 4.0, 6.0), 1)
    S.write_mha(os.path.join(root, "ct.mha"), hu, origin, (4.0, 4.0, 6.0))
    S.write_beam_model(os.path.join(root, "machine.txt"))
    iso = (1.5, -2.5, 40.0)                                   # the translation of the KAT
    beams = [{"name": "B%02d" % n, "collimator": float(ang[i, 0]), "gantry": float(ang[i, 1]), "couch": float(-ang[i, 2]),
              "iso": iso, "snout": 250.0, "spots": S.spot_list(n_layers=1, pitch=20.0, half_width=10.0, seed=3)}
             for n, i in enumerate(pick)]
    S.write_plan(os.path.join(root, "plan.txt"), beams)
    inp = os.path.join(root, "moqui_tps.in")
    S.write_input(inp, root, os.path.join(root, "out"), ParticlesPerHistory=1e4)
    out = dry_run(inp)
    assert len(out["beams"]) == len(pick)
    for beam, i in zip(out["beams"], pick):
        for sp in beam["spots"]:
            np.testing.assert_allclose(np.array(sp["rot"]).reshape(3, 3), rot[i], atol=3e-7)
            np.testing.assert_allclose(sp["trans"], iso, atol=1e-6)
    # and one point carried through a frame by hand: R * p + T as the reference computed it
    p = np.array([3.0, -4.0, 465.0])
    for i in sel[:8]:
        np.testing.assert_allclose(rot[i].astype(np.float64) @ p + np.array(iso), moved[i], rtol=2e-6, atol=2e-5)


def test_output_writers_match_the_files_the_reference_writes(tmp_path, golden_dir):
    """raw / mhd / mha byte for byte, npz member by member, against files written by the reference's own
    io::save_to_mhd, save_to_mha and save_to_npz (mqi_io.hpp:249-320, 493-591) for the same small volume and the same
    six-entry (voxel, spot) table: oracle/ref_kat.cpp sections 9-10 -> tests/golden/fmt_writers.npz."""
    import io
    import zipfile
    g = np.load(os.path.join(golden_dir, "fmt_writers.npz"))
    d = tmp_path / "w"
    d.mkdir()
    phantom_env = os.path.join(os.path.dirname(EXE), "phantom_env")
    subprocess.run([phantom_env, "--write-selftest", str(d)], check=True, timeout=60, stdout=subprocess.DEVNULL)
    for key, name in (("fmt_mhd_mhd", "fmt_mhd.mhd"), ("fmt_mhd_raw", "fmt_mhd.raw"), ("fmt_mha_mha", "fmt_mha.mha")):
        assert (d / name).read_bytes() == bytes(g[key]), name
    path = str(tmp_path / "dij.npz")
    subprocess.run([EXE, "--npz-selftest", path], check=True, timeout=60)
    ref = zipfile.ZipFile(io.BytesIO(bytes(g["fmt_npz_npz"])))
    mine = zipfile.ZipFile(path)
    assert mine.namelist() == ref.namelist() == ["indices.npy", "indptr.npy", "shape.npy", "data.npy", "format.npy"]
    for n in ("indices.npy", "indptr.npy", "shape.npy", "data.npy"):
        a, b = np.load(io.BytesIO(mine.read(n))), np.load(io.BytesIO(ref.read(n)))
        assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), n
    # format.npy: the reference declares a 0-d '|S3' array and writes sizeof(std::string) * 3 = 96 bytes for it -- "csr"
    # followed by whatever lies behind the string object (quirk B12, mqi_sparse_io.hpp:223-225); here the member is
    # the three characters and nothing else
    mf, rf = mine.read("format.npy"), ref.read("format.npy")
    assert np.load(io.BytesIO(mf)).tobytes() == b"csr" and np.load(io.BytesIO(rf)).tobytes() == b"csr"
    assert len(rf) > len(mf) == mf.index(b"csr") + 3


def test_input_parameter_format_is_parsed_like_the_reference_parser(tmp_path, golden_dir):
    """The moqui input-parameter file: the reference's own file_parser (mqi_file_handler.hpp:220-380, compiled from a
    scratch cut of that header, oracle/build_ref.sh patch 4) answered 42 queries on a file with comments, tabs, odd
    spacing, mixed-case and repeated keys, empty values and lists (oracle/ref_kat.cpp section 11 ->
    tests/golden/fmt_writers.npz); this file_parser has to give the same answers, quirks included (a tab is not a
    delimiter, the first of two keys that differ only in case wins, atoi / atof stop at the first odd character)."""
    g = np.load(os.path.join(golden_dir, "fmt_writers.npz"))
    inp = tmp_path / "in.txt"
    inp.write_bytes(bytes(g["fmt_parser_in_txt"]))
    r = subprocess.run([EXE, "--parse-selftest", str(inp)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    ref = bytes(g["fmt_parser_out_txt"]).decode()
    assert len(ref.splitlines()) == 42
    assert r.stdout.splitlines() == ref.splitlines()


def test_mask_files_are_read_like_the_reference_mask_reader(tmp_path, golden_dir):
    """ScoringMask / StatROIMaskFilename inputs: the reference's own mask_reader (read_mha_file, read_mask_files,
    mask_to_roi; mqi_file_handler.hpp:38-217) read two overlapping uint8 .mha masks -- one with an ITK-ordered header, one
    with upper-case keys and odd spacing -- in oracle/ref_kat.cpp section 12; the same files through this reader
    (tps_env --mask-selftest) give the same summed mask and the same run-length roi.  The one designed difference: a
    run still open at the end of the volume is closed there (the reference leaves that stride uninitialised)."""
    g = np.load(os.path.join(golden_dir, "fmt_writers.npz"))
    k = np.load(os.path.join(golden_dir, "kat_release.npz"))
    a, b = tmp_path / "a.mha", tmp_path / "b.mha"
    a.write_bytes(bytes(g["fmt_mask_a_mha"]))
    b.write_bytes(bytes(g["fmt_mask_b_mha"]))
    r = subprocess.run([EXE, "--mask-selftest", "%s,%s" % (a, b), "7", "5", "4"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    total = [int(x) for x in [ln for ln in lines if ln.startswith("total")][0].split()[1:]]
    assert total == k["mask_total"].tolist() and max(total) == 2       # overlaps add up
    runs = [tuple(int(x) for x in ln.split()[1:3]) for ln in lines if ln.startswith("run ")]
    ref = k["mask_runs"]
    n = int(ref[0])
    ref_runs = [(int(ref[1 + 2 * i]), int(ref[2 + 2 * i])) for i in range(n)]
    assert len(runs) == n == 29
    assert runs[:-1] == ref_runs[:-1] and runs[-1][0] == ref_runs[-1][0]
    assert runs[-1][0] + runs[-1][1] == 7 * 5 * 4                      # closed at the end of the volume
    assert total[0] == 2 and runs[0][0] != 0                           # a voxel where both masks start does not open a run
    # a mask of the wrong size is refused (the reference would read past its buffer)
    bad = subprocess.run([EXE, "--mask-selftest", str(a), "7", "5", "5"], capture_output=True, text=True, timeout=60)
    assert bad.returncode != 0


def test_beam_source_matches_the_reference_beam_model_code(tmp_path, golden_dir):
    """The treatment_machines beam model at work: the reference's own mqi::pbs (beam data file, spot -> beamlet
    interpolation, histories per spot), treatment_machine_ion::create_beamsource / create_coordinate_transform and
    beam_module_ion ran UNMODIFIED on the synthetic machine file and a 52-spot plan (oracle/ref_tps_kat.cpp; only the
    GDCM-backed dataset class below them is an in-memory stand-in) -> tests/golden/b1_beam_model.npz.  The C++ host code
    of this framework, fed the same machine file and the same plan as text, has to build the same beamlets: energy and
    its spread, start position 50 mm upstream of the snout, direction through the SAD, spot sizes and divergences,
    histories per spot, the couch angle negated, the isocentre as translation."""
    import ast
    g = np.load(os.path.join(golden_dir, "b1_beam_model.npz"))
    a = ast.literal_eval(str(g["meta"]))
    lines = bytes(g["text"]).decode().splitlines()
    ref_angles = [float(x) for x in lines[0].split()[1:]]
    ref_trans = [float(x) for x in lines[1].split()[1:]]
    ref_spots = [[float(x) for x in ln.split()[2:]] for ln in lines[2:]]
    root = str(tmp_path)
    hu, origin = S.head_ct((32, 32, 20), (4.0, 4.0, 6.0), 1)
    S.write_mha(os.path.join(root, "ct.mha"), hu, origin, (4.0, 4.0, 6.0))
    S.write_beam_model(os.path.join(root, "machine.txt"))
    spots = S.spot_list(n_layers=a["n_layers"], pitch=a["pitch"], half_width=a["half_width"], seed=a["seed"])
    S.write_plan(os.path.join(root, "plan.txt"), [{"name": "B0", "collimator": a["collimator"], "gantry": a["gantry"],
                                                     "couch": a["couch"], "iso": a["iso"], "snout": a["snout"], "spots": spots}])
    inp = os.path.join(root, "moqui_tps.in")
    S.write_input(inp, root, os.path.join(root, "out"), ParticlesPerHistory=a["pph"])
    beam = dry_run(inp)["beams"][0]
    assert beam["sid"] == a["snout"] + 50.0 and len(beam["spots"]) == len(ref_spots) == 52
    assert ref_angles == [a["collimator"], a["gantry"], -a["couch"], 0.0] and ref_trans == list(a["iso"])
    for sp, ref in zip(beam["spots"], ref_spots):
        assert sp["histories"] == int(ref[0])
        mine = [sp["energy"], sp["sigma_energy"]] + sp["mean"] + sp["sigma"]
        np.testing.assert_allclose(mine, ref[1:], rtol=2e-7, atol=1e-9)
        np.testing.assert_allclose(sp["trans"], ref_trans, atol=1e-6)
        np.testing.assert_allclose(np.array(sp["rot"]).reshape(3, 3), rot_matrix(a["collimator"], a["gantry"], a["couch"]), atol=3e-7)


def test_beamline_devices_match_the_reference_create_beamline(tmp_path, golden_dir):
    """Range shifter and aperture of a beam as the reference's own create_beamline / characterize_rangeshifter /
    characterize_aperture (tmi:238-276, pbs:279-397) built them, unmodified, from a plan held in memory
    (oracle/ref_tps_kat.cpp -> tests/golden/b1_beam_model.npz geo_*): thickness from the RangeShifterID table (one and
    two IDs) or from the water-equivalent thickness / 1.15 when the machine file has no table, position from the
    snout position and gap or from the isocentre distance, block thickness and tray distance, devices sorted upstream
    first.  tps_env builds its beamline nodes from the same quantities in the text plan."""
    g = np.load(os.path.join(golden_dir, "b1_beam_model.npz"))

    def ref(key):
        return [(ln.split()[1], [float(x) for x in ln.split()[2:]]) for ln in bytes(g["geo_" + key]).decode().splitlines()]

    block = [(-10.0, -8.0), (10.0, -8.0), (10.0, 8.0), (-10.0, 8.0)]
    cases = {"rs_id_block": dict(rangeshifter_ids=["RS1"], blocks=[block], block_thickness=20.0, block_tray_distance=60.0),
             "rs_two_ids": dict(rangeshifter_ids=["RS1", "RS1"]),
             "rs_wet": dict(rangeshifter_wet=(46.0, 120.0))}
    for key, extra in cases.items():
        root = str(tmp_path / key)
        os.makedirs(root)
        hu, origin = S.head_ct((32, 32, 20), (4.0, 4.0, 6.0), 1)
        S.write_mha(os.path.join(root, "ct.mha"), hu, origin, (4.0, 4.0, 6.0))
        S.write_beam_model(os.path.join(root, "machine.txt"))
        if key == "rs_wet":   # no [rangeshifter_thickness] table: the thickness comes from the plan
            text = open(os.path.join(root, "machine.txt")).read()
            i, j = text.index("[rangeshifter_thickness]"), text.index("[spot]")
            open(os.path.join(root, "machine.txt"), "w").write(text[:i] + text[j:])
        beam = {"name": "B0", "gantry": 45.0, "couch": 10.0, "collimator": 15.0, "iso": (1.5, -2.5, 40.0), "snout": 250.0,
                "spots": S.spot_list(n_layers=1, pitch=20.0, half_width=10.0, seed=3)}
        beam.update(extra)
        S.write_plan(os.path.join(root, "plan.txt"), [beam])
        inp = os.path.join(root, "moqui_tps.in")
        S.write_input(inp, root, os.path.join(root, "out"), ParticlesPerHistory=1e4)
        nodes = dry_run(inp)["beams"][0]["beamline"]
        want = ref(key)
        assert len(nodes) == len(want), key
        for n, (kind, v) in zip(nodes, want):          # same order: upstream first
            lx, ly, lz, px, py, pz = v
            np.testing.assert_allclose(n["pos_z"], pz, rtol=1e-6)
            np.testing.assert_allclose(n["ze"], [pz - lz / 2, pz + lz / 2], rtol=1e-6)
            np.testing.assert_allclose(n["xe"], [-lx / 2, lx / 2], rtol=1e-6)
            np.testing.assert_allclose(n["ye"], [-ly / 2, ly / 2], rtol=1e-6)
            if kind == "block":
                assert n["n"][2] == int(lz) and n["open_voxels"] == 20 * 16 * int(lz)   # 1 mm voxels, 20 x 16 mm opening
            else:
                assert n["n"] == [1, 1, 1]


def test_ct_mha_element_types(tmp_path):
    """MET_FLOAT (what the reference reads, mqi_tps_env.hpp:615-701) and MET_SHORT give the same HU volume; another
    element type, a compressed or a big-endian payload is refused instead of being misread as float."""
    root = str(tmp_path)
    inp = S.make_case(root, n=(24, 20, 10), spacing=(4.0, 4.0, 6.0), n_layers=1)
    hu, origin = S.head_ct((24, 20, 10), (4.0, 4.0, 6.0), 1)
    a = dry_run(inp)["hu"]
    assert a == {"sum": int(hu.astype(np.int64).sum()), "min": int(hu.min()), "max": int(hu.max())}
    S.write_mha(os.path.join(root, "ct.mha"), hu, origin, (4.0, 4.0, 6.0), element="MET_SHORT")
    assert dry_run(inp)["hu"] == a
    S.write_mha(os.path.join(root, "ct.mha"), (hu % 200).astype(np.uint8), origin, (4.0, 4.0, 6.0), element="MET_UCHAR")
    assert "unsupported ElementType" in dry_run(inp, expect_fail=True)
    # out-of-range floats are clamped, not wrapped
    big = hu.astype(np.float32)
    big[0, 0, 0], big[0, 0, 1] = 1.0e9, -1.0e9
    S.write_mha(os.path.join(root, "ct.mha"), big, origin, (4.0, 4.0, 6.0))
    h = dry_run(inp)["hu"]
    assert h["max"] == 32767 and h["min"] == -32768
    text = open(os.path.join(root, "ct.mha"), "rb").read().replace(b"CompressedData = False", b"CompressedData = True")
    open(os.path.join(root, "ct.mha"), "wb").write(text)
    assert "compressed" in dry_run(inp, expect_fail=True)
