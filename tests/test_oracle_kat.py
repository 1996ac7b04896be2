"""Pin the CPU restatement (oracle/mqi_oracle.c) against vectors dumped from the reference's own
headers (oracle/ref_kat.cpp -> tests/golden/kat_<variant>.npz).  Bit-exact unless stated."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

VARIANTS = [("debug", O.VARIANT_DEBUG), ("release", O.VARIANT_RELEASE)]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, "kat_%s.npz" % name))


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def c1_grid(k):
    e = k["geo_edges"]
    xe, ye, ze = e[:201], e[201:402], e[402:]
    rho = np.zeros(200 * 200 * 350, dtype=np.float32)
    return O.make_grid(xe, ye, ze, rho)


def test_tables_match_reference_checksum(golden_dir):
    k = load(golden_dir, "debug")
    raw = np.fromfile(O.TABLES, dtype=np.uint8)
    t = raw[16:16 + 3600 * 4].view(np.float32)
    c = raw[16 + 3600 * 4:].view(np.float32)
    assert t.astype(np.float64).sum() == float(k["tables_sum"])
    assert c.astype(np.float64).sum() == float(k["density_correction_sum"])


def test_edges_rule_matches_reference(golden_dir):
    k = load(golden_dir, "debug")
    e = k["geo_edges"]
    assert (bits(O.uniform_edges(-50, 50, 200)) == bits(e[:201])).all()
    assert (bits(O.uniform_edges(-350, 0, 350)) == bits(e[402:])).all()


@pytest.mark.parametrize("name,variant", VARIANTS)
def test_hu_to_density_bit_exact(golden_dir, name, variant):
    k = load(golden_dir, name)
    got = O.hu_to_density(k["hu"])
    assert (bits(got) == bits(k["hu_rho"])).all()
    # survey-time known answers
    assert got[list(k["hu"]).index(0)] == np.float32(9.88223474e-04)
    assert got[list(k["hu"]).index(-1000)] == np.float32(1.13160659e-05)
    assert got[list(k["hu"]).index(3000)] == np.float32(4.55248496e-03)


@pytest.mark.parametrize("name,variant", VARIANTS)
def test_rsp_and_radiation_length_bit_exact(golden_dir, name, variant):
    k = load(golden_dir, name)
    L = O.lib()
    got = np.array([L.mqo_spr(float(r), float(e), variant) for r, e in zip(k["rsp_rho"], k["rsp_ek"])],
                   dtype=np.float32)
    assert (bits(got) == bits(k["rsp_out"])).all()
    got = np.array([L.mqo_radiation_length(float(r), variant) for r in k["rl_rho"]], dtype=np.float32)
    assert (bits(got) == bits(k["rl_out"])).all()


def test_hash_keys_bit_exact(golden_dir):
    k = load(golden_dir, "debug")
    L = O.lib()
    got = np.array([L.mqo_hash(int(a), int(b), int(c)) for a, b, c in
                    zip(k["hash_k1"], k["hash_k2"], k["hash_cap"])], dtype=np.uint32)
    assert (got == k["hash_out"]).all()
    assert L.mqo_hash(0, 0, 1000003) == 687275
    assert L.mqo_hash(123456, 7, 1000003) == 252790
    assert L.mqo_hash(13999999, 4999, 393216000) == 281020009


def test_start_and_length(golden_dir):
    k = load(golden_dir, "debug")
    L = O.lib()
    i = k["sal_in"].reshape(-1, 3)
    o = k["sal_out"].reshape(-1, 2)
    for (a, b, c), exp in zip(i, o):
        out = (C.c_uint32 * 2)()
        L.mqo_start_and_length(int(a), int(b), int(c), out)
        assert (out[0], out[1]) == tuple(exp)


def test_physics_tables_and_relativistic_quantities_bit_exact(golden_dir):
    k = load(golden_dir, "debug")
    L = O.lib()
    exp = k["phys_out"].reshape(-1, 9)
    out = (C.c_float * 9)()
    got = np.empty_like(exp)
    for i, e in enumerate(k["phys_ek"]):
        L.mqo_physics_probe(C.c_float(float(e)), out)
        got[i] = np.frombuffer(out, dtype=np.float32)
    assert (bits(got) == bits(exp)).all()


def test_grid_index_intersect_bit_exact(golden_dir):
    k = load(golden_dir, "debug")
    L = O.lib()
    g, keep = c1_grid(k)
    P = k["geo_p"].reshape(-1, 3)
    D = k["geo_d"].reshape(-1, 3)
    IDX = k["geo_idx"].reshape(-1, 3)
    IDX1 = k["geo_idx1"].reshape(-1, 3)
    DA = k["geo_dir_after"].reshape(-1, 3)
    P1 = k["geo_p1"].reshape(-1, 3)
    n_valid = 0
    for n in range(len(P)):
        p = (C.c_float * 3)(*P[n])
        d = (C.c_float * 3)(*D[n])
        cell = (C.c_int * 3)()
        L.mqo_grid_index(C.byref(g), p, d, cell)
        assert tuple(cell) == tuple(IDX[n]), n
        valid = all(0 <= cell[a] < (200, 200, 350)[a] for a in range(3))
        if not valid:
            assert k["geo_cnb"][n] == np.uint64(0xFFFFFFFFFFFFFFFF)
            continue
        n_valid += 1
        assert cell[2] * 40000 + cell[1] * 200 + cell[0] == int(k["geo_cnb"][n])
        dist = L.mqo_grid_intersect_cell(C.byref(g), p, d, cell)
        assert bits([dist])[0] == bits([k["geo_dist"][n]])[0], n
        assert (bits(np.frombuffer(d, dtype=np.float32)) == bits(DA[n])).all()
        p1 = (C.c_float * 3)(*P1[n])
        c1 = (C.c_int * 3)(*cell)
        L.mqo_grid_index_update(C.byref(g), p1, d, c1)
        assert tuple(c1) == tuple(IDX1[n]), n
    assert n_valid > 4000


def test_grid_entry_bit_exact(golden_dir):
    k = load(golden_dir, "debug")
    L = O.lib()
    g, keep = c1_grid(k)
    P = k["entry_p"].reshape(-1, 3)
    D = k["entry_d"].reshape(-1, 3)
    CELL = k["entry_cell"].reshape(-1, 3)
    hits = 0
    for n in range(len(P)):
        p = (C.c_float * 3)(*P[n])
        d = (C.c_float * 3)(*D[n])
        cell = (C.c_int * 3)()
        dist = L.mqo_grid_intersect_entry(C.byref(g), p, d, cell)
        assert bits([dist])[0] == bits([k["entry_dist"][n]])[0], n
        assert tuple(cell) == tuple(CELL[n]), n
        hits += dist >= 0
    assert hits > 500
    # entry from (0.1,-0.2,0.5) heading -z: dist 0.5, cell (100,99,349)
    p = (C.c_float * 3)(0.1, -0.2, 0.5)
    d = (C.c_float * 3)(0, 0, -1)
    cell = (C.c_int * 3)()
    assert L.mqo_grid_intersect_entry(C.byref(g), p, d, cell) == 0.5
    assert tuple(cell) == (100, 99, 349)


def test_scattering_rotation_bit_exact(golden_dir):
    k = load(golden_dir, "debug")
    L = O.lib()
    D = k["rot_d"].reshape(-1, 3)
    A = k["rot_ang"].reshape(-1, 2)
    E = k["rot_out"].reshape(-1, 3)
    out = (C.c_float * 3)()
    for n in range(len(D)):
        L.mqo_rotate_direction((C.c_float * 3)(*D[n]), C.c_float(float(A[n][0])), C.c_float(float(A[n][1])), out)
        got = np.frombuffer(out, dtype=np.float32)
        same = (bits(got) == bits(E[n])).all() or (np.isnan(got).all() and np.isnan(E[n]).all())
        assert same, (n, got, E[n])


def test_philox_known_answers():
    # Random123 known-answer vectors for Philox4x32-10
    L = O.lib()
    out = (C.c_uint32 * 4)()
    L.mqo_philox4x32_10((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    L.mqo_philox4x32_10((C.c_uint32 * 4)(0xffffffff,) * 1, (C.c_uint32 * 2)(0xffffffff, 0xffffffff), out) if False else None
    L.mqo_philox4x32_10((C.c_uint32 * 4)(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff),
                        (C.c_uint32 * 2)(0xffffffff, 0xffffffff), out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    L.mqo_philox4x32_10((C.c_uint32 * 4)(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
                        (C.c_uint32 * 2)(0xa4093822, 0x299f31d0), out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox2x32_known_answers():
    # Random123 kat_vectors for Philox2x32-10 (the delta-electron sampler's generator)
    L = O.lib()
    out = (C.c_uint32 * 2)()
    for ctr, key, exp in (((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
                          ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
                          ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))):
        L.mqo_philox2x32_10((C.c_uint32 * 2)(*ctr), C.c_uint32(key), out)
        assert (out[0], out[1]) == exp


def test_tables_blob_is_what_the_reference_headers_hold():
    """moquimc_b200/data/mqi_tables_v1.bin (compiled into libmqi_b200.so) byte for byte against the tables the
    reference's own headers hold, dumped by oracle/_ref/ref_kat_release (generator: oracle/gen_tables.py).  Needs the
    reference build of the container; on a box without it the checksums in kat_release.npz pin the same arrays."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(here, "..", "oracle", "gen_tables.py"))
    gt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gt)
    blob = open(gt.BLOB, "rb").read()
    assert blob[:8] == b"MQITBL1\0" and len(blob) == 16 + 4 * (3600 + 3996)
    if os.path.exists(gt.KAT):
        assert gt.build_blob() == blob
    kat = np.load(os.path.join(here, "golden", "kat_release.npz"))
    f = np.frombuffer(blob[16:], dtype="<f4")
    np.testing.assert_allclose(f[:3600].astype(np.float64).sum(), float(kat["tables_sum"]), rtol=0, atol=0)
    np.testing.assert_allclose(f[3600:].astype(np.float64).sum(), float(kat["density_correction_sum"]), rtol=0, atol=0)
