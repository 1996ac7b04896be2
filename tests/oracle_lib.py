"""ctypes binding of oracle/libmqi_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
TABLES = os.path.join(ROOT, "moquimc_b200", "data", "mqi_tables_v1.bin")

VARIANT_RELEASE, VARIANT_DEBUG = 0, 1
SCORER_DOSE, SCORER_EDEP, SCORER_LETD_NUMER, SCORER_LETD_DENOM, SCORER_DOSE_SQ, SCORER_DIJ, SCORER_LETT_NUMER, SCORER_LETT_DENOM = range(8)
QUIRK_B2_DOUBLE_SCORE = 1
EMPTY = 0xFFFFFFFF


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("xe", C.POINTER(C.c_float)), ("ye", C.POINTER(C.c_float)), ("ze", C.POINTER(C.c_float)),
                ("rho", C.POINTER(C.c_float)), ("rot_fwd", C.c_float * 9), ("trans", C.c_float * 3)]


class Beamlet(C.Structure):
    _fields_ = [("phsp_uniform", C.c_int), ("energy_normal", C.c_int), ("energy", C.c_float),
                ("sigma_energy", C.c_float), ("mean", C.c_float * 6), ("sigma", C.c_float * 6),
                ("corr", C.c_float * 2), ("rot", C.c_float * 9), ("trans", C.c_float * 3)]


class Vertex(C.Structure):
    _fields_ = [("ke", C.c_float), ("pos", C.c_float * 3), ("dir", C.c_float * 3)]


class KeyValue(C.Structure):
    _fields_ = [("key1", C.c_uint32), ("key2", C.c_uint32), ("value", C.c_double)]


class Scorer(C.Structure):
    _fields_ = [("kind", C.c_int), ("dense", C.POINTER(C.c_double)), ("table", C.POINTER(KeyValue)),
                ("capacity", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("histories", "steps", "along_steps", "delta_events", "pp_events",
                                           "poe_events", "poi_events", "secondaries_pushed", "max_stack")]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ORACLE_DIR, "libmqi_oracle.so")
    src = os.path.join(ORACLE_DIR, "mqi_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libmqi_oracle.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    L.mqo_load_tables.argtypes = [C.c_char_p]
    L.mqo_hu_to_density.restype = C.c_float
    L.mqo_hu_to_density.argtypes = [C.c_int16]
    L.mqo_spr.restype = C.c_float
    L.mqo_spr.argtypes = [C.c_float, C.c_float, C.c_int]
    L.mqo_radiation_length.restype = C.c_float
    L.mqo_radiation_length.argtypes = [C.c_float, C.c_int]
    L.mqo_hash.restype = C.c_uint32
    L.mqo_hash.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.mqo_grid_intersect_cell.restype = C.c_float
    L.mqo_grid_intersect_entry.restype = C.c_float
    L.mqo_u32_to_uniform.restype = C.c_float
    L.mqo_u32_to_uniform.argtypes = [C.c_uint32]
    L.mqo_transport.restype = C.c_int
    rc = L.mqo_load_tables(TABLES.encode())
    if rc != 0:
        raise RuntimeError("mqo_load_tables failed: %d" % rc)
    _lib = L
    return L


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_grid(xe, ye, ze, rho, rot=None, trans=None):
    """Returns (Grid, keepalive)."""
    xe = np.ascontiguousarray(xe, dtype=np.float32)
    ye = np.ascontiguousarray(ye, dtype=np.float32)
    ze = np.ascontiguousarray(ze, dtype=np.float32)
    rho = np.ascontiguousarray(rho, dtype=np.float32).ravel()
    g = Grid()
    g.nx, g.ny, g.nz = len(xe) - 1, len(ye) - 1, len(ze) - 1
    assert rho.size == g.nx * g.ny * g.nz
    g.xe, g.ye, g.ze, g.rho = fptr(xe), fptr(ye), fptr(ze), fptr(rho)
    r = np.eye(3, dtype=np.float32).ravel() if rot is None else np.asarray(rot, dtype=np.float32).ravel()
    t = np.zeros(3, dtype=np.float32) if trans is None else np.asarray(trans, dtype=np.float32)
    g.rot_fwd = (C.c_float * 9)(*r)
    g.trans = (C.c_float * 3)(*t)
    return g, (xe, ye, ze, rho)


def uniform_edges(lo, hi, n):
    """grid3d(xe_min, xe_max, n_xe) edge rule: xe_min + i*dx in fp32 (mqi_grid3d.hpp:152-161)."""
    lo = np.float32(lo)
    hi = np.float32(hi)
    dx = np.float32((hi - lo) / np.float32(n))
    return (lo + np.arange(n + 1, dtype=np.float32) * dx).astype(np.float32)


def hu_to_density(hu):
    L = lib()
    hu = np.asarray(hu, dtype=np.int16)
    return np.array([L.mqo_hu_to_density(int(h)) for h in hu.ravel()], dtype=np.float32).reshape(hu.shape)


def make_beamlet(energy, mean, sigma, uniform=True, sigma_energy=0.0, corr=(0, 0), rot=None, trans=(0, 0, 0)):
    b = Beamlet()
    b.phsp_uniform = 1 if uniform else 0
    b.energy_normal = 1 if sigma_energy > 0 else 0
    b.energy = energy
    b.sigma_energy = sigma_energy
    b.mean = (C.c_float * 6)(*mean)
    b.sigma = (C.c_float * 6)(*sigma)
    b.corr = (C.c_float * 2)(*corr)
    r = np.eye(3, dtype=np.float32).ravel() if rot is None else np.asarray(rot, dtype=np.float32).ravel()
    b.rot = (C.c_float * 9)(*r)
    b.trans = (C.c_float * 3)(*trans)
    return b


def insert(kind, key1, key2, value, capacity=0, nvox=0):
    """mc::insert_hashtable restatement applied to a hit list, sequentially.  Returns the occupied
    slots (structured array, slot order) for Dij, or the dense float64 array."""
    L = lib()
    key1 = np.ascontiguousarray(key1, dtype=np.uint32)
    key2 = np.ascontiguousarray(key2, dtype=np.uint32)
    value = np.ascontiguousarray(value, dtype=np.float64)
    sc = Scorer()
    sc.kind = kind
    if kind == SCORER_DIJ:
        tab = np.zeros(capacity, dtype=[("key1", "<u4"), ("key2", "<u4"), ("value", "<f8")])
        tab["key1"] = EMPTY
        tab["key2"] = EMPTY
        sc.table = tab.ctypes.data_as(C.POINTER(KeyValue))
        sc.capacity = capacity
    else:
        tab = np.zeros(nvox, dtype=np.float64)
        sc.dense = tab.ctypes.data_as(C.POINTER(C.c_double))
    L.mqo_insert(C.byref(sc), key1.ctypes.data_as(C.POINTER(C.c_uint32)), key2.ctypes.data_as(C.POINTER(C.c_uint32)),
                 value.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(key1.size))
    if kind == SCORER_DIJ:
        occ = (tab["key1"] != EMPTY) & (tab["key2"] != EMPTY)
        return tab[occ].copy(), np.nonzero(occ)[0]
    return tab


def mask_to_roi(mask_total):
    """mask_reader::mask_to_roi restated: returns (start, stride, member) of the run-length CONTOUR roi."""
    L = lib()
    m = np.ascontiguousarray(mask_total, dtype=np.uint8).ravel()
    member = np.zeros(m.size, dtype=np.uint8)
    L.mqo_mask_to_roi.restype = C.c_uint32
    u8 = C.POINTER(C.c_uint8)
    u32 = C.POINTER(C.c_uint32)
    n = L.mqo_mask_to_roi(m.ctypes.data_as(u8), C.c_uint64(m.size), None, None, C.c_uint32(0), member.ctypes.data_as(u8))
    start = np.zeros(max(n, 1), dtype=np.uint32)
    stride = np.zeros(max(n, 1), dtype=np.uint32)
    L.mqo_mask_to_roi(m.ctypes.data_as(u8), C.c_uint64(m.size), start.ctypes.data_as(u32), stride.ctypes.data_as(u32),
                      C.c_uint32(n), None)
    return start[:n], stride[:n], member


def transport(grid, variant, beamlets, histories_per_spot, seed, h0, n, kinds, quirks=0, per_spot=False,
              dij_capacity=0, vertices=None, spot_ids=None, roi_members=None):
    """Run the oracle; returns (list of outputs per scorer, Stats).  Dense scorers -> float64[nvox];
    Dij -> structured array of occupied slots (key1, key2, value).  `grid` is one Grid or a list of
    Grids (beamline nodes first, the scored patient grid last); roi_members is an optional list (one
    entry per scorer, None = DIRECT roi) of expanded CONTOUR roi memberships (mask_to_roi()[2])."""
    L = lib()
    nodes = list(grid) if isinstance(grid, (list, tuple)) else [grid]
    grid = nodes[-1]
    narr = (Grid * len(nodes))(*nodes)
    nvox = grid.nx * grid.ny * grid.nz
    nb = len(beamlets)
    barr = (Beamlet * max(nb, 1))(*beamlets)
    cum = np.cumsum(np.asarray(histories_per_spot, dtype=np.uint64)).astype(np.uint64)
    sc = (Scorer * len(kinds))()
    keep = []
    for i, k in enumerate(kinds):
        sc[i].kind = k
        if k == SCORER_DIJ:
            tab = np.zeros(dij_capacity, dtype=[("key1", "<u4"), ("key2", "<u4"), ("value", "<f8")])
            tab["key1"] = EMPTY
            tab["key2"] = EMPTY
            sc[i].table = tab.ctypes.data_as(C.POINTER(KeyValue))
            sc[i].capacity = dij_capacity
            keep.append(tab)
        else:
            d = np.zeros(nvox, dtype=np.float64)
            sc[i].dense = d.ctypes.data_as(C.POINTER(C.c_double))
            keep.append(d)
    st = Stats()
    vptr = None
    sptr = None
    if vertices is not None:
        vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 7)
        vptr = vertices.ctypes.data_as(C.POINTER(Vertex))
        if spot_ids is not None:
            spot_ids = np.ascontiguousarray(spot_ids, dtype=np.uint32)
            sptr = spot_ids.ctypes.data_as(C.POINTER(C.c_uint32))
    rptr = None
    if roi_members is not None:
        assert len(roi_members) == len(kinds)
        rarr = (C.POINTER(C.c_uint8) * len(kinds))()
        for i, m in enumerate(roi_members):
            if m is not None:
                m = np.ascontiguousarray(m, dtype=np.uint8).ravel()
                assert m.size == nvox
                keep.append(m)
                rarr[i] = m.ctypes.data_as(C.POINTER(C.c_uint8))
        rptr = rarr
    L.mqo_transport_nodes.restype = C.c_int
    rc = L.mqo_transport_nodes(narr, C.c_int(len(nodes)), C.c_int(variant), C.c_uint32(quirks), barr,
                               cum.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint32(nb), vptr, sptr,
                               C.c_int(1 if per_spot else 0), C.c_uint64(seed), C.c_uint64(h0), C.c_uint64(n),
                               sc, C.c_int(len(kinds)), rptr, C.byref(st))
    if rc != 0:
        raise RuntimeError("mqo_transport_nodes rc=%d" % rc)
    outs = []
    for k, a in zip(kinds, keep):
        if k == SCORER_DIJ:
            occ = (a["key1"] != EMPTY) & (a["key2"] != EMPTY)
            outs.append(a[occ].copy())
        else:
            outs.append(a)
    return outs, st
