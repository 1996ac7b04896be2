"""ctypes binding of libmqi_b200.so (include/mqi_b200.h).

The product path is the CUDA library: importing works anywhere (so that CPU-only tests can check
that the library loads and exports its symbols), but every compute call fails loudly with
MqiError when no CUDA device is usable.  There is no CPU fallback and nothing here imports oracle/.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MQI_B200_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("MQI_B200_LIB") or os.path.join(HERE, "libmqi_b200.so")
HEADER = os.path.join(HERE, "..", "include", "mqi_b200.h")

PHYSICS_RELEASE, PHYSICS_DEBUG = 0, 1
SCORER_DOSE, SCORER_EDEP, SCORER_LETD_NUMER, SCORER_LETD_DENOM, SCORER_DOSE_SQ, SCORER_DIJ, SCORER_LETT_NUMER, SCORER_LETT_DENOM = range(8)
QUIRK_B2_DOUBLE_SCORE = 1
ACCUM_ATOMIC, ACCUM_WARP_MATCH = 0, 1
ENODEVICE = -2


class MqiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mqi error %d: %s" % (code, msg))
        self.code = code


class Vertex(C.Structure):
    _fields_ = [("ke", C.c_float), ("pos", C.c_float * 3), ("dir", C.c_float * 3)]


class Beamlet(C.Structure):
    _fields_ = [("phsp_uniform", C.c_int32), ("energy_normal", C.c_int32), ("energy", C.c_float),
                ("sigma_energy", C.c_float), ("mean", C.c_float * 6), ("sigma", C.c_float * 6),
                ("corr", C.c_float * 2), ("rot", C.c_float * 9), ("trans", C.c_float * 3)]


class RunStats(C.Structure):
    _fields_ = [("histories", C.c_uint64), ("steps", C.c_uint64), ("secondaries", C.c_uint64),
                ("stack_overflows", C.c_uint64), ("dij_table_full", C.c_uint64), ("kernel_ms", C.c_float),
                ("launches", C.c_uint32)]


_lib = None


def load():
    """Load libmqi_b200.so; raises if it has not been built (python -m moquimc_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MqiError(-5, "libmqi_b200.so is not built; run `python -m moquimc_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.mqi_last_error.restype = C.c_char_p
    L.mqi_version.restype = C.c_char_p
    vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    fp = C.POINTER(C.c_float)
    sig = {
        "mqi_device_count": [],
        "mqi_create": [i32, C.POINTER(vp)],
        "mqi_destroy": [vp],
        "mqi_device_memory": [vp, C.POINTER(u64), C.POINTER(u64)],
        "mqi_set_physics": [vp, i32, u32],
        "mqi_set_grid_hu": [vp, fp, i32, fp, i32, fp, i32, vp, f32, fp, fp],
        "mqi_set_grid_hu_device": [vp, fp, i32, fp, i32, fp, i32, vp, f32, fp, fp],
        "mqi_set_grid_density": [vp, fp, i32, fp, i32, fp, i32, fp, fp, fp],
        "mqi_add_scorer": [vp, i32, C.c_char_p, u64],
        "mqi_bind_scorer_buffer": [vp, i32, vp],
        "mqi_set_scorer_roi": [vp, i32, vp, u64, C.POINTER(u64)],
        "mqi_add_beamline_node": [vp, fp, i32, fp, i32, fp, i32, fp, fp, fp],
        "mqi_clear_beamline": [vp],
        "mqi_clear_scorers": [vp],
        "mqi_set_accumulation": [vp, i32],
        "mqi_set_beamlets": [vp, C.POINTER(Beamlet), u32, C.POINTER(u64)],
        "mqi_set_vertices": [vp, vp, u64, vp],
        "mqi_run": [vp, u64, u64, u64, i32],
        "mqi_run_async": [vp, u64, u64, u64, i32],
        "mqi_run_async_sharded": [vp, u64, u64, u64, i32, u32, u32],
        "mqi_set_stream": [vp, vp],
        "mqi_get_run_stats": [vp, C.POINTER(RunStats)],
        "mqi_set_option": [vp, C.c_char_p, C.c_int64],
        "mqi_get_dense": [vp, i32, vp, C.c_double],
        "mqi_get_sparse_count": [vp, i32, C.POINTER(u64)],
        "mqi_get_sparse": [vp, i32, vp, vp, vp, u64, C.c_double],
        "mqi_get_scorer_device_ptr": [vp, i32, C.POINTER(vp), C.POINTER(u64)],
        "mqi_stat_partial": [vp, i32, i32, u64, C.c_double, C.c_double, C.POINTER(C.c_double)],
        "mqi_stat_partial_buffers": [vp, vp, vp, u64, u64, C.c_double, C.c_double, C.POINTER(C.c_double)],
        "mqi_scale_scorer": [vp, i32, C.c_double],
        "mqi_reduce_dense": [C.POINTER(vp), i32, i32, i32],
        "mqi_allreduce_dense": [C.POINTER(vp), i32, i32],
        "mqi_stat_multi": [C.POINTER(vp), i32, i32, i32, u64, C.c_double, C.POINTER(C.c_double)],
        "mqi_dev_hu_to_density": [vp, vp, u64, f32, vp],
        "mqi_dev_rsp": [vp, vp, vp, u64, vp, vp],
        "mqi_dev_grid_step": [vp, vp, vp, u64, vp, vp, vp, vp, vp, vp],
        "mqi_dev_grid_entry": [vp, vp, vp, u64, vp, vp],
        "mqi_dev_hash": [vp, vp, vp, vp, u64, vp],
        "mqi_dev_sample_vertices": [vp, u64, u64, u64, vp, vp],
        "mqi_dev_insert": [vp, i32, vp, vp, vp, u64],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


def header_symbols():
    """Function names declared in include/mqi_b200.h (used by the CPU-side export test)."""
    import re
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MQI_API\s+(?:const\s+char\*|int)\s+(mqi_\w+)\s*\(", text)))


def device_count():
    return load().mqi_device_count()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def uniform_edges(lo, hi, n):
    """Edges as grid3d(xe_min, xe_max, n_xe) computes them in fp32 (mqi_grid3d.hpp:152-161)."""
    lo = np.float32(lo)
    hi = np.float32(hi)
    dx = np.float32((hi - lo) / np.float32(n))
    return (lo + np.arange(n + 1, dtype=np.float32) * dx).astype(np.float32)


def make_beamlet(energy, mean, sigma, uniform=True, sigma_energy=0.0, corr=(0.0, 0.0), rot=None,
                 trans=(0.0, 0.0, 0.0)):
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


def _handles(engines):
    return (C.c_void_p * len(engines))(*[e.h.value for e in engines])


def reduce_dense(engines, scorer, root=0):
    """Sum dense scorer `scorer` of the engines (one per GPU, one process) into engines[root]: one ncclReduce."""
    L = load()
    rc = L.mqi_reduce_dense(_handles(engines), len(engines), scorer, root)
    if rc != 0:
        raise MqiError(rc, L.mqi_last_error().decode())


def stat_multi(engines, s_sum, s_sq, n_histories, threshold):
    """calculate_stat over the summed stat grids of the engines without gathering them: reduce-scatter + slices.
    Returns (sum of sigma/mu, selected voxels, largest mean dose)."""
    L = load()
    out = (C.c_double * 3)()
    rc = L.mqi_stat_multi(_handles(engines), len(engines), s_sum, s_sq, n_histories, threshold, out)
    if rc != 0:
        raise MqiError(rc, L.mqi_last_error().decode())
    return out[0], out[1], out[2]


class Engine:
    """One transport engine on one GPU: the x_environment life cycle (initialize -> run -> finalize,
    moqui/base/environments/mqi_xenvironment.hpp:89-146) over the C ABI."""

    def __init__(self, device=0, physics=PHYSICS_RELEASE, quirks=0):
        self.L = load()
        self.h = C.c_void_p()
        self._check(self.L.mqi_create(device, C.byref(self.h)))
        self._check(self.L.mqi_set_physics(self.h, physics, quirks))
        self.physics = physics
        self.shape = None
        self.scorers = []
        self._keep = []

    def _check(self, rc):
        if rc < 0:
            raise MqiError(rc, self.L.mqi_last_error().decode())
        return rc

    def device_memory(self):
        """(free, total) bytes of HBM on the engine's device"""
        f, t = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.mqi_device_memory(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    def close(self):
        if self.h:
            self.L.mqi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- geometry
    def set_grid_hu(self, xe, ye, ze, hu, density_scale=1.0, rot=None, trans=None):
        xe, ye, ze = _f32(xe), _f32(ye), _f32(ze)
        hu = np.ascontiguousarray(hu, dtype=np.int16)
        assert hu.size == (len(xe) - 1) * (len(ye) - 1) * (len(ze) - 1)
        r = _f32(rot).ravel() if rot is not None else None
        t = _f32(trans) if trans is not None else None
        self._check(self.L.mqi_set_grid_hu(self.h, _fp(xe), len(xe), _fp(ye), len(ye), _fp(ze), len(ze),
                                           hu.ctypes.data, density_scale, _fp(r), _fp(t)))
        self.shape = (len(ze) - 1, len(ye) - 1, len(xe) - 1)

    def set_grid_hu_device(self, xe, ye, ze, d_hu_ptr, density_scale=1.0):
        xe, ye, ze = _f32(xe), _f32(ye), _f32(ze)
        self._check(self.L.mqi_set_grid_hu_device(self.h, _fp(xe), len(xe), _fp(ye), len(ye), _fp(ze), len(ze),
                                                  C.c_void_p(d_hu_ptr), density_scale, None, None))
        self.shape = (len(ze) - 1, len(ye) - 1, len(xe) - 1)

    def set_grid_density(self, xe, ye, ze, rho, rot=None, trans=None):
        xe, ye, ze = _f32(xe), _f32(ye), _f32(ze)
        rho = _f32(rho)
        r = _f32(rot).ravel() if rot is not None else None
        t = _f32(trans) if trans is not None else None
        self._check(self.L.mqi_set_grid_density(self.h, _fp(xe), len(xe), _fp(ye), len(ye), _fp(ze), len(ze),
                                                _fp(rho), _fp(r), _fp(t)))
        self.shape = (len(ze) - 1, len(ye) - 1, len(xe) - 1)

    def add_beamline_node(self, xe, ye, ze, rho, rot=None, trans=None):
        """A child of the world in front of the scored grid (range shifter slab, voxelised aperture),
        create_rangeshifter / create_voxelized_aperture mqi_tps_env.hpp:1605-1736.  rho in g/mm^3."""
        xe, ye, ze = _f32(xe), _f32(ye), _f32(ze)
        rho = _f32(rho)
        assert rho.size == (len(xe) - 1) * (len(ye) - 1) * (len(ze) - 1)
        r = _f32(rot).ravel() if rot is not None else None
        t = _f32(trans) if trans is not None else None
        return self._check(self.L.mqi_add_beamline_node(self.h, _fp(xe), len(xe), _fp(ye), len(ye), _fp(ze), len(ze),
                                                        _fp(rho), _fp(r), _fp(t)))

    def clear_beamline(self):
        self._check(self.L.mqi_clear_beamline(self.h))

    @property
    def nvox(self):
        return int(np.prod(self.shape))

    # ---- scorers
    def add_scorer(self, kind, name="", capacity=0):
        i = self._check(self.L.mqi_add_scorer(self.h, kind, name.encode(), capacity))
        self.scorers.append((kind, name))
        return i

    def set_scorer_roi(self, scorer, mask_total):
        """CONTOUR roi from the summed 0/1 mask volume (mask_reader::mask_to_roi); None = DIRECT roi.
        Returns the roi size (get_mask_size)."""
        n = C.c_uint64()
        if mask_total is None:
            self._check(self.L.mqi_set_scorer_roi(self.h, scorer, None, 0, C.byref(n)))
        else:
            m = np.ascontiguousarray(mask_total, dtype=np.uint8).ravel()
            self._check(self.L.mqi_set_scorer_roi(self.h, scorer, m.ctypes.data, m.size, C.byref(n)))
        return n.value

    def bind_scorer_buffer(self, scorer, device_ptr):
        self._check(self.L.mqi_bind_scorer_buffer(self.h, scorer, C.c_void_p(device_ptr)))

    def clear_scorers(self):
        self._check(self.L.mqi_clear_scorers(self.h))

    def set_accumulation(self, mode):
        self._check(self.L.mqi_set_accumulation(self.h, mode))

    def set_option(self, key, value):
        self._check(self.L.mqi_set_option(self.h, key.encode(), int(value)))

    # ---- source
    def set_beamlets(self, beamlets, histories_per_spot):
        n = len(beamlets)
        arr = (Beamlet * n)(*beamlets)
        hps = np.ascontiguousarray(histories_per_spot, dtype=np.uint64)
        assert hps.size == n
        self._check(self.L.mqi_set_beamlets(self.h, arr, n, hps.ctypes.data_as(C.POINTER(C.c_uint64))))
        self.total_histories = int(hps.sum())

    def set_vertices(self, vertices, spot_ids=None):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 7)
        s = None if spot_ids is None else np.ascontiguousarray(spot_ids, dtype=np.uint32)
        self._check(self.L.mqi_set_vertices(self.h, v.ctypes.data, len(v), None if s is None else s.ctypes.data))
        self.total_histories = len(v)

    # ---- run / results
    def run(self, seed, first, count, per_spot=False):
        self._check(self.L.mqi_run(self.h, seed, first, count, 1 if per_spot else 0))
        st = RunStats()
        self._check(self.L.mqi_get_run_stats(self.h, C.byref(st)))
        return st

    def run_sharded(self, seed, first, count, n_shards, shard, per_spot=False):
        """this device's interleaved share (chunks of 32 histories, chunk c to shard c % n_shards) of [first, first + count)"""
        self._check(self.L.mqi_run_async_sharded(self.h, seed, first, count, 1 if per_spot else 0, n_shards, shard))
        return self.run_stats()

    def run_async(self, seed, first, count, per_spot=False):
        self._check(self.L.mqi_run_async(self.h, seed, first, count, 1 if per_spot else 0))

    def run_stats(self):
        st = RunStats()
        self._check(self.L.mqi_get_run_stats(self.h, C.byref(st)))
        return st

    def set_stream(self, cuda_stream):
        self._check(self.L.mqi_set_stream(self.h, C.c_void_p(cuda_stream)))

    def get_dense(self, scorer, scale=1.0, out=None):
        """out: optional caller-owned host buffer (e.g. pinned memory) of nvox float64."""
        if out is None:
            out = np.empty(self.nvox, dtype=np.float64)
        assert out.size == self.nvox and out.dtype == np.float64
        self._check(self.L.mqi_get_dense(self.h, scorer, out.ctypes.data, scale))
        return out.reshape(self.shape)

    def get_sparse_count(self, scorer):
        nnz = C.c_uint64()
        self._check(self.L.mqi_get_sparse_count(self.h, scorer, C.byref(nnz)))
        return nnz.value

    def get_sparse(self, scorer, scale=1.0):
        nnz = C.c_uint64()
        self._check(self.L.mqi_get_sparse_count(self.h, scorer, C.byref(nnz)))
        n = nnz.value
        k1 = np.empty(n, dtype=np.uint32)
        k2 = np.empty(n, dtype=np.uint32)
        v = np.empty(n, dtype=np.float64)
        self._check(self.L.mqi_get_sparse(self.h, scorer, k1.ctypes.data, k2.ctypes.data, v.ctypes.data, n, scale))
        return k1, k2, v

    def scorer_device_ptr(self, scorer):
        p = C.c_void_p()
        n = C.c_uint64()
        self._check(self.L.mqi_get_scorer_device_ptr(self.h, scorer, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stat_partial(self, s_sum, s_sq, n_histories, threshold, max_mean=-1.0):
        out = (C.c_double * 3)()
        self._check(self.L.mqi_stat_partial(self.h, s_sum, s_sq, n_histories, threshold, max_mean, out))
        return out[0], out[1], out[2]

    def stat_partial_buffers(self, d_sum_ptr, d_sumsq_ptr, n_voxels, n_histories, threshold, max_mean=-1.0):
        out = (C.c_double * 3)()
        self._check(self.L.mqi_stat_partial_buffers(self.h, C.c_void_p(d_sum_ptr), C.c_void_p(d_sumsq_ptr), n_voxels,
                                                    n_histories, threshold, max_mean, out))
        return out[0], out[1], out[2]

    def scale_scorer(self, scorer, factor):
        self._check(self.L.mqi_scale_scorer(self.h, scorer, factor))

    # ---- deterministic device pieces
    def dev_hu_to_density(self, hu, density_scale=1.0):
        hu = np.ascontiguousarray(hu, dtype=np.int16)
        out = np.empty(hu.size, dtype=np.float32)
        self._check(self.L.mqi_dev_hu_to_density(self.h, hu.ctypes.data, hu.size, density_scale, out.ctypes.data))
        return out.reshape(hu.shape)

    def dev_rsp(self, rho, ek):
        rho, ek = _f32(rho), _f32(ek)
        rsp = np.empty(rho.size, dtype=np.float32)
        rl = np.empty(rho.size, dtype=np.float32)
        self._check(self.L.mqi_dev_rsp(self.h, rho.ctypes.data, ek.ctypes.data, rho.size, rsp.ctypes.data, rl.ctypes.data))
        return rsp, rl

    def dev_grid_step(self, p, d):
        p, d = _f32(p).reshape(-1, 3), _f32(d).reshape(-1, 3)
        n = len(p)
        cell = np.empty((n, 3), dtype=np.int32)
        cnb = np.empty(n, dtype=np.uint64)
        dist = np.empty(n, dtype=np.float32)
        dir_after = np.empty((n, 3), dtype=np.float32)
        p_exit = np.empty((n, 3), dtype=np.float32)
        cell_after = np.empty((n, 3), dtype=np.int32)
        self._check(self.L.mqi_dev_grid_step(self.h, p.ctypes.data, d.ctypes.data, n, cell.ctypes.data, cnb.ctypes.data,
                                             dist.ctypes.data, dir_after.ctypes.data, p_exit.ctypes.data,
                                             cell_after.ctypes.data))
        return cell, cnb, dist, dir_after, p_exit, cell_after

    def dev_grid_entry(self, p, d):
        p, d = _f32(p).reshape(-1, 3), _f32(d).reshape(-1, 3)
        n = len(p)
        dist = np.empty(n, dtype=np.float32)
        cell = np.empty((n, 3), dtype=np.int32)
        self._check(self.L.mqi_dev_grid_entry(self.h, p.ctypes.data, d.ctypes.data, n, dist.ctypes.data, cell.ctypes.data))
        return dist, cell

    def dev_hash(self, k1, k2, cap):
        k1 = np.ascontiguousarray(k1, dtype=np.uint32)
        k2 = np.ascontiguousarray(k2, dtype=np.uint32)
        cap = np.ascontiguousarray(cap, dtype=np.uint64)
        out = np.empty(k1.size, dtype=np.uint32)
        self._check(self.L.mqi_dev_hash(self.h, k1.ctypes.data, k2.ctypes.data, cap.ctypes.data, k1.size, out.ctypes.data))
        return out

    def dev_insert(self, scorer, key1, key2, value):
        key1 = np.ascontiguousarray(key1, dtype=np.uint32)
        key2 = np.ascontiguousarray(key2, dtype=np.uint32)
        value = np.ascontiguousarray(value, dtype=np.float64)
        assert key1.size == key2.size == value.size
        self._check(self.L.mqi_dev_insert(self.h, scorer, key1.ctypes.data, key2.ctypes.data, value.ctypes.data, key1.size))

    def dev_sample_vertices(self, seed, first, n):
        v = np.empty((n, 7), dtype=np.float32)
        s = np.empty(n, dtype=np.uint32)
        self._check(self.L.mqi_dev_sample_vertices(self.h, seed, first, n, v.ctypes.data, s.ctypes.data))
        return v, s
