"""The product's multi-GPU paths on real hardware (run on a box with >= 2 GPUs: `gpurun --gpus 8 -- python -m pytest
tests/test_gpu_multi.py -m gpu`; skipped on a single-GPU box).  One process drives all devices through the C ABI
(mqi_reduce_dense / mqi_stat_multi over NCCL) exactly as tps_env does with a GPUID list.

  * tps_env GPUID 0,...,N-1: dose equals the single-device dose (histories sharded by index range, one reduce);
  * a Dij run with the spots sharded over N devices equals the single-device table after a canonical sort;
  * mqi_stat_multi (reduce-scatter of the stat pair, criterion in slices) equals mqi_stat_partial on the summed grids;
  * the stopping loop of tps_env over N devices reaches the criterion with the same number of passes.
"""
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
N, SP = (96, 96, 60), (2.0, 2.0, 3.0)


def n_devices():
    try:
        return capi.device_count()
    except Exception:
        return 0


needs_multi = pytest.mark.skipif(n_devices() < 2, reason="needs at least two GPUs")


def run_tps(inp):
    r = subprocess.run([EXE, inp], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@needs_multi
def test_tps_env_on_all_devices_gives_the_single_device_dose(tmp_path):
    root = str(tmp_path)
    nd = n_devices()
    a, b = os.path.join(root, "o1"), os.path.join(root, "oN")
    inp = S.make_case(root, n=N, spacing=SP, n_layers=6, ParticlesPerHistory=1000.0, OutputDir=a, Scorer="Dose,LETd")
    run_tps(inp)
    i2 = os.path.join(root, "multi.in")
    S.write_input(i2, root, b, ParticlesPerHistory=1000.0, Scorer="Dose,LETd", GPUID=",".join(str(i) for i in range(nd)))
    out = run_tps(i2)
    assert "on %d GPU(s)" % nd in out
    for name in ("Dose", "LETd_numer", "LETd_denom"):
        d1 = np.fromfile(os.path.join(a, "G000_0_%s.raw" % name), dtype=np.float64)
        dn = np.fromfile(os.path.join(b, "G000_0_%s.raw" % name), dtype=np.float64)
        assert d1.sum() > 0
        # the same histories with the same streams, summed in another order
        np.testing.assert_allclose(dn, d1, rtol=1e-9, atol=d1.max() * 1e-13)


@needs_multi
def test_spot_sharded_dij_equals_the_single_device_table(tmp_path):
    root = str(tmp_path)
    nd = n_devices()
    a, b = os.path.join(root, "o1"), os.path.join(root, "oN")
    kw = dict(ParticlesPerHistory=1.0, Scorer="Dij", SimulationType="perSpot", UnitWeights=2000, OutputFormat="npz", DijCapacity=20_000_001)
    inp = S.make_case(root, n=N, spacing=SP, n_layers=5, OutputDir=a, **kw)
    run_tps(inp)
    i2 = os.path.join(root, "multi.in")
    S.write_input(i2, root, b, GPUID=",".join(str(i) for i in range(nd)), **kw)
    run_tps(i2)
    import scipy.sparse as sp
    m1 = sp.load_npz(os.path.join(a, "G000_0_Dij.npz")).tocsr()
    mn = sp.load_npz(os.path.join(b, "G000_0_Dij.npz")).tocsr()
    m1.sort_indices()
    mn.sort_indices()
    assert m1.shape == mn.shape and m1.nnz == mn.nnz > 0
    np.testing.assert_array_equal(m1.indptr, mn.indptr)
    np.testing.assert_array_equal(m1.indices, mn.indices)
    np.testing.assert_allclose(mn.data, m1.data, rtol=1e-9)


def small_engine(dev):
    e = capi.Engine(dev, physics=capi.PHYSICS_RELEASE)
    e.set_grid_hu(capi.uniform_edges(-40, 40, 80), capi.uniform_edges(-40, 40, 80), capi.uniform_edges(-160, 0, 107),
                  np.zeros((107, 80, 80), dtype=np.int16))
    ids = (e.add_scorer(capi.SCORER_DOSE, "Dose"), e.add_scorer(capi.SCORER_DOSE, "Dose_stat"), e.add_scorer(capi.SCORER_DOSE_SQ, "DoseSquare_stat"))
    return e, ids


@needs_multi
def test_stat_multi_equals_the_criterion_on_the_summed_grids():
    nd = n_devices()
    n = 400_000
    bl = [capi.make_beamlet(110.0, [0, 0, 0.5, 0, 0, -1], [8, 8, 0, 0, 0, 0], uniform=True)]
    engines = []
    for d in range(nd):
        e, ids = small_engine(d)
        e.set_beamlets(bl, [n])
        e.run_sharded(5, 0, n, nd, d)       # interleaved shards, as tps_env runs a beam over a GPUID list
        engines.append(e)
    s, c, mx = capi.stat_multi(engines, ids[1], ids[2], n, 0.5)
    # reference: everything on one device
    e1, _ = small_engine(0)
    e1.set_beamlets(bl, [n])
    e1.run(5, 0, n)
    s1, c1, mx1 = e1.stat_partial(ids[1], ids[2], n, 0.5)
    assert c == c1 > 100
    np.testing.assert_allclose(s, s1, rtol=1e-9)
    np.testing.assert_allclose(mx, mx1, rtol=1e-12)
    # the grids were not touched by the evaluation: reducing them now gives the single-device dose
    capi.reduce_dense(engines, ids[0], 0)
    np.testing.assert_allclose(engines[0].get_dense(ids[0]), e1.get_dense(ids[0]), rtol=1e-9, atol=1e-25)


@needs_multi
def test_stopping_loop_over_all_devices(tmp_path):
    root = str(tmp_path)
    nd = n_devices()
    res = {}
    for tag, gpus in (("one", "0"), ("all", ",".join(str(i) for i in range(nd)))):
        od = os.path.join(root, "o_" + tag)
        inp = S.make_case(os.path.join(root, tag), n=N, spacing=SP, n_layers=6, ParticlesPerHistory=2000.0, OutputDir=od,
                          StoppingStatistics="true", StoppingCriteria="3.0", StatThreshold="0.5", MaxStatPasses=60, GPUID=gpus)
        out = run_tps(inp)
        runs = [(int(a), float(b)) for a, b in re.findall(r"Run (\d+): current uncertainty ([0-9.eE+-]+) %", out)]
        res[tag] = (runs, np.fromfile(os.path.join(od, "G000_0_Dose.raw"), dtype=np.float64))
    (r1, d1), (rn, dn) = res["one"], res["all"]
    assert r1 and r1[-1][1] <= 3.0 and rn[-1][1] <= 3.0
    assert len(r1) == len(rn)
    # the same histories pass by pass => the same uncertainty after every pass (fp64 summation order apart)
    np.testing.assert_allclose([u for _, u in rn], [u for _, u in r1], rtol=1e-6)
    np.testing.assert_allclose(dn, d1, rtol=1e-9, atol=d1.max() * 1e-13)
