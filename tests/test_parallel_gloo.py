"""Host-side multi-GPU logic (moquimc_b200/parallel.py) with two gloo ranks on CPU.

The compute stand-in is the ORACLE (tests only): history h draws the same counter-based stream
whoever transports it, so two ranks that shard a history range and sum-reduce their dense grids must
reproduce the single-process grid to fp64 summation order -- the property the NCCL path of bench.py
and tps_env relies on.  Also covered: whole-spot sharding for Dij (disjoint rows), the 21 robust
scenarios round-robin, and the statistical stopping loop (rank-local running sums, reduce-scatter of the stat pair).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from moquimc_b200 import parallel as P  # noqa: E402

NX, NY, NZ = 20, 20, 60
N_HIST = 600


def small_setup():
    import oracle_lib as O
    xe, ye, ze = O.uniform_edges(-20, 20, NX), O.uniform_edges(-20, 20, NY), O.uniform_edges(-120, 0, NZ)
    rho = np.full(NX * NY * NZ, O.hu_to_density(np.array([0]))[0], dtype=np.float32)
    g, keep = O.make_grid(xe, ye, ze, rho)
    beamlets = [O.make_beamlet(90.0, [x0, 0, 0.5, 0, 0, -1], [3, 3, 0, 0, 0, 0], uniform=True) for x0 in (-8.0, 0.0, 8.0)]
    return O, g, keep, beamlets, [N_HIST // 3] * 3


def numpy_criterion(sum_slice, sq_slice, n, max_mean, threshold=0.5):
    """calculate_standard_deviation + calculate_stat (mqi_variables.hpp:20-48, mqi_tps_env.hpp:1409-1425) on one
    slice of the summed grids: (sum of sigma / mu, selected voxels, largest mean dose of the slice)"""
    s, q = sum_slice.numpy(), sq_slice.numpy()
    mean = s / n
    if max_mean < 0:
        return 0.0, 0, float(mean.max())
    var = (q / n - mean * mean) / (n - 1.0)
    sel = (mean > threshold * max_mean) & (mean > 0)
    return float((np.sqrt(np.maximum(var[sel], 0.0)) / mean[sel]).sum()), int(sel.sum()), float(mean.max())


def worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O, g, keep, beamlets, hps = small_setup()
        nvox = NX * NY * NZ
        # ---- dense scorer: shard the history range, one reduce
        first, count = P.history_shard(N_HIST, rank, world)
        (d,), _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=11, h0=first, n=count, kinds=[O.SCORER_DOSE])
        grid = torch.from_numpy(np.ascontiguousarray(d.reshape(-1)))
        P.reduce_dense(grid, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, "dense.npy"), grid.numpy())
        # ---- Dij: whole spots per rank, no reduction; rank r's rows are spots [s0, s0 + ns)
        s0, ns, h0, nh = P.spot_shard(hps, rank, world)
        outs, _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=11, h0=h0, n=nh, kinds=[O.SCORER_DIJ], per_spot=True,
                              dij_capacity=400_009)
        k1, k2, v = outs[0]["key1"], outs[0]["key2"], outs[0]["value"]
        assert len(k2) and k2.min() >= s0 and k2.max() < s0 + ns
        np.savez(os.path.join(out_dir, "dij_%d.npz" % rank), k1=k1, k2=k2, v=v)
        # ---- statistical stopping: fresh streams per pass, rank-local running sums, reduce-scatter of the stat pair,
        # identical decision on all ranks
        n_pad = P.StoppingLoop.padded_len(nvox)
        ts, tq = torch.zeros(n_pad, dtype=torch.float64), torch.zeros(n_pad, dtype=torch.float64)

        def transport_pass(k):
            (a, b), _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=100 + k, h0=first, n=count,
                                    kinds=[O.SCORER_DOSE, O.SCORER_DOSE_SQ])
            ts[:nvox] += torch.from_numpy(np.ascontiguousarray(a.reshape(-1)))
            tq[:nvox] += torch.from_numpy(np.ascontiguousarray(b.reshape(-1)))
            return count
        loop = P.StoppingLoop(60.0, transport_pass, numpy_criterion, threshold=0.5, max_passes=6)
        tracked, current, passes = loop.run(ts, tq)
        assert 0 < loop.exchanged_values < n_pad        # only the chunks that can pass the dose threshold travel
        np.save(os.path.join(out_dir, "stat_%d.npy" % rank), np.array([tracked, current, passes] + loop.history))
        P.reduce_dense(ts, dst=0)          # the dose itself travels once, after the last pass
        if rank == 0:
            np.save(os.path.join(out_dir, "stat_sum.npy"), ts[:nvox].numpy())
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_outputs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gloo"))
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    return out


def test_shard_helpers():
    for n, w in ((10, 3), (7, 8), (10**9 + 7, 8)):
        parts = [P.history_shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
    hps = [5, 0, 7, 3, 9]
    parts = [P.spot_shard(hps, r, 2) for r in range(2)]
    assert parts == [(0, 2, 0, 5), (2, 3, 5, 19)]
    hps = [3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5, 8]
    for w in (1, 2, 3, 8):
        blocks = [P.spot_shard_blocks(hps, r, w, 2) for r in range(w)]
        flat = sorted(b for bl in blocks for b in bl)
        assert flat[0][0] == 0 and sum(n for _, n in flat) == sum(hps)           # the blocks tile the histories exactly
        assert all(a + n == b for (a, n), (b, _) in zip(flat, flat[1:]))
        cum = np.concatenate(([0], np.cumsum(hps)))
        assert all(a in cum and a + n in cum for a, n in flat)                   # and cut between spots only
    sc = P.robust_scenarios()
    assert len(sc) == 21 and sc[0] == {"XShift": 0.0, "YShift": 0.0, "ZShift": 0.0, "DensityScaling": 1.0}
    assert sorted(sum((P.scenario_shard(21, r, 8) for r in range(8)), [])) == list(range(21))
    assert max(len(P.scenario_shard(21, r, 8)) for r in range(8)) == 3


def test_sharded_dense_dose_equals_single_process(two_rank_outputs):
    O, g, keep, beamlets, hps = small_setup()
    (d,), _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=11, h0=0, n=N_HIST, kinds=[O.SCORER_DOSE])
    got = np.load(os.path.join(two_rank_outputs, "dense.npy"))
    assert d.sum() > 0
    np.testing.assert_allclose(got, d.reshape(-1), rtol=1e-12, atol=d.max() * 1e-14)


def test_spot_sharded_dij_rows_are_disjoint_and_complete(two_rank_outputs):
    O, g, keep, beamlets, hps = small_setup()
    outs, _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=11, h0=0, n=N_HIST, kinds=[O.SCORER_DIJ], per_spot=True,
                          dij_capacity=400_009)
    k1, k2, v = outs[0]["key1"], outs[0]["key2"], outs[0]["value"]
    ref = {(int(a), int(b)): float(c) for a, b, c in zip(k1, k2, v)}
    got = {}
    for r in range(2):
        z = np.load(os.path.join(two_rank_outputs, "dij_%d.npz" % r))
        for a, b, c in zip(z["k1"], z["k2"], z["v"]):
            assert (int(a), int(b)) not in got           # disjoint: concatenation, no reduction
            got[(int(a), int(b))] = float(c)
    assert got.keys() == ref.keys()
    np.testing.assert_allclose([got[k] for k in sorted(ref)], [ref[k] for k in sorted(ref)], rtol=1e-12)


def test_stopping_loop_is_consistent_across_ranks(two_rank_outputs):
    a = np.load(os.path.join(two_rank_outputs, "stat_0.npy"))
    b = np.load(os.path.join(two_rank_outputs, "stat_1.npy"))
    np.testing.assert_array_equal(a, b)                  # same totals => same decision on every rank
    tracked, current, passes = a[:3]
    hist = a[3:]
    assert passes == len(hist) >= 1 and tracked == passes * N_HIST
    assert (current <= 60.0) or passes == 6
    assert all(x >= y * 0.8 for x, y in zip(hist, hist[1:]))
    # the running total equals a single-process accumulation of the same passes
    O, g, keep, beamlets, hps = small_setup()
    tot = np.zeros(NX * NY * NZ)
    for k in range(int(passes)):
        (d,), _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=100 + k, h0=0, n=N_HIST, kinds=[O.SCORER_DOSE])
        tot += d.reshape(-1)
    np.testing.assert_allclose(np.load(os.path.join(two_rank_outputs, "stat_sum.npy")), tot, rtol=1e-12, atol=tot.max() * 1e-14)
    # and the sliced criterion equals the criterion on whole grids
    sq = np.zeros(NX * NY * NZ)
    for k in range(int(passes)):
        (d2,), _ = O.transport(g, O.VARIANT_RELEASE, beamlets, hps, seed=100 + k, h0=0, n=N_HIST, kinds=[O.SCORER_DOSE_SQ])
        sq += d2.reshape(-1)
    s, c, mx = numpy_criterion(torch.from_numpy(tot), torch.from_numpy(sq), int(tracked), float((tot / tracked).max()))
    np.testing.assert_allclose(current, 100.0 * s / c, rtol=1e-9)


# ------------------------------------------------------------------------------------------------
# the chunk selection with rank-local bounds (StoppingLoop.kept_chunks) on synthetic grids, four ranks
# ------------------------------------------------------------------------------------------------
def _selection_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nvox = 40 * P.StoppingLoop.CHUNK + 123           # not a whole number of chunks: the buffers are padded
        n_pad = P.StoppingLoop.padded_len(nvox)
        # every rank has its own peak (height 1, in its own chunk) and a share of 0.3 of a common spot X in yet another chunk:
        # X sums to 1.2, the largest value of the summed grid, but no rank holds threshold x ITS OWN largest value there
        # (0.3 < 0.5 x 1); only the bound threshold / world x m_r (0.125) keeps its chunk
        rng = np.random.default_rng(7 + rank)
        x = np.arange(nvox, dtype=np.float64)
        C = P.StoppingLoop.CHUNK
        bump = lambda c, h: h * np.exp(-0.5 * ((x - c * C) / (0.05 * C)) ** 2)   # noqa: E731
        local = rng.random(nvox)
        d = (bump((5.5, 12.5, 20.5, 28.5)[rank], 1.0) + bump(35.5, 0.3)) * (0.95 + 0.1 * local)
        ts, tq = torch.zeros(n_pad, dtype=torch.float64), torch.zeros(n_pad, dtype=torch.float64)
        n_hist = 1000

        def transport_pass(k):
            ts[:nvox] += torch.from_numpy(d * n_hist)
            tq[:nvox] += torch.from_numpy((3.0 * d * (1.0 + 0.2 * local)) * n_hist)   # E[x^2] > mean^2 everywhere: positive variances
            return n_hist // world

        loop = P.StoppingLoop(0.0, transport_pass, numpy_criterion, threshold=0.5, max_passes=2, histories_per_pass=n_hist)
        tracked, current, passes = loop.run(ts, tq)
        assert (passes, tracked, loop.collectives_per_pass) == (2, 2 * n_hist, 5), (passes, tracked, loop.collectives_per_pass, loop.history)
        assert 0 < loop.exchanged_values < n_pad
        full_s, full_q = ts.clone(), tq.clone()
        dist.all_reduce(full_s)
        dist.all_reduce(full_q)
        if rank == 0:
            np.save(os.path.join(out_dir, "sel.npy"), np.array([current, loop.exchanged_values, n_pad]))
            np.save(os.path.join(out_dir, "sel_sum.npy"), full_s.numpy())
            np.save(os.path.join(out_dir, "sel_sq.npy"), full_q.numpy())
    finally:
        dist.destroy_process_group()


def test_chunk_selection_with_local_bounds_is_exact_on_four_ranks(tmp_path):
    """StoppingLoop keeps a chunk if ANY rank holds more than threshold / world x that rank's OWN largest value in it (no
    exchange for the bound).  The synthetic grids are the case that loses voxels if the division by world is forgotten (the
    hottest voxel of the sum is nobody's hot spot: checked by mutating the rule); the criterion must equal the evaluation on
    the whole summed grids."""
    out = str(tmp_path)
    mp.spawn(_selection_worker, args=(4, 29500 + ((os.getpid() + 977) % 2000), out), nprocs=4, join=True)
    current, exchanged, n_pad = np.load(os.path.join(out, "sel.npy"))
    s, q = np.load(os.path.join(out, "sel_sum.npy")), np.load(os.path.join(out, "sel_sq.npy"))
    n = 2000
    ss, c, mx = numpy_criterion(torch.from_numpy(s), torch.from_numpy(q), n, float((s / n).max()))
    assert c > 0 and exchanged < n_pad
    np.testing.assert_allclose(current, 100.0 * ss / c, rtol=1e-10)
