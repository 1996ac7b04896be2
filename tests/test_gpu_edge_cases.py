"""GPU edge cases of the transport path (run on the B200 box with -m gpu): the reference-compatible
explicit-vertex source, empty and ragged inputs, argument errors, and combinations of the multi-node
world with the sparse scorer and regions of interest.  Everything goes through the C ABI; the oracle is
the checker on identical Philox streams."""
import numpy as np
import pytest

import dose_metrics as M
import oracle_lib as O
from moquimc_b200 import capi

pytestmark = pytest.mark.gpu

NX, NY, NZ = 60, 50, 120


def edges():
    return capi.uniform_edges(-30, 30, NX), capi.uniform_edges(-25, 25, NY), capi.uniform_edges(-120, 0, NZ)


def water():
    return np.full(NX * NY * NZ, O.hu_to_density(np.array([0]))[0], dtype=np.float32)


def engine(physics=capi.PHYSICS_RELEASE, kinds=(capi.SCORER_DOSE,), capacity=0):
    e = capi.Engine(0, physics=physics)
    xe, ye, ze = edges()
    e.set_grid_hu(xe, ye, ze, np.zeros((NZ, NY, NX), dtype=np.int16))
    ids = [e.add_scorer(k, "s%d" % k, capacity) for k in kinds]
    return e, ids


def test_explicit_vertices_match_oracle_and_device_source():
    """mqi_set_vertices (upload_vertices + scorer_offset_vector, mqi_upload_data.hpp:415-468): vertices
    sampled by the caller.  The device sampler's vertices fed back explicitly give the oracle's dose on the
    same vertices, and per-spot ids key the Dij rows."""
    n_spots, per = 3, 1500
    n = n_spots * per
    bl = [capi.make_beamlet(100.0, [(s - 1) * 12.0, 0, 0.5, 0, 0, -1], [4, 4, 0, 0, 0, 0], uniform=True) for s in range(n_spots)]
    e, (s_dij, s_dose) = engine(kinds=(capi.SCORER_DIJ, capi.SCORER_DOSE), capacity=1_000_003)
    e.set_beamlets(bl, [per] * n_spots)
    v, spot = e.dev_sample_vertices(seed=21, first=0, n=n)
    assert np.array_equal(spot, np.repeat(np.arange(n_spots), per))
    e.set_vertices(v, spot)
    st = e.run(seed=21, first=0, count=n, per_spot=True)
    assert st.histories == n and st.dij_table_full == 0
    dense = e.get_dense(s_dose)
    k1, k2, val = e.get_sparse(s_dij)
    acc = np.zeros(dense.size)
    np.add.at(acc, k1, val)
    np.testing.assert_allclose(acc, dense.ravel(), rtol=1e-9, atol=1e-22)
    assert set(np.unique(k2)) == set(range(n_spots))
    # every spot's row lies around its own axis
    for s in range(n_spots):
        col = np.zeros(dense.size)
        np.add.at(col, k1[k2 == s], val[k2 == s])
        lat = col.reshape(NZ, NY, NX).sum(axis=(0, 1))
        assert abs((lat * (np.arange(NX) - 29.5)).sum() / lat.sum() - (s - 1) * 12.0) < 1.0
    xe, ye, ze = edges()
    g, keep = O.make_grid(xe, ye, ze, water())
    (od,), _ = O.transport(g, O.VARIANT_RELEASE, [], [], seed=21, h0=0, n=n, kinds=[O.SCORER_DOSE], vertices=v, spot_ids=spot)
    od = od.reshape(NZ, NY, NX)
    assert abs(dense.sum() / od.sum() - 1.0) < 5e-3
    gi, oi = dense.sum(axis=(1, 2)), od.sum(axis=(1, 2))
    assert np.abs(gi - oi).max() / oi.max() < 0.04
    # a sub-range of the vertex list: histories [first, first + count) of the launch
    e.clear_scorers()
    st = e.run(seed=21, first=per, count=per, per_spot=True)
    _, k2b, _ = e.get_sparse(s_dij)
    assert st.histories == per and set(np.unique(k2b)) == {1}
    with pytest.raises(capi.MqiError):
        e.run(seed=21, first=n - 10, count=11)


def test_empty_run_small_runs_and_argument_errors():
    e, (s,) = engine()
    b = capi.make_beamlet(80.0, [0, 0, 0.5, 0, 0, -1], [2, 2, 0, 0, 0, 0], uniform=True)
    with pytest.raises(capi.MqiError):
        e.run(1, 0, 10)                       # no source yet
    e.set_beamlets([b], [1000])
    st = e.run(1, 0, 0)                       # empty range: nothing launched, nothing scored
    assert st.histories == 0 and st.launches == 0 and e.get_dense(s).sum() == 0.0
    for count in (1, 31, 32, 33, 767, 769):   # around the warp / CTA sizes of the persistent grid
        e.clear_scorers()
        st = e.run(3, 0, count)
        assert st.histories == count
        assert e.get_dense(s).sum() > 0
    with pytest.raises(capi.MqiError):
        e.run(1, 995, 10)                     # beyond the beam source
    # histories that never enter the grid are counted as transported and score nothing
    away = capi.make_beamlet(80.0, [500.0, 0, 0.5, 0, 0, -1], [2, 2, 0, 0, 0, 0], uniform=True)
    e.set_beamlets([away], [500])
    e.clear_scorers()
    st = e.run(1, 0, 500)
    assert st.histories == 500 and e.get_dense(s).sum() == 0.0
    with pytest.raises(capi.MqiError):
        e.add_scorer(99, "bad")
    with pytest.raises(capi.MqiError):
        e.set_grid_hu(np.float32([0, 1, 1]), np.float32([0, 1]), np.float32([0, 1]), np.zeros(2, np.int16))   # edges must increase
    for _ in range(7):
        e.add_beamline_node(np.float32([-1, 1]), np.float32([-1, 1]), np.float32([10, 11]), np.float32([1e-8]))
    with pytest.raises(capi.MqiError):
        e.add_beamline_node(np.float32([-1, 1]), np.float32([-1, 1]), np.float32([10, 11]), np.float32([1e-8]))


def test_ragged_beamline_node_dij_and_roi_in_a_multi_node_world():
    """A range-shifter node with non-uniform edges (bisection path of the cell search), the sparse scorer
    keyed by spot and a CONTOUR roi on it, all in one multi-node launch, against the oracle."""
    n_spots, per = 2, 2500
    n = n_spots * per
    rng = np.random.default_rng(3)
    ze_rs = np.cumsum(np.concatenate([[30.0], rng.uniform(0.3, 4.0, size=12)])).astype(np.float32)   # ragged slab stack
    xe_rs = np.float32([-40, -7.5, -1.0, 0.25, 9.0, 40])
    ye_rs = np.float32([-40, 0.5, 40])
    rho_rs = rng.choice(np.float32([1.19e-3, 0.9e-3, 1e-8]), size=(len(ze_rs) - 1) * 2 * 5).astype(np.float32)
    mask = np.zeros((NZ, NY, NX), dtype=np.uint8)
    mask[20:110, 10:40, 5:50] = 1
    _, _, member = O.mask_to_roi(mask)
    bl = [capi.make_beamlet(110.0, [(s - 0.5) * 16.0, 0, 90.0, 0, 0, -1], [5, 5, 0, 0, 0, 0], uniform=True) for s in range(n_spots)]
    e, (s_dij, s_dose) = engine(kinds=(capi.SCORER_DIJ, capi.SCORER_DOSE), capacity=2_000_003)
    e.add_beamline_node(xe_rs, ye_rs, ze_rs, rho_rs)
    e.set_scorer_roi(s_dij, mask)
    e.set_beamlets(bl, [per] * n_spots)
    st = e.run(seed=9, first=0, count=n, per_spot=True)
    assert st.histories == n
    k1, k2, val = e.get_sparse(s_dij)
    dense = e.get_dense(s_dose)
    assert member[k1].all()                                   # the roi filters the sparse scorer ...
    acc = np.zeros(dense.size)
    np.add.at(acc, k1, val)
    inside = member.astype(bool)
    np.testing.assert_allclose(acc[inside], dense.ravel()[inside], rtol=1e-9, atol=1e-22)
    assert dense.ravel()[~inside].sum() > 0                   # ... and only it
    xe, ye, ze = edges()
    g_rs, k0 = O.make_grid(xe_rs, ye_rs, ze_rs, rho_rs)
    g, k1_ = O.make_grid(xe, ye, ze, water())
    ob = [O.make_beamlet(110.0, [(s - 0.5) * 16.0, 0, 90.0, 0, 0, -1], [5, 5, 0, 0, 0, 0], uniform=True) for s in range(n_spots)]
    (tab, od), _ = O.transport([g_rs, g], O.VARIANT_RELEASE, ob, [per] * n_spots, seed=9, h0=0, n=n,
                               kinds=[O.SCORER_DIJ, O.SCORER_DOSE], per_spot=True, dij_capacity=2_000_003,
                               roi_members=[member, None])
    assert abs(dense.sum() / od.sum() - 1.0) < 5e-3
    assert abs(val.sum() / tab["value"].sum() - 1.0) < 5e-3
    got = set(zip(k1.tolist(), k2.tolist()))
    exp = set(zip(tab["key1"].tolist(), tab["key2"].tolist()))
    assert len(got & exp) > 0.93 * max(len(got), len(exp))
    gi, oi = dense.reshape(NZ, NY, NX).sum(axis=(1, 2)), od.reshape(NZ, NY, NX).sum(axis=(1, 2))
    assert abs(M.r80_mm(gi) - M.r80_mm(oi)) < 0.3


def test_device_memory_reports_hbm_and_tracks_a_dij_table():
    e, _ = engine()
    free0, total = e.device_memory()
    assert 0 < free0 <= total and total > 100e9          # a B200 carries 180 GB
    e.add_scorer(capi.SCORER_DIJ, "Dij", 64_000_001)     # 16 B per slot = 1.02 GB, allocated with the first run
    e.set_beamlets([capi.make_beamlet(100.0, [0, 0, 0.5, 0, 0, -1], [2, 2, 0, 0, 0, 0], uniform=True)], [64])
    e.run(seed=1, first=0, count=64, per_spot=True)
    free1, _ = e.device_memory()
    assert free0 - free1 > 1.0e9


def test_dij_write_combining_changes_nothing_but_the_number_of_inserts():
    """Per-lane write-combining in front of the hash table (option dij_write_combine): same (voxel, spot) key set,
    same sums up to the order of the additions, with a second Dij scorer (inserted directly) and a dense Dose scorer
    alongside; rows still add up to the dense dose."""
    def run(wc):
        e, ids = engine(kinds=(capi.SCORER_DIJ, capi.SCORER_DOSE, capi.SCORER_DIJ), capacity=1_000_003)
        e.set_option("dij_write_combine", wc)
        bl = [capi.make_beamlet(90.0 + 15.0 * i, [-12.0 + 8.0 * i, 3.0, 0.5, 0, 0, -1], [3, 3, 0, 0.002, 0.002, 0], uniform=False)
              for i in range(4)]
        e.set_beamlets(bl, [3000] * 4)
        st = e.run(seed=5, first=0, count=12000, per_spot=True)
        assert st.histories == 12000 and st.dij_table_full == 0
        out = []
        for s in (ids[0], ids[2]):
            k1, k2, v = e.get_sparse(s)
            out.append({(int(a), int(b)): c for a, b, c in zip(k1, k2, v)})
        return out, e.get_dense(ids[1]).ravel()
    (a0, a2), dense_a = run(1)
    (b0, b2), dense_b = run(0)
    assert a0.keys() == b0.keys() == a2.keys() == b2.keys() and len(a0) > 10000
    kk = sorted(a0)
    ref = np.array([b0[k] for k in kk])
    for got in (a0, a2, b2):
        np.testing.assert_allclose(np.array([got[k] for k in kk]), ref, rtol=1e-9)
    np.testing.assert_allclose(dense_a, dense_b, rtol=1e-9, atol=1e-22)
    acc = np.zeros(dense_a.size)
    np.add.at(acc, np.array([k[0] for k in kk]), np.array([a0[k] for k in kk]))
    np.testing.assert_allclose(acc, dense_a, rtol=1e-9, atol=dense_a.max() * 1e-13)


@pytest.mark.parametrize("fetch_order", [0, 1])
def test_interleaved_shards_tile_the_history_range(fetch_order):
    """mqi_run_async_sharded: chunks of 32 histories dealt round-robin over n shards.  The shards together transport
    every history of the range exactly once with the streams of a single launch -- here all shards run one after the other on
    one device -- and each shard sees the same mix of spots (the point of interleaving: a plan sorted by energy layer).
    Both fetch orders (option fetch_order: a launch's chunks first to last, or last to first so that the longest histories
    of an ascending plan start first) transport the same histories: the reference dose is taken in forward order."""
    n_spots, per = 5, 1237          # a range that is not a multiple of 32 * shards
    n = n_spots * per
    bl = [capi.make_beamlet(70.0 + 20.0 * s, [(s - 2) * 6.0, 0, 0.5, 0, 0, -1], [3, 3, 0, 0, 0, 0], uniform=True) for s in range(n_spots)]
    e, (s_dose, s_dij) = engine(kinds=(capi.SCORER_DOSE, capi.SCORER_DIJ), capacity=2_000_003)
    e.set_beamlets(bl, [per] * n_spots)
    e.set_option("fetch_order", 0)
    st = e.run(seed=9, first=100, count=n - 200, per_spot=True)
    full, full_keys = e.get_dense(s_dose).copy(), set(zip(*e.get_sparse(s_dij)[:2]))
    assert st.histories == n - 200
    e.set_option("fetch_order", fetch_order)
    e.clear_scorers()
    st = e.run(seed=9, first=100, count=n - 200, per_spot=True)
    assert st.histories == n - 200
    np.testing.assert_allclose(e.get_dense(s_dose), full, rtol=1e-9, atol=full.max() * 1e-13)
    for shards in (2, 3, 8):
        e.clear_scorers()
        counts, rows = [], []
        for r in range(shards):
            st = e.run_sharded(9, 100, n - 200, shards, r, per_spot=True)
            counts.append(st.histories)
        assert sum(counts) == n - 200 and max(counts) - min(counts) <= 32
        np.testing.assert_allclose(e.get_dense(s_dose), full, rtol=1e-9, atol=full.max() * 1e-13)
        assert set(zip(*e.get_sparse(s_dij)[:2])) == full_keys
    # every shard holds every spot: a shard alone deposits in all five rows
    e.clear_scorers()
    e.run_sharded(9, 0, n, 4, 3, per_spot=True)
    _, k2, _ = e.get_sparse(s_dij)
    assert sorted(set(int(k) for k in k2)) == list(range(n_spots))
