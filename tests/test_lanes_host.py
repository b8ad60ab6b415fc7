"""CPU: the host-side lane plumbing of traversome_b200.lanes -- array form of a batch, sorted id sets, the per-rank result
blocks of a gathered batch, the success rule for stalled lanes, column residency -- without a GPU (stand-in engine)."""
import numpy as np
import pytest

from tests.helpers import OracleEngine
from traversome_b200 import _cabi, synthetic
from traversome_b200.lanes import IdSet, LaneBatch, LaneBlock, LaneContext, LaneRequest, LaneSharder, LeaveOneOut, ShardedTheta


def test_lane_batch_from_requests_builds_masks_and_starts_in_one_scatter():
    V = 7
    reqs = [LaneRequest(3, 1.5, {0, 2, 5}), LaneRequest(4, 0.0, range(V), theta0=np.full(V, 1.0 / V)),
            LaneRequest(3, 2.5, IdSet(np.array([1, 6])), theta0=np.array([0.25, 0.75]))]
    b = LaneBatch.from_requests(reqs, V)
    assert b.cols.tolist() == [3, 4, 3] and b.C.tolist() == [1.5, 0.0, 2.5]
    want = np.zeros((V, 3), np.uint8)
    want[[0, 2, 5], 0] = 1
    want[:, 1] = 1
    want[[1, 6], 2] = 1
    assert np.array_equal(b.mask, want)
    assert np.allclose(b.theta0[:, 0], want[:, 0] / 3.0)                 # no start given: uniform over the subset
    assert np.allclose(b.theta0[:, 1], 1.0 / V) and b.theta0[1, 2] == 0.25 and b.theta0[6, 2] == 0.75
    full = LaneBatch.from_requests([LaneRequest(0, 0.0, range(V)), LaneRequest(1, 0.0, range(V))], V)
    assert full.mask is None and full.theta0 is None                      # every variant in every lane: no mask at all
    s = b.slice(1, 3)
    assert len(s) == 2 and s.cols.tolist() == [4, 3] and np.array_equal(s.mask, want[:, 1:3])
    assert b.digest() == LaneBatch.from_requests(reqs, V).digest() != b.slice(0, 2).digest()


def test_id_set_behaves_like_a_sorted_set():
    s = IdSet(np.array([2, 5, 9], dtype=np.int64))
    assert list(s) == [2, 5, 9] and len(s) == 3 and 5 in s and 4 not in s and 10 not in s
    assert LaneRequest(0, 0.0, s).ids is s.arr                            # no copy, no sort
    assert LaneRequest(0, 0.0, {9, 2, 5}).ids == [2, 5, 9]


def _start_like_the_fitter(warm, ids):
    """ModelFitMaxLike._start_point for one subset (the per-lane form the array form must reproduce)."""
    x = warm[ids]
    tot = x.sum()
    if not np.isfinite(tot) or tot <= 0.0:
        return None
    x = np.maximum(x / tot, 1e-3 / len(ids))
    return x / x.sum()


@pytest.mark.parametrize("warm_kind", ["optimum", "none", "degenerate"])
def test_leave_one_out_block_equals_the_per_lane_requests(warm_kind):
    V = 11
    rng = np.random.default_rng(5)
    chosen = np.array([0, 2, 3, 6, 7, 10], dtype=np.int64)
    lo = LeaveOneOut(chosen, [4, 0, 5, 2])
    assert len(lo) == 4 and [list(s) for s in lo] == [[0, 2, 3, 6, 10], [2, 3, 6, 7, 10], [0, 2, 3, 6, 7], [0, 2, 6, 7, 10]]
    warm = None
    if warm_kind != "none":
        warm = np.zeros(V)
        warm[chosen] = rng.dirichlet(np.ones(len(chosen)))
        warm[3] = 0.0                                                     # an exact zero in the accepted optimum: floored
        if warm_kind == "degenerate":                                     # all the mass on the variant lane 0 leaves out
            warm[:] = 0.0
            warm[7] = 1.0
    cols, C, mask, theta0 = LaneBlock(2, -3.5, lo, warm).arrays(V)
    reqs = [LaneRequest(2, -3.5, ids, None if warm is None else _start_like_the_fitter(warm, ids.arr)) for ids in lo]
    want = LaneBatch.from_requests(reqs, V)
    assert cols.tolist() == want.cols.tolist() and C.tolist() == want.C.tolist() and np.array_equal(mask, want.mask)
    if warm is None:
        assert theta0 is None
    else:
        assert np.allclose(theta0, want.theta0, rtol=0, atol=1e-15) and np.allclose(theta0.sum(axis=0), 1.0)
        assert np.all(theta0[mask == 0] == 0.0)


def test_fit_items_mixes_blocks_and_request_lists_in_one_batch():
    wl = synthetic.make_workload(6, 200, 0.8, seed=4)
    A = wl.scipy_csr().toarray() > 0
    ok_drop = [v for v in range(wl.V) if np.all(np.delete(A, v, axis=1).any(axis=1))]      # still covers every read path
    assert len(ok_drop) >= 2
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    ctx.engine_factory = OracleEngine
    c0 = ctx.add_weights(wl.n_obs.astype(np.int64))
    chosen = np.arange(wl.V, dtype=np.int64)
    lo = LeaveOneOut(chosen, ok_drop[:2])
    full = [LaneRequest(c0, 2.0, range(wl.V))]
    res_full, res_blk = ctx.fit_items([full, LaneBlock(c0, 2.0, lo, None)])
    assert ctx.n_batches == 1 and ctx.n_lanes_fitted == 3 and len(res_full) == 1 and len(res_blk) == 2
    one_by_one = ctx.fit([LaneRequest(c0, 2.0, ids) for ids in lo])
    for j in range(2):
        assert abs(res_blk.fun[j] - one_by_one[j].fun) <= 1e-9 * abs(one_by_one[j].fun) and bool(res_blk.success[j])
        assert np.allclose(res_blk.x(j), one_by_one[j].x, atol=1e-7) and np.allclose(res_blk[j].x, res_blk.x(j))
        assert res_blk.ids(j).tolist() == list(lo[j])
    assert res_full[0].fun <= res_blk.fun.min() + 1e-9


def test_pipelined_batches_equal_one_by_one():
    wl = synthetic.make_workload(6, 200, 0.8, seed=4)
    W = synthetic.bootstrap_weights(wl.n_obs, 10, seed=2)
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    ctx.engine_factory = OracleEngine
    cols = [ctx.add_weights(W[:, b]) for b in range(10)]
    batches = [LaneBatch(np.array(cols[0:4])), LaneBatch(np.array(cols[4:5])), LaneBatch(np.array(cols[5:10]), C=np.arange(5.0)),
               LaneBatch(np.array(cols[2:6])), LaneBatch(np.array(cols[9:10]))]
    one = [ctx.fit_batch(b) for b in batches]
    n0, l0 = ctx.n_batches, ctx.n_lanes_fitted
    two = ctx.fit_batches(batches)
    assert ctx.n_batches - n0 == 5 and ctx.n_lanes_fitted - l0 == 15 and ctx._engine2 is not None
    for a, b in zip(one, two):
        for k in ("theta", "loglike", "status"):
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]))
    ws = [np.ascontiguousarray(W[:, 0:3]), np.ascontiguousarray(W[:, 3:4]), np.ascontiguousarray(W[:, 4:10])]
    got = ctx.fit_weight_batches(ws, C=[np.zeros(3), np.ones(1), None])
    for w, g, c in zip(ws, got, (0.0, 1.0, 0.0)):
        ref = ctx.fit_weights(w)
        assert np.allclose(g["loglike"], ref["loglike"] + c, rtol=1e-12, atol=0) and np.array_equal(g["theta"], ref["theta"])
    ctx.pipeline_batches = False
    assert all(np.array_equal(a["loglike"], b["loglike"]) for a, b in zip(one, ctx.fit_batches(batches)))

    class Boom(OracleEngine):
        def fit(self, *a, **k):
            raise RuntimeError("boom")
    ctx2 = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    ctx2.engine_factory = Boom
    c2 = [ctx2.add_weights(W[:, b]) for b in range(4)]
    with pytest.raises(RuntimeError):
        ctx2.fit_batches([LaneBatch(np.array(c2[:2])), LaneBatch(np.array(c2[2:])), LaneBatch(np.array(c2[:1]))])


def test_sharded_theta_indexing_and_materialisation():
    rng = np.random.default_rng(0)
    blocks = [rng.random((4, 3)), rng.random((4, 2)), rng.random((4, 3))]
    full = np.concatenate(blocks, axis=1)
    th = ShardedTheta(blocks, [0, 3, 5], 4, 8)
    assert th.shape == (4, 8)
    for b in range(8):
        assert np.array_equal(th[:, b], full[:, b]) and np.array_equal(th[[1, 3], b], full[[1, 3], b])
    assert np.array_equal(th[np.array([0, 2]), -1], full[[0, 2], -1])
    assert np.array_equal(np.asarray(th), full) and np.array_equal(th[:, 2:6], full[:, 2:6])
    assert np.array_equal(th - full, np.zeros_like(full)) and np.array_equal(th * 2.0, full * 2.0) and np.array_equal(th.T, full.T)


def test_stalled_lanes_succeed_only_near_the_optimum():
    ctx = LaneContext(3, [100, 100, 100])
    out = {"status": np.array([_cabi.CONVERGED, _cabi.STALLED, _cabi.STALLED, _cabi.MAXITER, _cabi.CONVERGED]),
           "kkt": np.array([[1e-10, 1e-8], [5e-7, 5e-5], [1e-3, 1e-2], [1e-10, 1e-9], [1e-10, 1e-9]]),
           "loglike": np.array([-1.0, -2.0, -3.0, -4.0, -np.inf])}
    assert ctx.lane_ok(out).tolist() == [True, True, False, False, False]


def test_columns_are_uploaded_once_and_named_by_slot():
    wl = synthetic.make_workload(5, 120, 0.5, seed=9)
    W = synthetic.bootstrap_weights(wl.n_obs, 6, seed=2)
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    ctx.engine_factory = OracleEngine
    cols = [ctx.add_weights(W[:, b]) for b in range(6)]
    with pytest.raises(ValueError):
        ctx.add_weights(np.array([1, -1] + [0] * (wl.R - 2)))
    r1 = ctx.fit([LaneRequest(cols[4], 0.0, range(wl.V)), LaneRequest(cols[1], 0.5, range(wl.V)), LaneRequest(cols[1], 0.0, {0})])
    eng = ctx.engine()
    res = ctx._resident[id(eng)]
    assert sorted(res.slot_of) == [cols[1], cols[4]] and res.used == 2     # only what the batch needed
    r2 = ctx.fit([LaneRequest(cols[1], 0.5, range(wl.V)), LaneRequest(cols[2], 0.0, range(wl.V))])
    assert res.used == 3 and np.array_equal(eng.cols[:, res.slot_of[cols[2]]], W[:, 2])
    assert np.array_equal(r1[1].x, r2[0].x) and r1[1].fun == r2[0].fun     # the same lane in another batch
    assert r1[0].success and len(r1[0].x) == wl.V and len(r1[2].x) == 1 and not r1[2].success    # one variant cannot cover every read path
    assert LaneSharder.bounds(10, 0, 4) == (0, 3) and LaneSharder.bounds(10, 3, 4) == (8, 10)
