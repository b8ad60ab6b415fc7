"""GPU (-m gpu): batches whose lane count is not a multiple of the pass kernels' lanes per thread.  tvs_fit re-lays such a
batch at the next multiple of 4 (padding lanes: zeros, finished from the start, no result) so that it runs the 2- and
4-lane kernels like any other batch: every lane must come out as it does in an even batch, its results must land at its
own position, and nothing of the padding may leak into the outputs.

Reference arithmetic: PathMultinomialModel.get_like_formula (ModelGenerator.py:153-174); minimize_neg_likelihood
(ModelFitMaxLike.py:19-58).  Tolerances (BASELINE.json north_star): log-likelihood 1e-9 relative, frequencies 1e-6 absolute.
"""
import os

import numpy as np
import pytest

from oracle import exact
from traversome_b200 import _cabi, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(lib_built):
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return True


def _lanes(wl, B, seed):
    rng = np.random.default_rng(seed)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=seed + 1)
    A = wl.scipy_csr()
    mask = np.ones((wl.V, B), dtype=np.uint8)
    for b in range(1, B, 3):                                    # every third lane: a subset that still covers its read paths
        m = rng.random(wl.V) < 0.7
        uncovered = np.asarray(A[:, m].sum(axis=1)).ravel() == 0
        for r in np.nonzero(uncovered & (w[:, b] > 0))[0]:
            m[wl.col[wl.row_ptr[r]]] = True
        mask[:, b] = m
    theta0 = rng.random((wl.V, B)) * mask
    theta0 /= theta0.sum(axis=0)
    C = rng.normal(size=B) * 10.0
    return w, mask, theta0, C


# (V, R, density, lane counts): 64 variants -> the staged kernels (pass3, up to 4 lanes per thread);
#                              384 variants -> blk_kernel (2 lanes per thread from 192 lanes on)
@pytest.mark.parametrize("V,R,dens,counts", [(64, 3000, 0.15, (7, 61, 203, 1001)), (384, 1500, 0.05, (9, 193, 451))])
def test_any_lane_count_gives_the_lanes_of_an_even_batch(gpu, V, R, dens, counts):
    wl = synthetic.make_workload(V, R, dens, seed=91)
    Bmax = max(counts) + 3
    Bmax += (-Bmax) % 4
    w, mask, theta0, C = _lanes(wl, Bmax, 5)
    A = wl.scipy_csr()
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes(w, mask=mask, C=C)
        even = eng.fit(theta0=theta0, max_iter=3000)
        assert np.all(np.isin(even["status"], (_cabi.CONVERGED, _cabi.STALLED)))
        for B in counts:
            for use_theta0 in (True, False):
                eng.set_lanes(np.ascontiguousarray(w[:, :B]), mask=np.ascontiguousarray(mask[:, :B]), C=C[:B])
                out = eng.fit(theta0=np.ascontiguousarray(theta0[:, :B]) if use_theta0 else None, max_iter=3000)
                assert out["theta"].shape == (V, B) and out["loglike"].shape == (B,)
                assert np.all(np.isin(out["status"], (_cabi.CONVERGED, _cabi.STALLED))), np.unique(out["status"])
                assert np.abs(out["theta"] - even["theta"][:, :B]).max() < 1e-6
                assert np.all(np.abs(out["loglike"] - even["loglike"][:B]) < 1e-9 * np.abs(even["loglike"][:B]))
                assert np.all(out["theta"][mask[:, :B] == 0] == 0.0) and np.all(np.abs(out["theta"].sum(axis=0) - 1.0) < 1e-12)
                assert np.all(out["iters"] > 0)
            for b in (0, 1, B - 1):                                   # the first, a masked and the LAST lane against the oracle
                r = exact.exact_fit(A, wl.L, w[:, b], mask=mask[:, b].astype(bool))
                assert np.abs(out["theta"][:, b] - r["theta"]).max() < 1e-6
                assert abs(out["loglike"][b] - (r["loglike_no_C"] + C[b])) < 1e-9 * abs(r["loglike_no_C"])
            ll = eng.loglik(out["theta"])                             # the evaluation entry point at the same odd width
            assert np.all(np.abs(ll - out["loglike"]) < 1e-9 * np.abs(out["loglike"]))


def test_padding_can_be_switched_off_and_changes_nothing_but_the_kernel(gpu):
    wl = synthetic.make_workload(384, 1500, 0.05, seed=92)
    w, mask, theta0, C = _lanes(wl, 201, 6)
    res = {}
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        for pad in ("1", "0"):
            os.environ["TVS_PAD"] = pad
            try:
                eng.set_lanes(w, mask=mask, C=C)
                res[pad] = eng.fit(max_iter=3000)
                res[pad]["lane_passes"] = eng.last_fit_stats()["lane_passes"]
            finally:
                del os.environ["TVS_PAD"]
    assert np.abs(res["1"]["theta"] - res["0"]["theta"]).max() < 1e-6
    assert np.all(np.abs(res["1"]["loglike"] - res["0"]["loglike"]) < 1e-9 * np.abs(res["0"]["loglike"]))
    assert np.array_equal(res["1"]["status"], res["0"]["status"])


def test_column_store_lanes_of_an_odd_batch(gpu):
    """The selection path: lanes name registered columns (16-bit mirror gathered on the device), odd count, masks."""
    wl = synthetic.make_workload(384, 1500, 0.05, seed=93)
    n_cols = 5
    Wc = synthetic.bootstrap_weights(wl.n_obs, n_cols, seed=4)
    B = 197
    rng = np.random.default_rng(2)
    slots = rng.integers(0, n_cols, B).astype(np.int32)
    mask = np.ones((wl.V, B), dtype=np.uint8)
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.columns_reserve(n_cols)
        eng.columns_upload(0, Wc)
        eng.set_lanes_columns(slots, mask=mask)
        got = eng.fit(max_iter=3000)
        assert np.array_equal(eng.get_weights(), Wc[:, slots])
        eng.set_lanes(np.ascontiguousarray(Wc[:, np.r_[slots, slots[-1:]]]), mask=np.ones((wl.V, B + 1), dtype=np.uint8))
        ref = eng.fit(max_iter=3000)
    assert np.abs(got["theta"] - ref["theta"][:, :B]).max() < 1e-6
    assert np.all(np.abs(got["loglike"] - ref["loglike"][:B]) < 1e-9 * np.abs(ref["loglike"][:B]))
    same = slots[:-1] == slots[1:]                                  # lanes of the same column: the same problem
    assert np.all(np.abs(got["loglike"][:-1][same] - got["loglike"][1:][same]) < 1e-9 * np.abs(got["loglike"][1:][same]))
