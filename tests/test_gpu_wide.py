"""GPU (-m gpu): wide candidate sets -- BASELINE configs[3] (2,048 variants, 5 %: blk_kernel, csrc/tvs_blk.cuh) and
configs[4] (4,096 variants, 30 %: the dense contraction) at their real V with R scaled down (the kernel is chosen by V
and the density, not by R) -- against the oracle (oracle/loglik.py, oracle/exact.py), through the C ABI.

Reference arithmetic: PathMultinomialModel.get_like_formula (ModelGenerator.py:153-174); minimize_neg_likelihood
(ModelFitMaxLike.py:19-58); cover_all_observed_subpaths (ModelFitMaxLike.py:560-580) for the infeasible lane.
Tolerances (BASELINE.json north_star): log-likelihood 1e-9 relative, frequencies 1e-6 absolute.
"""
import numpy as np
import pytest

from oracle import exact, loglik
from traversome_b200 import _cabi, synthetic

pytestmark = pytest.mark.gpu

TOL_LL = 1e-9
TOL_THETA = 1e-6


@pytest.fixture(scope="module")
def gpu(lib_built):
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return True


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def lanes_with_subsets(wl, B, seed):
    """B bootstrap lanes: lane 0 all variants, odd lanes random 70 % subsets that still cover every observed read
    path, the LAST lane a subset that leaves an observed read path uncovered (reference objective = -inf)."""
    rng = np.random.default_rng(seed)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=seed + 1)
    A = wl.scipy_csr()
    mask = np.ones((wl.V, B), dtype=np.uint8)
    for b in range(1, B, 2):
        m = rng.random(wl.V) < 0.7
        uncovered = np.asarray(A[:, m].sum(axis=1)).ravel() == 0
        for r in np.nonzero(uncovered & (w[:, b] > 0))[0]:          # re-admit one variant of every uncovered row
            m[wl.col[wl.row_ptr[r]]] = True
        mask[:, b] = m
    r_obs = int(np.nonzero(w[:, B - 1] > 0)[0][0])
    mask[:, B - 1] = 1
    mask[wl.col[wl.row_ptr[r_obs]:wl.row_ptr[r_obs + 1]], B - 1] = 0
    return w, mask


def check_against_exact(wl, w, mask, out, lanes):
    A = wl.scipy_csr()
    for b in lanes:
        r = exact.exact_fit(A, wl.L, w[:, b], mask=mask[:, b].astype(bool))
        assert out["status"][b] in (_cabi.CONVERGED, _cabi.STALLED), (b, out["status"][b], out["kkt"][b])
        assert np.abs(out["theta"][:, b] - r["theta"]).max() < TOL_THETA, (b, np.abs(out["theta"][:, b] - r["theta"]).max())
        assert rel(out["loglike"][b], r["loglike_no_C"]) < TOL_LL
        assert abs(out["theta"][:, b].sum() - 1.0) < 1e-12
        assert np.all(out["theta"][mask[:, b] == 0, b] == 0.0)


@pytest.mark.parametrize("B", [40, 37, 2, 129])
def test_wide_sparse_loglik_matches_oracle(gpu, B):
    """V = 2,048 at 5 %: even lane counts run 2 lanes per thread with bulk (TMA) operand copies, odd ones the 1-lane
    cp.async variant; 129 lanes = three tiles, the last with one lane."""
    wl = synthetic.make_workload(2048, 3000, 0.05, seed=41)
    A = wl.scipy_csr()
    rng = np.random.default_rng(B)
    w, mask = lanes_with_subsets(wl, B, seed=5 + B)
    theta = rng.dirichlet(np.ones(wl.V), size=B).T.copy()
    C = rng.normal(size=B) * 100
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes(w, mask=mask, C=C)
        got = eng.loglik(theta)
    for b in sorted(set([0, 1, B // 2, B - 2, B - 1])):
        if b < 0:
            continue
        ref = loglik.loglik_rows(A, wl.L, w[:, b], C[b], theta[:, b], mask[:, b].astype(bool))
        if np.isinf(ref):
            assert np.isinf(got[b]) and got[b] < 0
        else:
            assert rel(got[b], ref) < 1e-12, (b, got[b], ref)
    assert np.isinf(got[B - 1]) and got[B - 1] < 0


@pytest.mark.parametrize("B", [12, 7])
def test_wide_sparse_fit_matches_exact_optimum(gpu, B):
    """cfg4's code path (V = 2,048 -> blk_kernel + L-BFGS): >= 4 lanes against the exact optimum, incl. masked subsets
    and a lane whose subset does not cover an observed read path."""
    wl = synthetic.make_workload(2048, 12000, 0.05, seed=43)
    w, mask = lanes_with_subsets(wl, B, seed=60 + B)
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes(w, mask=mask)
        out = eng.fit(max_iter=3000)
    assert out["status"][B - 1] == _cabi.INFEASIBLE and np.isinf(out["loglike"][B - 1])
    assert np.all(out["status"][:B - 1] == _cabi.CONVERGED), np.bincount(out["status"], minlength=4)
    check_against_exact(wl, w, mask, out, [0, 1, 2, 3, B - 2])


def test_blk_kernel_equals_staged_kernel_where_both_apply(gpu, monkeypatch):
    """TVS_BLK=1 sends a shape the staged kernel (pass3) handles through blk_kernel: same fits (both converge to KKT 1e-9
    of the same optimum), log-likelihoods of the same theta equal to rounding.  Counts up to 65,535 exercise the split
    of a count into 16-bit-representable entries."""
    wl = synthetic.make_workload(200, 2500, 0.08, seed=47)
    rng = np.random.default_rng(3)
    big = rng.random(wl.val.size) < 0.1
    wl.val = np.where(big, rng.integers(1, 65536, size=wl.val.size), wl.val).astype(np.int32)
    B = 24
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=4)
    theta = rng.dirichlet(np.ones(wl.V), size=B).T.copy()
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("TVS_BLK", mode)
        with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
            eng.set_lanes(w)
            res[mode] = (eng.loglik(theta), eng.fit())
    ll0, ll1 = res["0"][0], res["1"][0]
    assert np.abs(ll1 - ll0).max() < 1e-12 * np.abs(ll0).max()
    f0, f1 = res["0"][1], res["1"][1]
    assert np.all(f0["status"] == _cabi.CONVERGED) and np.all(f1["status"] == _cabi.CONVERGED)
    assert np.abs(f0["theta"] - f1["theta"]).max() < 1e-7
    assert np.abs(f0["loglike"] - f1["loglike"]).max() < 1e-9 * np.abs(f0["loglike"]).max()


def test_wide_dense_matches_oracle_and_exact_optimum(gpu):
    """cfg5's code path (V = 4,096 at 30 % -> dense fp64 contraction): log-likelihood against the oracle, fits of >= 4
    lanes (masked subsets, an uncovered lane) against the exact optimum."""
    wl = synthetic.make_workload(4096, 5000, 0.30, seed=78)
    A = wl.scipy_csr()
    B = 8
    rng = np.random.default_rng(8)
    w, mask = lanes_with_subsets(wl, B, seed=70)
    theta = rng.dirichlet(np.ones(wl.V), size=B).T.copy()
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes(w, mask=mask)
        got = eng.loglik(theta)
        out = eng.fit(max_iter=3000)
    for b in (0, 1, B - 2):
        ref = loglik.loglik_rows(A, wl.L, w[:, b], 0.0, theta[:, b], mask[:, b].astype(bool))
        assert rel(got[b], ref) < 1e-11, (b, got[b], ref)
    assert np.isinf(got[B - 1]) and got[B - 1] < 0
    assert out["status"][B - 1] == _cabi.INFEASIBLE
    assert np.all(out["status"][:B - 1] == _cabi.CONVERGED), np.bincount(out["status"], minlength=4)
    check_against_exact(wl, w, mask, out, [0, 1, 2, B - 2])
