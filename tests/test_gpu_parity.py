"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): log-likelihood 1e-9 relative, fitted frequencies 1e-6
absolute, selected models identical.  Integer outputs (resampled weights) bit-exact.
"""
import numpy as np
import pytest

from oracle import exact, loglik
from tests.helpers import fit_context, model_from_tables, resample_reference
from traversome_b200 import _cabi, csr, synthetic

pytestmark = pytest.mark.gpu

TOL_LL = 1e-9
TOL_THETA = 1e-6


@pytest.fixture(scope="module")
def gpu(lib_built):
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return True


def engine_for(wl):
    return _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


# --------------------------------------------------------------------------------------- likelihood
@pytest.mark.parametrize("V,R,density,B", [(2, 40, 0.7, 1), (16, 700, 0.2, 37), (64, 3000, 0.1, 96), (300, 900, 0.05, 40)])
def test_loglik_matches_oracle(gpu, V, R, density, B):
    wl = synthetic.make_workload(V, R, density, seed=11 + V)
    A = wl.scipy_csr()
    rng = np.random.default_rng(V)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=3)
    theta = rng.dirichlet(np.ones(V), size=B).T.copy()
    mask = (rng.random((V, B)) < 0.8).astype(np.uint8)
    mask[:, 0] = 1
    C = rng.normal(size=B) * 100
    with engine_for(wl) as eng:
        eng.set_lanes(w, mask=mask, C=C)
        got = eng.loglik(theta)
    for b in range(B):
        ref = loglik.loglik_rows(A, wl.L, w[:, b], C[b], theta[:, b], mask[:, b].astype(bool))
        if np.isinf(ref):
            assert np.isinf(got[b]) and got[b] < 0
        else:
            assert rel(got[b], ref) < 1e-12, (b, got[b], ref)


@pytest.mark.parametrize("counts", [(1, 2, 3, 40, 64), (1, 33, 47, 1000, 65535)])
def test_loglik_and_fit_with_large_occurrence_counts(gpu, counts):
    """Counts that are exact in the top 16 bits of an fp64 (1..32, 40, 64: pass3_kernel's packed entries) and counts that
    are not (33, 47, 1000, 65535: the handle falls back to the unstaged kernel; blk_kernel splits such a count into
    several entries).  Reference: from_variants counts are arbitrary positive integers (utils.py:416-501)."""
    wl = synthetic.make_workload(48, 1500, 0.15, seed=77)
    rng = np.random.default_rng(5)
    val = wl.val.copy()
    val[:] = rng.choice(np.asarray(counts, dtype=val.dtype), size=val.size)
    B = 12
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=9)
    theta = rng.dirichlet(np.ones(48), size=B).T.copy()
    import scipy.sparse as sp
    A = sp.csr_matrix((val.astype(np.float64), wl.col, wl.row_ptr), shape=(wl.row_ptr.size - 1, 48))
    with _cabi.Engine(wl.row_ptr, wl.col, val, wl.L) as eng:
        eng.set_lanes(w)
        got = eng.loglik(theta)
        for b in range(B):
            ref = loglik.loglik_rows(A, wl.L, w[:, b], 0.0, theta[:, b], np.ones(48, dtype=bool))
            assert rel(got[b], ref) < 1e-12, (b, got[b], ref)
        out = eng.fit()
    assert (out["status"] == 0).all()
    for b in (0, B - 1):
        ref = exact.exact_fit(A, wl.L, w[:, b])
        assert np.abs(out["theta"][:, b] - ref["theta"]).max() < TOL_THETA
        assert rel(out["loglike"][b], ref["loglike_no_C"]) < TOL_LL


def test_loglik_known_answers_from_reference(gpu, plastome, synthetic_small):
    cases = [(plastome["variant_sizes"], plastome["models"][0]["bins_list"], plastome["known_answer_ll"])]
    cases += [(c["variant_sizes"], c["bins_list"], c["known_answer_ll"]) for c in synthetic_small["cases"]]
    for L, bins, kas in cases:
        model = model_from_tables(L, bins)
        matrix, w, C, N = csr.lanes_from_models([model])
        row_ptr, col, val = matrix.arrays()
        with _cabi.Engine(row_ptr, col, val, np.asarray(L, np.int64)) as eng:
            for ka in kas:
                theta = np.array([float.fromhex(t) for t in ka["theta"]])[:, None]
                mask = np.ones((len(L), 1), np.uint8)
                if ka["subset"]:
                    mask[:] = 0
                    mask[ka["subset"], 0] = 1
                eng.set_lanes(w, mask=mask, C=C)
                got = float(eng.loglik(theta)[0])
                ref = float.fromhex(ka["loglike"])
                if np.isinf(ref):
                    assert np.isinf(got) and got < 0
                else:
                    assert rel(got, ref) < TOL_LL


# ---------------------------------------------------------------------------------------------- fit
def check_fit_against_exact(wl, w, mask, out, lanes):
    A = wl.scipy_csr()
    for b in lanes:
        m = None if mask is None else mask[:, b].astype(bool)
        r = exact.exact_fit(A, wl.L, w[:, b], mask=m)
        assert out["status"][b] in (_cabi.CONVERGED, _cabi.STALLED), (b, out["status"][b], out["kkt"][b])
        assert np.abs(out["theta"][:, b] - r["theta"]).max() < TOL_THETA, (b, np.abs(out["theta"][:, b] - r["theta"]).max())
        assert rel(out["loglike"][b], r["loglike_no_C"]) < TOL_LL
        assert abs(out["theta"][:, b].sum() - 1.0) < 1e-12
        if m is not None:
            assert np.all(out["theta"][~m, b] == 0.0)


@pytest.mark.parametrize("V,R,density,B,n_true", [(2, 60, 0.8, 5, None), (8, 400, 0.3, 33, 4), (16, 1500, 0.2, 70, None),
                                                   (64, 4000, 0.1, 130, None), (64, 4000, 0.1, 64, 8)])
def test_fit_matches_exact_optimum(gpu, V, R, density, B, n_true):
    wl = synthetic.make_workload(V, R, density, seed=100 + V + B, n_true=n_true)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=9)
    with engine_for(wl) as eng:
        eng.set_lanes(w)
        out = eng.fit()
    lanes = sorted(set([0, 1, B // 2, B - 1]))
    check_fit_against_exact(wl, w, None, out, lanes)
    assert np.all(out["status"] == _cabi.CONVERGED)


@pytest.mark.parametrize("V,R,density,B", [(96, 3000, 0.08, 40), (200, 2500, 0.05, 24)])
def test_fit_medium_and_large_variant_counts(gpu, V, R, density, B):
    """V = 96: the Newton phase with the 128-variant solve kernel; V = 200: beyond its limit (tvs_newton.cuh), the
    batched L-BFGS alone runs every lane to the KKT certificate."""
    wl = synthetic.make_workload(V, R, density, seed=300 + V)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=4)
    with engine_for(wl) as eng:
        eng.set_lanes(w)
        out = eng.fit()
    assert np.all(out["status"] == _cabi.CONVERGED)
    check_fit_against_exact(wl, w, None, out, [0, 1, B - 1])


def test_dense_contraction_pass_equals_sparse_pass_and_exact_optimum(gpu, monkeypatch):
    """Dense shapes (V > 400, >= 22 % of A non-zero: BASELINE configs[4]) run the likelihood pass as two fp64 GEMMs on a
    dense copy of A (tvs_api.cu: launch_dense).  Same log-likelihoods as the sparse kernels to fp64 round-off, the
    same optimum, subset masks and finished lanes honoured; TVS_DENSE=0 keeps the sparse path."""
    V, B = 448, 24
    wl = synthetic.make_workload(V, 1800, 0.3, seed=812)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=3)
    mask = np.ones((V, B), np.uint8)
    mask[5:40, 3] = 0
    mask[::2, 7] = 0
    theta = np.random.default_rng(4).dirichlet(np.ones(V), size=B).T * mask
    theta /= theta.sum(axis=0)
    res = {}
    for mode in ("0", "auto"):
        if mode == "0":
            monkeypatch.setenv("TVS_DENSE", "0")
        else:
            monkeypatch.delenv("TVS_DENSE")
        with engine_for(wl) as eng:
            eng.set_lanes(w, mask=mask)
            ll = eng.loglik(theta)
            out = eng.fit()
            res[mode] = (ll, out, eng.last_fit_stats()["kernel_launches"])
    (ll_s, out_s, _), (ll_d, out_d, _) = res["0"], res["auto"]
    assert np.all(np.abs(ll_s - ll_d) <= 1e-13 * np.abs(ll_s))
    assert np.all(out_s["status"] == _cabi.CONVERGED) and np.all(out_d["status"] == _cabi.CONVERGED)
    assert np.all(np.abs(out_s["loglike"] - out_d["loglike"]) <= 1e-11 * np.abs(out_s["loglike"]))
    assert np.abs(out_s["theta"] - out_d["theta"]).max() < 1e-6
    assert np.all(out_d["theta"][mask == 0] == 0.0)
    check_fit_against_exact(wl, w, mask, out_d, [0, 3, 7, B - 1])
    # a single lane, and a lane whose subset leaves a read path uncovered, on the dense path
    monkeypatch.delenv("TVS_DENSE", raising=False)
    A = wl.dense()
    keep = np.zeros(V, bool)
    keep[:3] = True
    uncovered = bool(np.any((A[:, keep].sum(1) == 0) & (w[:, 0] > 0)))
    assert uncovered, "three variants should not cover every read path of this fixture"
    m2 = np.ones((V, 2), np.uint8)
    m2[:, 1] = keep
    with engine_for(wl) as eng:
        eng.set_lanes(w[:, :2], mask=m2)
        o2 = eng.fit()
        assert o2["status"][0] == _cabi.CONVERGED and o2["status"][1] == _cabi.INFEASIBLE and np.isinf(o2["loglike"][1])
        eng.set_lanes(w[:, :1])
        o1 = eng.fit()
    assert o1["status"][0] == _cabi.CONVERGED
    assert np.abs(o1["theta"][:, 0] - o2["theta"][:, 0]).max() < 1e-7
    # a sparse matrix of the same width does not take the dense path (density below the threshold): both settings agree bitwise
    wl2 = synthetic.make_workload(V, 1500, 0.04, seed=813)
    w2 = synthetic.bootstrap_weights(wl2.n_obs, 4, seed=3)
    outs = []
    for mode in ("0", "auto"):
        monkeypatch.setenv("TVS_DENSE", "0") if mode == "0" else monkeypatch.delenv("TVS_DENSE")
        with engine_for(wl2) as eng:
            eng.set_lanes(w2)
            outs.append(eng.fit(max_iter=30)["theta"])
    assert np.array_equal(outs[0], outs[1])


def test_fit_with_subset_masks_and_infeasible_lane(gpu):
    V, B = 12, 20
    wl = synthetic.make_workload(V, 800, 0.35, seed=77)
    rng = np.random.default_rng(1)
    w = np.repeat(wl.n_obs[:, None], B, axis=1).astype(np.int32)
    mask = np.ones((V, B), np.uint8)
    A = wl.dense()
    for b in range(1, B):
        mask[(b - 1) % V, b] = 0                       # leave-one-out lanes
    covered = [bool(np.all((A[:, mask[:, b].astype(bool)].sum(1) > 0) | (w[:, b] == 0))) for b in range(B)]
    # force an uncovered lane: keep a single variant
    mask[:, B - 1] = 0
    mask[0, B - 1] = 1
    covered[B - 1] = bool(np.all((A[:, [0]].sum(1) > 0)))
    with engine_for(wl) as eng:
        eng.set_lanes(w, mask=mask)
        out = eng.fit()
    good = [b for b in range(B) if covered[b]]
    bad = [b for b in range(B) if not covered[b]]
    assert bad, "fixture should contain an uncovered lane"
    check_fit_against_exact(wl, w, mask, out, good)
    for b in bad:
        assert out["status"][b] == _cabi.INFEASIBLE and np.isinf(out["loglike"][b])


def test_fit_plastome_full_data_and_bootstraps(gpu, plastome):
    """cfg1: GPU optimum == exact optimum (23/35, 12/35); reference SLSQP answers lie within their own
    documented spread of it; LL within 1e-9 relative of the reference's."""
    L = plastome["variant_sizes"]
    models = [model_from_tables(L, m["bins_list"]) for m in plastome["models"]]
    matrix, w, C, N = csr.lanes_from_models(models)
    row_ptr, col, val = matrix.arrays()
    with _cabi.Engine(row_ptr, col, val, np.asarray(L, np.int64)) as eng:
        eng.set_lanes(w, C=C)
        out = eng.fit()
    assert np.all(out["status"] == _cabi.CONVERGED)
    assert abs(out["theta"][0, 0] - 23 / 35.) < 1e-7
    for fit in plastome["fits"]:
        b = fit["model"]
        ref = np.array([p for _, p in fit["prop"]])
        assert rel(out["loglike"][b], fit["loglike"]) < TOL_LL
        assert out["loglike"][b] >= fit["loglike"] - 1e-8
        assert np.abs(out["theta"][:, b] - ref).max() < 5e-4          # reference's own SLSQP tolerance
        A, wr, Cr, Nr = loglik.rows_from_bins(L, plastome["models"][b]["bins_list"])
        r = exact.exact_fit(A, L, wr)
        assert np.abs(out["theta"][:, b] - r["theta"]).max() < TOL_THETA
        # bootstraps.replicates.tab prints %.4f of the proportions
    tab = plastome["output_tables"]["bootstraps.replicates.tab"].strip().split("\n")[1:]
    n_round_differs = 0
    for line, b in zip(tab, range(1, len(models))):
        ref = [float(x) for x in line.split("\t")[1:]]
        got = [round(float(out["theta"][v, b]), 4) for v in range(2)]
        n_round_differs += int(any(abs(g - r_) > 1.01e-4 for g, r_ in zip(got, ref)))
    assert n_round_differs == 0


def test_fit_is_deterministic_and_warm_start_agrees(gpu):
    wl = synthetic.make_workload(32, 2500, 0.15, seed=5)
    w = synthetic.bootstrap_weights(wl.n_obs, 50, seed=2)
    with engine_for(wl) as eng:
        eng.set_lanes(w)
        a = eng.fit()
        b = eng.fit()
        assert np.array_equal(a["theta"], b["theta"]) and np.array_equal(a["loglike"], b["loglike"])
        warm = eng.fit(theta0=np.repeat(a["theta"][:, :1], 50, axis=1))
    assert np.abs(warm["theta"] - a["theta"]).max() < 1e-7
    assert np.all(warm["status"] == _cabi.CONVERGED)
    assert warm["iters"][0] <= 30


@pytest.mark.parametrize("V,B", [(8, 5), (40, 200), (150, 12)])
def test_fit_escapes_the_zero_saddle(gpu, V, B):
    """z_v = 0 is a stationary point of the z-parametrisation for every variant.  Started from a one-hot theta0 (all
    other variants at the floor) the solver must still reach the optimum, where those variants are positive: the
    multiplier test g_v <= 1 keeps such lanes alive and the saddle escape moves the components off zero.  Covers the
    Newton phase (small B, V <= 128), the L-BFGS phase (B > 128) and V > 128."""
    wl = synthetic.make_workload(V, 60 * V, 0.3, seed=900 + V)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=3)
    theta0 = np.zeros((V, B))
    theta0[0, :] = 1.0
    with engine_for(wl) as eng:
        eng.set_lanes(w)
        out = eng.fit(theta0=theta0)
    assert np.all(out["status"] == _cabi.CONVERGED), np.bincount(out["status"], minlength=4)
    check_fit_against_exact(wl, w, None, out, [0, B // 2, B - 1])


def test_edge_cases(gpu):
    # single variant, single row, zero-weight rows, lane count not a multiple of 32
    rp = np.array([0, 1], np.int32)
    with _cabi.Engine(rp, np.array([0], np.int32), np.array([2], np.int32), np.array([1000], np.int64)) as eng:
        eng.set_lanes(np.array([[7]], np.int32))
        out = eng.fit()
        assert out["theta"][0, 0] == 1.0 and out["status"][0] == _cabi.CONVERGED
        assert rel(out["loglike"][0], 7 * np.log(2.0 / 1000.0)) < 1e-13
    wl = synthetic.make_workload(6, 200, 0.4, seed=8)
    w = synthetic.bootstrap_weights(wl.n_obs, 3, seed=1)
    w[::2, 1] = 0                                        # many empty rows in lane 1
    with engine_for(wl) as eng:
        eng.set_lanes(w)
        out = eng.fit()
    check_fit_against_exact(wl, w, None, out, [0, 1, 2])
    with pytest.raises(_cabi.TvsError):
        with engine_for(wl) as eng:
            eng.fit()                                    # lanes not set -> TVS_ESTATE


def test_device_log_and_reciprocal_are_within_2_ulp(gpu):
    """The pass kernel replaces IEEE division and libdevice log by MUFU-seeded Newton / an atanh series."""
    rng = np.random.default_rng(0)
    p = np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), 10.0 ** rng.uniform(-12, 0, 400000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 100000), 2.0 ** rng.integers(-1000, 1000, 1000).astype(np.float64),
                        np.nextafter(2.0 ** rng.integers(-50, 50, 1000).astype(np.float64), 0.0),
                        np.array([np.sqrt(2.0), np.sqrt(0.5), 1.0, 0.5, 2.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308])])
    lg, rc = _cabi.selftest_math(p)
    ref_l = np.log(p)
    ref_r = 1.0 / p
    ulp_l = np.abs(lg - ref_l) / np.maximum(np.spacing(np.abs(ref_l)), 5e-324)
    # near p = 1 the result is tiny: the contract is ABSOLUTE accuracy (terms of a sum of magnitude >= 1e3)
    bad_l = (ulp_l > 2) & (np.abs(lg - ref_l) > 2.3e-16)
    assert not bad_l.any(), (p[bad_l][:5], lg[bad_l][:5], ref_l[bad_l][:5])
    ulp_r = np.abs(rc - ref_r) / np.spacing(np.abs(ref_r))
    fin = np.isfinite(ref_r) & (np.abs(ref_r) > 1e-300)
    assert ulp_r[fin].max() <= 2
    wl = synthetic.make_workload(4, 50, 0.5, seed=1)
    with engine_for(wl) as eng:
        lt = eng.selftest_table_log(p)
    ulp_t = np.abs(lt - ref_l) / np.maximum(np.spacing(np.abs(ref_l)), 5e-324)
    bad_t = (ulp_t > 2) & (np.abs(lt - ref_l) > 3.4e-16)
    assert not bad_t.any(), (p[bad_t][:5], lt[bad_t][:5], ref_l[bad_t][:5])
    z = _cabi.selftest_math(np.array([0.0, -1.0, np.inf]))
    assert z[0][0] == -np.inf and np.isnan(z[0][1]) and z[0][2] == np.inf and z[1][0] == np.inf


# --------------------------------------------------------------------------------------- resampling
def test_device_resample_is_bit_exact_vs_numpy_philox(gpu):
    wl = synthetic.make_workload(8, 300, 0.3, seed=21)
    with engine_for(wl) as eng:
        eng.resample(6, wl.n_obs, seed=0x1234567890ABCDEF, keep_lane0=True)
        got = eng.get_weights()
    ref = resample_reference(wl.n_obs, 6, 0x1234567890ABCDEF, True)
    assert np.array_equal(got, ref)
    assert np.all(got.sum(0) == wl.N)


# ------------------------------------------------------------------- full-size, size-independent props
def test_cfg2_full_size_properties(gpu):
    """BASELINE.json configs[1] at full size: every lane converges, KKT certified, lane 0 (unresampled)
    equals the exact optimum, refitting from the answer is a fixed point, and a lane's LL at its own
    optimum is never below its LL at another lane's optimum."""
    wl = synthetic.make_config("cfg2")
    B = 1000
    with engine_for(wl) as eng:
        eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
        out = eng.fit()
        assert np.all(out["status"] == _cabi.CONVERGED)
        assert out["kkt"][:, 0].max() <= 1e-9 and out["kkt"][:, 1].max() <= 1e-7
        assert np.abs(out["theta"].sum(0) - 1).max() < 1e-12
        w = eng.get_weights()
        assert np.all(w.sum(0) == wl.N) and np.array_equal(w[:, 0], wl.n_obs)
        ll_own = eng.loglik(out["theta"])
        assert np.abs(ll_own - out["loglike"]).max() < 1e-9 * np.abs(ll_own).max()
        shifted = np.roll(out["theta"], 1, axis=1)
        ll_other = eng.loglik(shifted)
        assert np.all(ll_own >= ll_other - 1e-9 * np.abs(ll_own))
        again = eng.fit(theta0=out["theta"], tol=1e-9)     # the answer is a fixed point: nothing left to do
        assert np.abs(again["theta"] - out["theta"]).max() < 1e-9 and again["iters"].max() <= 1
    check_fit_against_exact(wl, w, None, out, [0, 1, 999])
