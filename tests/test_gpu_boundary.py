"""GPU (-m gpu): the reference-facing boundary -- ``ModelFitMaxLike`` (point_estimate,
reverse_model_selection, get_neg_likelihood_of_var_freq) and the batched bootstrap driver -- running on
the CUDA solver through the C ABI, against the reference's own results (tests/golden) and the oracle.

north_star tolerances: LL 1e-9 relative and frequencies 1e-6 absolute against the exact optimum of the
reference's objective (oracle/exact.py; SURVEY.md 8c explains why the reference's raw SLSQP output, which
scatters by ~3e-4 between its own restarts, cannot be the 1e-6 target); selected models and the
bootstrap support table identical to the reference's.
"""
import math

import numpy as np
import pytest

from oracle import exact, loglik
from tests.helpers import OracleEngine, fit_context, hexf, model_from_tables, model_from_workload
from traversome_b200 import _cabi, bootstrap, synthetic
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike
from traversome_b200.lanes import LaneContext, LaneRequest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(lib_built):
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return True


def _selection(case, criterion):
    return [f for f in case["fits"] if f["kind"] == "reverse_model_selection" and f["criterion"] == criterion][0]


@pytest.mark.parametrize("criterion", ["BIC", "AIC"])
def test_reverse_model_selection_matches_reference(gpu, synthetic_small, criterion):
    for case in synthetic_small["cases"]:
        L = case["variant_sizes"]
        model = model_from_tables(L, case["bins_list"])
        kw = fit_context(model, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        fitter = ModelFitMaxLike(**kw)
        prop, like, crit = fitter.reverse_model_selection(n_proc=1, criterion=criterion)
        ref = _selection(case, criterion)
        assert sorted(prop) == [cid for cid, _ in ref["prop"]], case["name"]            # identical model
        for cid, p in ref["prop"]:
            assert abs(prop[cid] - p) < 5e-4                                             # reference's SLSQP spread
        assert abs(like - ref["loglike"]) < 1e-8 * abs(ref["loglike"])
        assert like >= ref["loglike"] - 1e-7
        # against the exact optimum of the selected model: the north_star tolerances
        A, w, C, N = loglik.rows_from_bins(L, case["bins_list"])
        groups = {int(k): v for k, v in case["repr_to_merged_variants"]}
        reps = sorted({int(dict(case["be_unidentifiable_to"])[cid]) for cid in prop})
        mask = np.zeros(len(L), bool)
        mask[reps] = True
        r = exact.exact_fit(A, L, w, mask=mask)
        for rep in reps:
            for member in groups[rep]:
                assert abs(prop[member] - r["theta"][rep] / len(groups[rep])) < 1e-6
        assert abs(like - (r["loglike_no_C"] + C)) < 1e-9 * abs(like)
        # the criterion is the one of the last accepted FIT (zero-proportion drops do not refit, :333-347)
        assert abs(crit - ref["criterion_value"]) < 1e-8 * abs(crit)
        fitter.close()


def test_point_estimate_and_failure_mode(gpu, synthetic_small):
    for case in synthetic_small["cases"]:
        model = model_from_tables(case["variant_sizes"], case["bins_list"])
        kw = fit_context(model, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        fitter = ModelFitMaxLike(**kw)
        for fit in case["fits"]:
            if fit["kind"] != "point_estimate":
                continue
            chosen = set(fit["chosen_ids"]) if fit["chosen_ids"] else None
            if "error" in fit:
                with pytest.raises(Exception, match="Likelihood maximization failed."):
                    fitter.point_estimate(chosen_ids=chosen, criterion=fit["criterion"])
                continue
            prop, like, crit = fitter.point_estimate(chosen_ids=chosen, criterion=fit["criterion"])
            assert list(prop) == [cid for cid, _ in fit["prop"]]
            for cid, p in fit["prop"]:
                assert abs(prop[cid] - p) < 5e-4
            assert abs(like - fit["loglike"]) < 1e-8 * abs(like) and like >= fit["loglike"] - 1e-7
            assert abs(crit - fit["criterion_value"]) < 1e-8 * abs(crit)
        fitter.close()


def test_neg_loglike_callable_matches_known_answers(gpu, plastome, synthetic_small):
    """get_neg_likelihood_of_var_freq().loglike_func == the reference's lambdified objective (known answers
    from get_like_formula(floats, math.log), ModelGenerator.py:110-178)."""
    cases = [(plastome["variant_sizes"], plastome["models"][0]["bins_list"], plastome["known_answer_ll"],
              plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])]
    cases += [(c["variant_sizes"], c["bins_list"], c["known_answer_ll"], c["be_unidentifiable_to"], c["repr_to_merged_variants"])
              for c in synthetic_small["cases"]]
    for L, bins, kas, unid, groups in cases:
        model = model_from_tables(L, bins)
        fitter = ModelFitMaxLike(**fit_context(model, unid, groups))
        for ka in kas:
            ids = sorted(ka["subset"]) if ka["subset"] else list(range(len(L)))
            info = fitter.get_neg_likelihood_of_var_freq(within_variant_ids=set(ids))
            theta = np.array([hexf(t) for t in ka["theta"]])
            got = -float(info.loglike_func(theta[ids])[0])
            ref = hexf(ka["loglike"])
            assert info.variable_size == ka["variable_size"] and info.sample_size == ka["sample_size"]
            if math.isinf(ref):
                assert math.isinf(got) and got < 0
            else:
                assert abs(got - ref) < 1e-9 * abs(ref)
        fitter.close()


def test_cfg1_bootstrap_flow_on_gpu(gpu, plastome):
    """BASELINE.json configs[0]: whole-data selection + 100 bootstrap replicates of the reference run, all
    replicates in lock step on the GPU: same chosen model in every replicate, same support table,
    %.4f proportions of bootstraps.replicates.tab reproduced (up to the reference's own SLSQP noise)."""
    L = plastome["variant_sizes"]
    models = [model_from_tables(L, m["bins_list"]) for m in plastome["models"]]
    kw0 = fit_context(models[0], plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
    full = ModelFitMaxLike(**kw0)
    raw = full.reverse_model_selection(n_proc=1)
    assert abs(raw[0][0] - 23 / 35.) < 1e-7 and abs(raw[0][1] - 12 / 35.) < 1e-7
    assert abs(raw[1] - plastome["final"]["res_loglike"]) < 1e-9 * abs(raw[1])
    assert abs(raw[2] - plastome["final"]["res_criterion"]) < 1e-9 * abs(raw[2])
    shared = {k: kw0[k] for k in ("variant_paths", "variant_subpath_counters", "repr_to_merged_variants", "be_unidentifiable_to")}
    reps = []
    for m in models[1:]:
        kw = fit_context(m, plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
        reps.append((m, kw["sbp_to_sbp_id"]))
    s = bootstrap.do_subsampling(reps, shared, full, raw, n_replicate=100, threshold=0.95)
    assert s.bs_eligible and s.n_replicates_run == 100
    got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
    assert got == plastome["final"]["support"]
    ctx = full._ensure_lane()
    assert ctx.n_batches <= 6
    tab = plastome["output_tables"]["bootstraps.replicates.tab"].strip().split("\n")[1:]
    for line, v_prop, fit, m in zip(tab, s.variant_proportions_reps, plastome["fits"][1:], plastome["models"][1:]):
        ref4 = [float(x) for x in line.split("\t")[1:]]
        assert all(abs(round(v_prop[c], 4) - ref4[c]) <= 1.01e-4 for c in (0, 1))
        A, w, C, N = loglik.rows_from_bins(L, m["bins_list"])
        r = exact.exact_fit(A, L, w)
        assert abs(v_prop[0] - r["theta"][0]) < 1e-6 and abs(v_prop[1] - r["theta"][1]) < 1e-6
    # the result tables of this run against the files the reference wrote: same layout, fids and
    # support; %.4f cells within the reference's SLSQP noise, log-likelihood and BIC within 1e-9 relative
    from traversome_b200 import output
    cid_to_fid = output.cid_to_fid_table(s.vp_unique_results_sorted, raw[0])
    mine = output.final_result_tab(s, raw[0], raw[1], raw[2], cid_to_fid).strip().split("\n")
    ref = plastome["output_tables"]["final.result.tab"].strip().split("\n")
    assert mine[0] == ref[0] and len(mine) == len(ref) == 2
    a, b = mine[1].split("\t"), ref[1].split("\t")
    assert a[:2] == b[:2]                                                        # solution id, support
    assert all(abs(float(x) - float(y)) <= 1e-9 * abs(float(y)) for x, y in zip(a[2:4], b[2:4]))
    assert all(abs(float(x) - float(y)) <= 1.01e-4 for x, y in zip(a[4:6], b[4:6]))
    assert all(abs(float(x) - float(y)) <= 2e-4 for x, y in zip(a[6].split(","), b[6].split(",")))
    rep_mine = output.replicates_tab(s.variant_proportions_reps, cid_to_fid).split("\n")
    rep_ref = plastome["output_tables"]["bootstraps.replicates.tab"].split("\n")
    assert rep_mine[0] == rep_ref[0] and len(rep_mine) == len(rep_ref)
    assert [ln.split("\t")[0] for ln in rep_mine] == [ln.split("\t")[0] for ln in rep_ref]
    full.close()


def test_indexed_lanes_equal_expanded_lanes(gpu):
    from traversome_b200 import synthetic
    wl = synthetic.make_workload(10, 500, 0.3, seed=4)
    cols = synthetic.bootstrap_weights(wl.n_obs, 3, seed=1)
    idx = np.array([0, 2, 2, 1, 0, 1, 2], np.int32)
    mask = np.ones((wl.V, idx.size), np.uint8)
    mask[3, 2] = 0
    mask[5, 4] = 0
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes_indexed(cols, idx, mask=mask)
        assert np.array_equal(eng.get_weights(), cols[:, idx])
        a = eng.fit()
        eng.set_lanes(cols[:, idx], mask=mask)
        b = eng.fit()
    assert np.array_equal(a["theta"], b["theta"], equal_nan=True) and np.array_equal(a["loglike"], b["loglike"])
    assert np.array_equal(a["status"], b["status"]) and (a["status"] == _cabi.CONVERGED).sum() >= 5
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:       # 16-bit upload == 32-bit upload
        eng.set_lanes(cols[:, idx].astype(np.uint16), mask=mask)
        assert np.array_equal(eng.get_weights(), cols[:, idx])
        c = eng.fit()
    assert np.array_equal(a["theta"], c["theta"], equal_nan=True) and np.array_equal(a["loglike"], c["loglike"])
    with pytest.raises(_cabi.TvsError):
        with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
            eng.set_lanes_indexed(cols, np.array([0, 3], np.int32))


def test_lane_context_batches_mixed_replicates_and_subsets(gpu):
    from traversome_b200 import synthetic
    wl = synthetic.make_workload(12, 900, 0.3, seed=9)
    ctx = LaneContext(wl.V, wl.L)
    ctx.matrix._sigs = [tuple((int(c), int(v)) for c, v in zip(wl.col[a:b], wl.val[a:b]))
                        for a, b in zip(wl.row_ptr[:-1], wl.row_ptr[1:])]
    ctx.matrix._row_of = {s: i for i, s in enumerate(ctx.matrix._sigs)}
    W = synthetic.bootstrap_weights(wl.n_obs, 4, seed=2)
    cols = [ctx.add_weights(W[:, j]) for j in range(4)]
    reqs = [LaneRequest(cols[j], 10.0 * j, set(range(wl.V)) - ({k} if k >= 0 else set())) for j in range(4) for k in (-1, 0, 5)]
    res = ctx.fit(reqs)
    A = wl.scipy_csr()
    n_ok = 0
    for rq, fr in zip(reqs, res):
        m = np.zeros(wl.V, bool)
        m[rq.ids] = True
        j = cols.index(rq.col)
        try:
            r = exact.exact_fit(A, wl.L, W[:, j], mask=m)
        except ValueError:                      # the subset leaves an observed read path uncovered
            assert not fr.success and fr.status == _cabi.INFEASIBLE
            continue
        n_ok += 1
        assert fr.success and np.abs(fr.x - r["theta"][rq.ids]).max() < 1e-6
        assert abs(-fr.fun - (r["loglike_no_C"] + rq.C)) < 1e-9 * abs(fr.fun)
    assert n_ok >= 4
    ctx.close()


@pytest.mark.parametrize("criterion", ["BIC", "AIC"])
def test_backward_selection_decisions_equal_exact_replay(gpu, criterion):
    """cfg3-shaped case at test size: 24 candidates of which 8 carry reads.  The backward selection driven by the CUDA
    solver takes exactly the decisions of the same driver fed with exact CPU optima (oracle), with a pinned shuffle:
    same dropped variants in the same order, same final model, criterion within 1e-9 relative."""
    wl = synthetic.make_workload(24, 1200, 0.5, seed=77, n_true=8)

    def run(engine_factory):
        model = model_from_workload(wl)
        kw = fit_context(model)
        ctx = LaneContext(wl.V, wl.L)
        if engine_factory is not None:
            ctx.engine_factory = engine_factory
        fitter = ModelFitMaxLike(context=ctx, **kw)
        rng = np.random.default_rng(5)
        fitter._shuffle = lambda x: rng.shuffle(x)
        out = fitter.reverse_model_selection(n_proc=1, criterion=criterion)
        n = ctx.n_lanes_fitted
        ctx.close()
        return out, n

    (gp, gl, gc), n_gpu = run(None)
    (cp, cl, cc), n_cpu = run(OracleEngine)
    assert sorted(gp) == sorted(cp)
    assert n_gpu == n_cpu                                   # same number of candidate fits: same decision path
    assert all(abs(gp[k] - cp[k]) < 1e-6 for k in gp)
    assert abs(gl - cl) < 1e-9 * abs(cl) and abs(gc - cc) < 1e-9 * abs(cc)
    assert n_gpu >= 5 and len(gp) < wl.V                    # the selection did drop variants and tested candidates


@pytest.mark.parametrize("device_bins", [False, True])
def test_cfg1_bootstrap_from_record_pool_on_gpu(gpu, plastome, device_bins):
    """cfg1 without any per-replicate object graph: records -> replayed draws -> bit-exact bins -> weight columns ->
    lock-step backward selection on the CUDA solver.  Same support table as the reference run; every replicate's
    proportions within the reference's own SLSQP spread and within 1e-6 of the exact optimum of ITS replicate.
    device_bins: the bin statistics of all 100 replicates in one device call (include/tvs_pool.h)."""
    import random
    from traversome_b200.resample import RecordPool
    L = plastome["variant_sizes"]
    pool = RecordPool.from_tables(plastome)
    pattern = np.zeros((pool.n_rows, len(L)), dtype=bool)
    ctx = LaneContext(len(L), L)
    for r, rp in enumerate(plastome["read_paths"]):
        fv = {int(v): int(c) for v, c in rp["from_variants"]}
        assert ctx.matrix.row_index_distinct(fv) == r
        pattern[r, list(fv)] = True
    groups = {int(k): [int(x) for x in v] for k, v in plastome["repr_to_merged_variants"]}
    unid = {int(k): int(v) for k, v in plastome["be_unidentifiable_to"]}
    w0, C0, N0, obs0 = pool.lane_of(np.arange(pool.n_records))
    full = ModelFitMaxLike.from_lane(ctx, (ctx.add_weights(w0), C0, N0), pattern, np.nonzero(obs0)[0], groups, unid)
    raw = full.reverse_model_selection(n_proc=1)
    assert abs(raw[0][0] - 23 / 35.) < 1e-7 and abs(raw[1] - plastome["final"]["res_loglike"]) < 1e-9 * abs(raw[1])
    s = bootstrap.do_subsampling_from_records(pool, pattern, ctx, full, raw, random.Random(12345), n_replicate=100, device_bins=device_bins)
    assert s.bs_eligible and s.n_replicates_run == 100
    got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
    assert got == plastome["final"]["support"]
    A = ctx.matrix.dense()
    for k, (v_prop, fit) in enumerate(zip(s.variant_proportions_reps, plastome["fits"][1:])):
        for cid, p in fit["prop"]:
            assert abs(v_prop[cid] - p) < 5e-4
        if k < 10:
            w = ctx.dense_columns([k + 1])[:, 0]
            r = exact.exact_fit(A, L, w)
            assert abs(v_prop[0] - r["theta"][0]) < 1e-6 and abs(v_prop[1] - r["theta"][1]) < 1e-6
    assert ctx.n_batches <= 6
    ctx.close()


def test_lane_context_concurrent_halves_on_gpu(gpu):
    """A 3,200-lane batch through LaneContext: two concurrent handles give the same fits as one (each lane is an
    independent problem; only the batch a lane shares its iterations with differs)."""
    from traversome_b200 import synthetic
    from traversome_b200.lanes import LaneContext, LaneRequest
    wl = synthetic.make_workload(24, 3000, 0.25, seed=61)
    W = synthetic.bootstrap_weights(wl.n_obs, 3200, seed=2)
    res = []
    for limit in (0, 3000):
        ctx = LaneContext(wl.V, wl.L)
        ctx.concurrent_min_lanes = limit
        for r in range(wl.R):
            lo, hi = int(wl.row_ptr[r]), int(wl.row_ptr[r + 1])
            ctx.matrix.row_index_distinct(dict(zip(wl.col[lo:hi].tolist(), wl.val[lo:hi].tolist())))
        reqs = [LaneRequest(ctx.add_weights(W[:, b]), 0.0, range(wl.V)) for b in range(W.shape[1])]
        res.append(ctx.fit(reqs))
        assert (ctx._engine2 is not None) == (limit > 0)
        ctx.close()
    assert all(r.success for r in res[0]) and all(r.success for r in res[1])
    assert max(np.abs(a.x - b.x).max() for a, b in zip(*res)) < 1e-7
    assert max(abs(a.fun - b.fun) / abs(a.fun) for a, b in zip(*res)) < 1e-12


def test_batches_in_flight_equal_batches_one_at_a_time(gpu):
    """LaneContext.fit_batches / fit_weight_batches: several batches in flight on separate handles (own CUDA stream and
    host thread each) give, batch for batch, bitwise the fits of the same batch alone; masks, constants, warm starts,
    odd and tiny batches included."""
    from traversome_b200 import synthetic
    from traversome_b200.lanes import LaneBatch, LaneContext
    wl = synthetic.make_workload(40, 2000, 0.2, seed=62)
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    rng = np.random.default_rng(9)
    sizes = [301, 64, 7, 500, 1, 130, 222, 33, 96]
    batches, blocks = [], []
    for j, n in enumerate(sizes):
        cols = ctx.add_resampled(wl.n_obs, n, seed=70 + j)
        mask = None
        theta0 = None
        if j % 3 == 1:                                          # subsets that keep every read path covered: drop one rare variant
            mask = np.ones((wl.V, n), dtype=np.uint8)
            theta0 = rng.random((wl.V, n))
            theta0 /= theta0.sum(axis=0)
        batches.append(LaneBatch(np.asarray(cols), C=rng.normal(size=n) if j % 2 else None, mask=mask, theta0=theta0))
    alone = [ctx.fit_batch(b) for b in batches]
    piped = ctx.fit_batches(batches)
    assert ctx._engine2 is not None and len(ctx._engines_more) == 2            # four handles
    for a, b in zip(alone, piped):
        assert bool(ctx.lane_ok(a).all())
        for k in ("theta", "loglike", "kkt", "iters", "status"):
            assert np.array_equal(a[k], b[k]), k
    eng = ctx.engine()
    for b in batches[:5]:
        eng.set_lanes_columns(ctx._slots(eng, b.cols))
        blocks.append(eng.get_weights().astype(np.uint16))
    got = ctx.fit_weight_batches(blocks, C=[b.C for b in batches[:5]])
    for a, b, src in zip(alone, got, batches[:5]):
        if src.theta0 is None:                                  # the same lanes from the same start: the same iterates
            assert np.array_equal(a["theta"], b["theta"]) and np.array_equal(a["loglike"], b["loglike"])
        else:
            assert np.abs(a["theta"] - b["theta"]).max() < 1e-6 and np.all(np.abs(a["loglike"] - b["loglike"]) < 1e-9 * np.abs(a["loglike"]))
    ctx.pipeline_depth = 2
    again = ctx.fit_batches(batches[:3])
    assert all(np.array_equal(a["theta"], b["theta"]) for a, b in zip(alone, again))
    ctx.close()


def test_weight_forms_and_the_column_store(gpu):
    """The lane weights live on the device as 16-bit counts when every count fits and as int32 otherwise; the column
    store (tvs_columns_*) registers a replicate's weight vector once and batches name it by slot.  All forms give the
    same fits; counts above 65,535 take the int32 path; a negative weight is refused."""
    from traversome_b200 import synthetic
    wl = synthetic.make_workload(12, 800, 0.3, seed=31)
    B = 10
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=6)
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes(w)                                        # int32 upload, 16-bit copy made on the device
        a = eng.fit()
        eng.set_lanes(w.astype(np.uint16))                      # 16-bit upload, stays 16-bit
        b = eng.fit()
        assert np.array_equal(eng.get_weights(), w)             # widened on request
        assert np.array_equal(a["theta"], b["theta"]) and np.array_equal(a["loglike"], b["loglike"])
        # the column store: int32 and uint16 uploads, lanes by slot in any order and with repeats
        eng.columns_reserve(4)
        eng.columns_upload(0, w[:, :4])
        eng.columns_reserve(B + 3)                              # growing keeps the registered columns
        eng.columns_upload(4, w[:, 4:].astype(np.uint16))
        order = np.array([3, 0, 9, 9, 5, 1, 2, 4, 6, 7, 8], dtype=np.int32)
        eng.set_lanes_columns(order)
        c = eng.fit()
        assert np.array_equal(eng.get_weights(), w[:, order])
        assert np.abs(c["theta"] - a["theta"][:, order]).max() < 1e-7       # same lanes in another batch geometry
        assert np.array_equal(c["theta"][:, 2], c["theta"][:, 3])           # the repeated column: the same lane twice
        # device-drawn columns depend on (seed, id) only, not on the slot or the call that drew them
        eng.columns_resample(B, 3, wl.n_obs, seed=77, first_id=5)
        eng.set_lanes_columns(np.array([B, B + 1, B + 2], dtype=np.int32))
        d1 = eng.get_weights()
        eng.columns_resample(0, 2, wl.n_obs, seed=77, first_id=6)
        eng.set_lanes_columns(np.array([0, 1], dtype=np.int32))
        d2 = eng.get_weights()
        assert np.array_equal(d1[:, 1:3], d2) and np.all(d1.sum(0) == wl.N)
        with pytest.raises(_cabi.TvsError):
            eng.set_lanes_columns(np.array([B + 3], dtype=np.int32))        # slot outside the store
        # counts above 65,535: int32 path, same optimum (theta is invariant under a scaling of the weights)
        big = w.astype(np.int64) * 40000
        assert big.max() > 65535 and big.max() < 2 ** 31
        eng.set_lanes(big.astype(np.int32))
        e = eng.fit()
        assert np.all(e["status"] == _cabi.CONVERGED) and np.abs(e["theta"] - a["theta"]).max() < 1e-7
        assert np.array_equal(eng.get_weights(), big.astype(np.int32))
        neg = w.copy()
        neg[5, 2] = -1
        with pytest.raises(_cabi.TvsError):
            eng.set_lanes(neg)
    check_fit_against_exact_lane(wl, w, a, [0, 3, B - 1])


def check_fit_against_exact_lane(wl, w, out, lanes):
    A = wl.scipy_csr()
    for b in lanes:
        r = exact.exact_fit(A, wl.L, w[:, b])
        assert np.abs(out["theta"][:, b] - r["theta"]).max() < 1e-6
        assert abs(out["loglike"][b] - r["loglike_no_C"]) < 1e-9 * abs(r["loglike_no_C"])
