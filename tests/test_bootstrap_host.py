"""CPU: host logic of the batched bootstrap / model-selection drivers (decision replay, tally,
lock-step batching), with the GPU engine replaced by the oracle-backed stand-in of tests/helpers.py."""
import numpy as np
import pytest

from tests.helpers import OracleEngine, fit_context, model_from_tables
from traversome_b200 import bootstrap
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike
from traversome_b200.lanes import LaneContext, LaneSharder


def _ctx(L):
    ctx = LaneContext(len(L), L)
    ctx.engine_factory = OracleEngine
    return ctx


def _golden_selection(case, criterion):
    for fit in case["fits"]:
        if fit["kind"] == "reverse_model_selection" and fit["criterion"] == criterion:
            return fit
    raise KeyError(criterion)


@pytest.mark.parametrize("criterion", ["BIC", "AIC"])
def test_reverse_model_selection_decisions_match_reference(synthetic_small, criterion):
    """Same chosen variant set as the reference's own run (golden), proportions within the
    reference's SLSQP spread, criterion value within 1e-9 relative."""
    for case in synthetic_small["cases"]:
        model = model_from_tables(case["variant_sizes"], case["bins_list"])
        kw = fit_context(model, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        fitter = ModelFitMaxLike(context=_ctx(case["variant_sizes"]), **kw)
        prop, like, crit = fitter.reverse_model_selection(n_proc=1, criterion=criterion)
        ref = _golden_selection(case, criterion)
        assert sorted(prop) == [cid for cid, _ in ref["prop"]], case["name"]
        for cid, p in ref["prop"]:
            assert abs(prop[cid] - p) < 5e-4
        assert abs(like - ref["loglike"]) < 1e-8 * abs(ref["loglike"])
        assert abs(crit - ref["criterion_value"]) < 1e-8 * abs(ref["criterion_value"])
        assert like >= ref["loglike"] - 1e-7            # the exact optimum is never worse than SLSQP's


def test_point_estimate_matches_reference_and_failure_mode(synthetic_small):
    for case in synthetic_small["cases"]:
        model = model_from_tables(case["variant_sizes"], case["bins_list"])
        kw = fit_context(model, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        fitter = ModelFitMaxLike(context=_ctx(case["variant_sizes"]), **kw)
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
            assert abs(crit - fit["criterion_value"]) < 1e-8 * abs(fit["criterion_value"])


def test_lock_step_selection_equals_serial(synthetic_small):
    """All replicates' coroutines advanced together give exactly the serial results, in fewer engine calls."""
    case = synthetic_small["cases"][2]
    L = case["variant_sizes"]
    rng = np.random.default_rng(0)
    models = []
    for k in range(5):                      # 5 "replicates": resampled num_matched of the same bins
        bins = [dict(b, rp_bins=[dict(rp, num_matched=int(rng.poisson(rp["num_matched"])) + 1) for rp in b["rp_bins"]])
                for b in case["bins_list"]]
        models.append(model_from_tables(L, bins))
    serial = []
    for m in models:
        kw = fit_context(m, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        f = ModelFitMaxLike(context=_ctx(L), **kw)
        f._shuffle = lambda x: None
        serial.append(f.reverse_model_selection(n_proc=1))
    ctx = _ctx(L)
    fitters = []
    for m in models:
        kw = fit_context(m, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        f = ModelFitMaxLike(context=ctx, **kw)
        f._shuffle = lambda x: None
        fitters.append(f)
    batched = bootstrap.run_selections(fitters)
    for (p0, l0, c0), (p1, l1, c1) in zip(serial, batched):
        assert list(p0) == list(p1)
        assert all(abs(p0[k] - p1[k]) < 1e-9 for k in p0) and abs(l0 - l1) < 1e-9 * abs(l0) and abs(c0 - c1) < 1e-9 * abs(c0)
    assert ctx.n_batches < sum(1 for _ in models) * 3      # far fewer engine calls than serial fits
    assert ctx.n_lanes_fitted == sum(f.n_gpu_fits for f in fitters)


@pytest.mark.parametrize("groups", [2, 3])
def test_selection_in_groups_equals_one_group_and_is_reproducible(synthetic_small, groups):
    """Replicates advanced as interleaved groups on host threads (their batches take the engine strictly in turn) select
    the same models as one lock-step group; with the global generator seeded, two runs draw the same permutations."""
    case = synthetic_small["cases"][2]
    L = case["variant_sizes"]
    rng = np.random.default_rng(1)
    models = []
    for k in range(7):
        bins = [dict(b, rp_bins=[dict(rp, num_matched=int(rng.poisson(rp["num_matched"])) + 1) for rp in b["rp_bins"]])
                for b in case["bins_list"]]
        models.append(model_from_tables(L, bins))

    def run(n_groups, pin):
        ctx = _ctx(L)
        fitters = []
        for m in models:
            f = ModelFitMaxLike(context=ctx, **fit_context(m, case["be_unidentifiable_to"], case["repr_to_merged_variants"]))
            if pin:
                f._shuffle = lambda x: None
            fitters.append(f)
        np.random.seed(11)
        out = bootstrap.run_selections(fitters, groups=n_groups, random_size=2)
        assert ctx.n_lanes_fitted == sum(f.n_gpu_fits for f in fitters)
        assert all(f._shuffle is np.random.shuffle for f in fitters) or pin
        return out, ctx.n_batches

    one, nb1 = run(1, True)
    many, nbg = run(groups, True)
    assert nbg > nb1                                                        # each group has its own rounds
    for (p0, l0, c0), (p1, l1, c1) in zip(one, many):
        assert list(p0) == list(p1)
        assert all(abs(p0[k] - p1[k]) < 1e-9 for k in p0) and abs(l0 - l1) < 1e-9 * abs(l0) and abs(c0 - c1) < 1e-9 * abs(c0)
    a, _ = run(groups, False)
    b, _ = run(groups, False)
    for (p0, l0, c0), (p1, l1, c1) in zip(a, b):
        assert list(p0) == list(p1) and l0 == l1 and c0 == c1


def test_a_failing_group_does_not_hang_the_others():
    from traversome_b200.bootstrap import _Turns
    import threading
    turns = _Turns(3)
    order = []

    def member(g, n):
        try:
            for k in range(n):
                with turns.gate(g):
                    order.append(g)
                    if g == 1 and k == 1:
                        raise RuntimeError("boom")
        except RuntimeError:
            pass
        finally:
            turns.finish(g)
    ts = [threading.Thread(target=member, args=(g, n)) for g, n in ((0, 4), (1, 5), (2, 2))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=20)
    assert not any(t.is_alive() for t in ts)
    assert order == [0, 1, 2, 0, 1, 2, 0, 0]


def test_support_tally_matches_reference_tables(plastome):
    """summarize_replicates / check_bs_threshold on the reference's own replicate results reproduce its
    support table (final.result.tab) -- traversome.py:561-617."""
    reps = [dict((int(c), p) for c, p in fit["prop"]) for fit in plastome["fits"][1:]]
    raw = dict((int(c), p) for c, p in plastome["final"]["variant_proportions"])
    count_unique = {}
    for v in reps:
        assert bootstrap.check_bs_threshold(count_unique, v, len(reps), 0.95)
    s = bootstrap.summarize_replicates(reps, raw, plastome["final"]["res_loglike"], plastome["final"]["res_criterion"])
    got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
    assert got == plastome["final"]["support"]
    top = s.vp_unique_results[s.vp_unique_results_sorted[0]]
    assert top["raw_like"] == plastome["final"]["res_loglike"] and top["raw_criterion"] == plastome["final"]["res_criterion"]
    std = np.std(np.array(top["values"]), axis=0, ddof=1)      # traversome.py output_result writes the sample std
    line = plastome["output_tables"]["final.result.tab"].strip().split("\n")[1].split("\t")
    assert [float(x) for x in line[-1].split(",")] == [round(float(x), 4) for x in std] or \
        all(abs(float(a) - b) < 1.01e-4 for a, b in zip(line[-1].split(","), std))


def test_bs_threshold_early_stop_index():
    count = {}
    seq = [{0: 1.0}, {1: 1.0}, {0: .5, 1: .5}, {0: 1.0}]
    flags = [bootstrap.check_bs_threshold(count, v, 4, 0.95) for v in seq]
    assert flags == [True, False, False, False]
    count = {}
    assert all(bootstrap.check_bs_threshold(count, {0: 1.0}, 10, 0.95) for _ in range(10))


def test_do_subsampling_on_plastome_replicates(plastome):
    """cfg1 end to end above the C ABI (engine = oracle stand-in): 100 replicate models of the reference run ->
    identical support table; every replicate's proportions within the reference's own tolerance."""
    L = plastome["variant_sizes"]
    models = [model_from_tables(L, m["bins_list"]) for m in plastome["models"]]
    ctx = _ctx(L)
    kw0 = fit_context(models[0], plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
    full = ModelFitMaxLike(context=ctx, **kw0)
    raw = full.reverse_model_selection(n_proc=1)
    shared = {k: kw0[k] for k in ("variant_paths", "variant_subpath_counters", "repr_to_merged_variants", "be_unidentifiable_to")}
    reps = []
    for m in models[1:]:
        kw = fit_context(m, plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
        reps.append((m, kw["sbp_to_sbp_id"]))
    s = bootstrap.do_subsampling(reps, shared, full, raw, n_replicate=100, threshold=0.95)
    assert s.bs_eligible and s.n_replicates_run == 100
    got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
    assert got == plastome["final"]["support"]
    for v_prop, fit in zip(s.variant_proportions_reps, plastome["fits"][1:]):
        for cid, p in fit["prop"]:
            assert abs(v_prop[cid] - p) < 5e-4
    assert ctx.n_batches <= 6            # 1 (full) + a handful of lock-step rounds for 100 replicates


def test_shard_bounds_cover_all_lanes():
    for B in (1, 2, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [LaneSharder.bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_vectorised_cover_rule_equals_reference_set_logic(synthetic_small, plastome):
    """ModelFitMaxLike.cover_all_observed_subpaths (numpy tables) == the reference's set unions (ModelFitMaxLike.py:568-580)."""
    rng = np.random.default_rng(2)
    cases = [(c["variant_sizes"], c["bins_list"], c["be_unidentifiable_to"], c["repr_to_merged_variants"]) for c in synthetic_small["cases"]]
    cases.append((plastome["variant_sizes"], plastome["models"][3]["bins_list"], plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"]))
    for L, bins, unid, groups in cases:
        model = model_from_tables(L, bins)
        kw = fit_context(model, unid, groups)
        # make some read paths unobserved and drop some from the id table, like a bootstrap replicate does
        for k, spi in enumerate(model.all_sub_paths.values()):
            if k % 7 == 3:
                spi.mapped_records = []
        fitter = ModelFitMaxLike(context=_ctx(L), **kw)
        fitter.update_observed_sp_ids()
        V = len(L)
        for _ in range(60):
            ids = set(int(v) for v in rng.choice(V, rng.integers(0, V + 1), replace=False))
            model_sp = set()
            for v in ids:
                for sp_ in kw["variant_subpath_counters"][kw["variant_paths"][v]]:
                    if sp_ in kw["sbp_to_sbp_id"]:
                        model_sp.add(kw["sbp_to_sbp_id"][sp_])
            assert fitter.cover_all_observed_subpaths(ids) == fitter.observed_sbp_id_set.issubset(model_sp)

        # with the cover counts of a base set cached (what a selection round does), leave-one-out subsets, the base set
        # itself and unrelated sets all answer like the set logic
        def by_sets(ids):
            model_sp = set()
            for v in ids:
                for sp_ in kw["variant_subpath_counters"][kw["variant_paths"][v]]:
                    if sp_ in kw["sbp_to_sbp_id"]:
                        model_sp.add(kw["sbp_to_sbp_id"][sp_])
            return fitter.observed_sbp_id_set.issubset(model_sp)

        for _ in range(12):
            base = set(int(v) for v in rng.choice(V, rng.integers(1, V + 1), replace=False))
            fitter._set_cover_base(base)
            assert fitter.cover_all_observed_subpaths(base) == by_sets(base)
            for v in base:
                assert fitter.cover_all_observed_subpaths(base - {v}) == by_sets(base - {v})
            other = set(int(v) for v in rng.choice(V, rng.integers(1, V + 1), replace=False))
            assert fitter.cover_all_observed_subpaths(other) == by_sets(other)


def test_do_subsampling_from_records_reproduces_the_reference_run(plastome):
    """cfg1 from the record pool in array form: replayed draws (random.Random(12345)), bit-exact bins, lock-step
    selection -> the reference's support table and, per replicate, its proportions (engine = oracle stand-in)."""
    import random
    from traversome_b200.resample import RecordPool
    L = plastome["variant_sizes"]
    model0 = model_from_tables(L, plastome["models"][0]["bins_list"])
    pool = RecordPool.from_tables(plastome)
    pattern = np.zeros((pool.n_rows, len(L)), dtype=bool)
    ctx = _ctx(L)
    for r, rp in enumerate(plastome["read_paths"]):
        fv = {int(v): int(c) for v, c in rp["from_variants"]}
        assert ctx.matrix.row_index_distinct(fv) == r
        pattern[r, list(fv)] = True
    kw0 = fit_context(model0, plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
    w0, C0, N0, obs0 = pool.lane_of(np.arange(pool.n_records))
    full = ModelFitMaxLike.from_lane(ctx, (ctx.add_weights(w0), C0, N0), pattern, np.nonzero(obs0)[0],
                                     kw0["repr_to_merged_variants"], kw0["be_unidentifiable_to"])
    raw = full.reverse_model_selection(n_proc=1)
    assert abs(raw[0][0] - 23 / 35.) < 1e-7
    assert abs(raw[1] - plastome["final"]["res_loglike"]) < 1e-9 * abs(raw[1])
    s = bootstrap.do_subsampling_from_records(pool, pattern, ctx, full, raw, random.Random(12345), n_replicate=100)
    assert s.bs_eligible and s.n_replicates_run == 100
    got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
    assert got == plastome["final"]["support"]
    for v_prop, fit in zip(s.variant_proportions_reps, plastome["fits"][1:]):
        for cid, p in fit["prop"]:
            assert abs(v_prop[cid] - p) < 5e-4


def test_jackknife_from_records_equals_serial_fits(plastome):
    """Jackknife replicates (random.sample replay), AIC, a user-fixed variant: the lock-step batch gives, replicate by
    replicate, what one fitter at a time gives on the same lanes."""
    import random
    from traversome_b200.resample import RecordPool
    from traversome_b200.lanes import LaneRequest
    L = plastome["variant_sizes"]
    pool = RecordPool.from_tables(plastome)
    pattern = np.zeros((pool.n_rows, len(L)), dtype=bool)
    ctx = _ctx(L)
    for r, rp in enumerate(plastome["read_paths"]):
        fv = {int(v): int(c) for v, c in rp["from_variants"]}
        ctx.matrix.row_index_distinct(fv)
        pattern[r, list(fv)] = True
    groups = {int(k): [int(x) for x in v] for k, v in plastome["repr_to_merged_variants"]}
    unid = {int(k): int(v) for k, v in plastome["be_unidentifiable_to"]}
    w0, C0, N0, obs0 = pool.lane_of(np.arange(pool.n_records))
    full = ModelFitMaxLike.from_lane(ctx, (ctx.add_weights(w0), C0, N0), pattern, np.nonzero(obs0)[0], groups, unid)
    raw = full.reverse_model_selection(n_proc=1, criterion="AIC", user_fixed_ids=[1])
    s = bootstrap.do_subsampling_from_records(pool, pattern, ctx, full, raw, random.Random(99), n_replicate=10, criterion="AIC",
                                              jackknife=10, user_fixed_ids=[1], threshold=0.5)
    assert s.n_replicates_run == 10
    rng = random.Random(99)
    for k in range(10):
        sample = pool.sample_jackknife(rng, int(pool.n_records / 10.0))
        assert len(sample) == pool.n_records - pool.n_records // 10
        w, C, N, observed = pool.lane_of(sample)
        ctx1 = _ctx(L)
        for rp in plastome["read_paths"]:
            ctx1.matrix.row_index_distinct({int(v): int(c) for v, c in rp["from_variants"]})
        f = ModelFitMaxLike.from_lane(ctx1, (ctx1.add_weights(w), C, N), pattern, np.nonzero(observed)[0], groups, unid)
        prop, like, crit = f.reverse_model_selection(n_proc=1, criterion="AIC", user_fixed_ids=[1])
        got = s.variant_proportions_reps[k]
        assert list(got) == list(prop) and all(abs(got[c] - prop[c]) < 1e-12 for c in prop)
        assert 1 in prop


def test_large_batches_run_as_two_concurrent_halves(synthetic_small):
    """LaneContext.fit cuts a batch of >= concurrent_min_lanes lanes in two halves on two handles (two threads); the
    results come back in request order and equal those of the single-handle path."""
    from traversome_b200.lanes import LaneRequest
    case = synthetic_small["cases"][0]
    L = case["variant_sizes"]
    model = model_from_tables(L, case["bins_list"])
    V = len(L)
    rng = np.random.default_rng(5)
    outs = []
    for limit in (0, 4):
        ctx = _ctx(L)
        ctx.concurrent_min_lanes = limit
        col, C, _N = ctx.add_model(model)
        reqs = []
        rng = np.random.default_rng(5)
        for b in range(9):
            ids = sorted(set(rng.choice(V, rng.integers(max(2, V - 2), V + 1), replace=False).tolist()))
            reqs.append(LaneRequest(col, C, ids))
        res = ctx.fit(reqs)
        outs.append(res)
        assert (ctx._engine2 is not None) == (limit == 4)
        if limit == 4:
            assert ctx._engine.calls == 1 and ctx._engine2.calls == 1
        ctx.close()
    for a, b in zip(*outs):
        assert a.success == b.success and a.status == b.status
        if a.success:
            assert np.allclose(a.x, b.x, atol=1e-12) and abs(a.fun - b.fun) <= 1e-12 * abs(a.fun)
