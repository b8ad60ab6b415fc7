"""CPU: host-side logic of the boundary -- record types, get_like_formula mirror, integer CSR build."""
import math

import numpy as np

from oracle import loglik
from tests.helpers import fit_context, hexf, model_from_tables
from traversome_b200 import csr, synthetic
from traversome_b200.utils import Criterion, aic, bic


def _log0(v):
    return math.log(v) if v > 0 else float("-inf")


def test_get_like_formula_mirror_bit_exact(plastome, synthetic_small):
    """The mirror evaluated with floats equals the reference's own numbers bit for bit."""
    cases = [(plastome["variant_sizes"], plastome["models"][0]["bins_list"], plastome["known_answer_ll"])]
    cases += [(c["variant_sizes"], c["bins_list"], c["known_answer_ll"]) for c in synthetic_small["cases"]]
    for L, bins, kas in cases:
        for ka in kas:
            model = model_from_tables(L, bins)
            theta = [hexf(t) for t in ka["theta"]]
            info = model.get_like_formula(theta, _log0, set(ka["subset"]) if ka["subset"] else None)
            ref = hexf(ka["loglike"])
            assert info.loglike_expression == ref or (math.isinf(ref) and info.loglike_expression == ref)
            assert info.variable_size == ka["variable_size"] and info.sample_size == ka["sample_size"]
            assert model.sample_size == ka["sample_size"]


def test_get_like_formula_prunes_in_place(synthetic_small):
    case = synthetic_small["cases"][1]           # has bins with X < 1
    model = model_from_tables(case["variant_sizes"], case["bins_list"])
    before = sum(len(b.rp_bins) for b in model.bins_list)
    model.get_like_formula([1.0 / len(case["variant_sizes"])] * len(case["variant_sizes"]), _log0)
    after = sum(len(b.rp_bins) for b in model.bins_list)
    n_small = sum(1 for bins in case["bins_list"] for b in bins["rp_bins"] if hexf(b["num_possible_X"]) < 1)
    assert n_small > 0 and before - after == n_small
    assert all(rp.num_possible_X >= 1 for b in model.bins_list for rp in b.rp_bins)


def test_csr_build_is_integer_exact(plastome, synthetic_small):
    cases = [(plastome["variant_sizes"], plastome["models"][0]["bins_list"])]
    cases += [(c["variant_sizes"], c["bins_list"]) for c in synthetic_small["cases"]]
    for L, bins in cases:
        model = model_from_tables(L, bins)
        matrix = csr.ReadPathMatrix(len(L))
        w_map, C, N = csr.lane_from_bins(model.bins_list, matrix)
        A_ref, w_ref, C_ref, N_ref = loglik.rows_from_bins(L, bins)
        assert np.array_equal(matrix.dense(), A_ref)
        w = np.zeros(matrix.R, dtype=np.int64)
        for r, n in w_map.items():
            w[r] = n
        assert np.array_equal(w, w_ref) and N == N_ref and C == C_ref
        row_ptr, col, val = matrix.arrays()
        assert row_ptr[0] == 0 and row_ptr[-1] == col.size == val.size
        for r in range(matrix.R):
            seg = col[row_ptr[r]:row_ptr[r + 1]]
            assert np.all(np.diff(seg) > 0)
        # the reference's readpath table: every from_variants entry is reproduced
        dense = matrix.dense()
        for b in bins:
            for rp in b["rp_bins"]:
                if hexf(rp["num_possible_X"]) < 1:
                    continue
                r = matrix.row_index({int(v): int(c) for v, c in rp["from_variants"]}, add=False)
                for v, c in rp["from_variants"]:
                    assert dense[r, v] == c


def test_lanes_from_models_share_rows_across_replicates(plastome):
    L = plastome["variant_sizes"]
    models = [model_from_tables(L, plastome["models"][i]["bins_list"]) for i in range(0, 12)]
    matrix, w, C, N = csr.lanes_from_models(models)
    assert w.shape == (matrix.R, 12) and w.dtype == np.int32
    for b, i in enumerate(range(0, 12)):
        A_ref, w_ref, C_ref, N_ref = loglik.rows_from_bins(L, plastome["models"][i]["bins_list"])
        assert N[b] == N_ref == int(w[:, b].sum()) and C[b] == C_ref
        # LL through the shared matrix == LL through the replicate's own tables
        th = np.array([0.61, 0.39])
        assert abs(loglik.loglik_rows(matrix.dense(), L, w[:, b], C[b], th) -
                   loglik.loglik_rows(A_ref, L, w_ref, C_ref, th)) < 1e-9


def test_cover_matches_reference_rule(synthetic_small):
    case = synthetic_small["cases"][1]
    model = model_from_tables(case["variant_sizes"], case["bins_list"])
    ctx = fit_context(model, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
    matrix = csr.ReadPathMatrix(model.num_put_variants)
    w_map, _, _ = csr.lane_from_bins(model.bins_list, matrix)
    observed = sorted(w_map)
    V = model.num_put_variants
    rng = np.random.default_rng(0)
    for _ in range(50):
        ids = set(int(v) for v in rng.choice(V, rng.integers(1, V + 1), replace=False))
        # reference rule (ModelFitMaxLike.py:568-580) on the dict tables
        model_sp = set()
        for v in ids:
            for sp_ in ctx["variant_subpath_counters"][ctx["variant_paths"][v]]:
                model_sp.add(ctx["sbp_to_sbp_id"][sp_])
        obs = {i for i, spi in enumerate(model.all_sub_paths.values()) if spi.mapped_records}
        assert matrix.covers(ids, observed) == obs.issubset(model_sp)


def test_criteria():
    assert aic(-10.0, 3) == 26.0
    assert bic(-10.0, 3, 100) == math.log(100) * 3 + 20.0
    assert Criterion.BIC == "BIC" and Criterion("AIC") is Criterion.AIC


def test_synthetic_workload_shape_and_determinism():
    a = synthetic.make_workload(16, 500, 0.2, seed=5)
    b = synthetic.make_workload(16, 500, 0.2, seed=5)
    assert np.array_equal(a.col, b.col) and np.array_equal(a.n_obs, b.n_obs)
    assert a.R <= 500 and a.n_obs.min() >= 1 and a.N == int(a.n_obs.sum())
    assert np.all(np.diff(a.row_ptr) >= 1)
    for r in range(a.R):
        assert np.all(np.diff(a.col[a.row_ptr[r]:a.row_ptr[r + 1]]) > 0)
    assert set(np.unique(a.val)) <= {1, 2, 3}
    w = synthetic.bootstrap_weights(a.n_obs, 5, seed=1)
    assert w.shape == (a.R, 5) and np.all(w.sum(0) == a.N) and np.array_equal(w[:, 0], a.n_obs)
