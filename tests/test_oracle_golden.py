"""CPU: the oracle is pinned against outputs of the reference itself (tests/golden, made by oracle/make_golden.py)."""
import math

import numpy as np
import pytest

from oracle import exact, loglik, slsqp_port
from tests.helpers import hexf


def test_known_answer_ll_plastome_bit_exact(plastome):
    m0 = plastome["models"][0]
    for ka in plastome["known_answer_ll"]:
        theta = [hexf(t) for t in ka["theta"]]
        subset = set(ka["subset"]) if ka["subset"] else None
        ll, k, n = loglik.loglik_bins(plastome["variant_sizes"], m0["bins_list"], theta, subset)
        ref = hexf(ka["loglike"])
        assert ll == ref or (math.isinf(ll) and math.isinf(ref) and ll < 0)     # bit for bit
        assert k == ka["variable_size"] and n == ka["sample_size"]


def test_known_answer_ll_synthetic_bit_exact(synthetic_small):
    for case in synthetic_small["cases"]:
        for ka in case["known_answer_ll"]:
            theta = [hexf(t) for t in ka["theta"]]
            subset = set(ka["subset"]) if ka["subset"] else None
            ll, k, n = loglik.loglik_bins(case["variant_sizes"], case["bins_list"], theta, subset)
            assert ll == hexf(ka["loglike"])
            assert k == ka["variable_size"] and n == ka["sample_size"]


def test_row_merged_form_matches_bin_form(plastome, synthetic_small):
    cases = [(plastome["variant_sizes"], plastome["models"][i]["bins_list"]) for i in (0, 1, 50)]
    cases += [(c["variant_sizes"], c["bins_list"]) for c in synthetic_small["cases"]]
    rng = np.random.default_rng(3)
    for L, bins in cases:
        A, w, C, N = loglik.rows_from_bins(L, bins)
        theta = rng.dirichlet(np.ones(len(L)))
        a = loglik.loglik_rows(A, L, w, C, theta)
        b, _, n = loglik.loglik_bins(L, bins, list(theta))
        assert n == N
        assert abs(a - b) <= 1e-12 * abs(b)


def test_exact_fit_plastome_is_23_over_35(plastome):
    A, w, C, N = loglik.rows_from_bins(plastome["variant_sizes"], plastome["models"][0]["bins_list"])
    r = exact.exact_fit(A, plastome["variant_sizes"], w)
    assert abs(r["theta"][0] - 23 / 35.) < 1e-12 and abs(r["theta"][1] - 12 / 35.) < 1e-12
    assert r["comp"] < 1e-13 and r["dual"] < 1e-12
    # the reference's own answer (SLSQP, ftol 1e-5) sits within its documented spread of the optimum
    ref = plastome["fits"][0]
    assert abs(ref["prop"][0][1] - r["theta"][0]) < 3e-4
    assert abs((r["loglike_no_C"] + C - ref["loglike"]) / ref["loglike"]) < 1e-9
    assert r["loglike_no_C"] + C >= ref["loglike"] - 1e-9        # exact optimum is never worse


def test_exact_fit_vs_reference_bootstrap_fits(plastome):
    worst_t, worst_l = 0.0, 0.0
    for fit in plastome["fits"][1:]:
        m = plastome["models"][fit["model"]]
        A, w, C, N = loglik.rows_from_bins(plastome["variant_sizes"], m["bins_list"])
        r = exact.exact_fit(A, plastome["variant_sizes"], w)
        ref = np.array([p for _, p in fit["prop"]])
        worst_t = max(worst_t, float(np.abs(r["theta"] - ref).max()))
        worst_l = max(worst_l, abs((r["loglike_no_C"] + C - fit["loglike"]) / fit["loglike"]))
        assert r["loglike_no_C"] + C >= fit["loglike"] - 1e-8
    assert worst_t < 5e-4      # the reference's SLSQP stops this far from the optimum
    assert worst_l < 1e-9


def test_exact_fit_vs_reference_synthetic_point_estimates(synthetic_small):
    for case in synthetic_small["cases"]:
        L = case["variant_sizes"]
        A, w, C, N = loglik.rows_from_bins(L, case["bins_list"])
        groups = {int(k): v for k, v in case["repr_to_merged_variants"]}
        for fit in case["fits"]:
            if fit["kind"] != "point_estimate" or "error" in fit:
                continue
            ids = sorted(groups) if fit["chosen_ids"] is None else sorted(fit["chosen_ids"])
            mask = np.zeros(len(L), bool)
            mask[ids] = True
            r = exact.exact_fit(A, L, w, mask=mask)
            got = {}
            for rep in ids:
                for member in groups[rep]:
                    got[member] = r["theta"][rep] / len(groups[rep])
            for cid, p in fit["prop"]:
                assert abs(got[cid] - p) < 5e-4
            ll = r["loglike_no_C"] + C
            # the reference's SLSQP (ftol 1e-5) stops up to ~1e-4 LL units BELOW the maximum: on case k8 its own
            # answer is 1.7e-9 relative under the exact optimum -- never above it.
            assert abs((ll - fit["loglike"]) / fit["loglike"]) < 1e-8
            assert ll >= fit["loglike"] - 1e-8


def test_slsqp_port_reproduces_reference_spread(plastome):
    A, w, C, N = loglik.rows_from_bins(plastome["variant_sizes"], plastome["models"][0]["bins_list"])
    fun, n = slsqp_port.neg_loglike_closure(A, plastome["variant_sizes"], w, C)
    res = slsqp_port.minimize_neg_likelihood_port(fun, n, rng=np.random.default_rng(11))
    assert res is not False
    assert abs(res.x[0] - 23 / 35.) < 5e-4
    assert abs((-res.fun - plastome["fits"][0]["loglike"]) / res.fun) < 1e-9


def test_exact_fit_rejects_uncovered_subset(plastome):
    A, w, C, N = loglik.rows_from_bins(plastome["variant_sizes"], plastome["models"][0]["bins_list"])
    with pytest.raises(ValueError):
        exact.exact_fit(A, plastome["variant_sizes"], w, mask=np.array([True, False]))
