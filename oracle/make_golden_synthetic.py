"""TEST INFRASTRUCTURE ONLY -- second half of oracle/make_golden.py.

Builds REFERENCE model objects (traversome.utils.Bins/BinInfo/SubPathInfo,
traversome.ModelGenerator.PathMultinomialModel) directly from seeded arrays and fits them with the
reference's own ModelFitMaxLike (point_estimate + reverse_model_selection), then writes the
tables and answers to tests/golden/synthetic_small.json.gz.  Fits go through the sympy stand-in
for symengine (see oracle/symengine_standin.py) and are labelled so; np.random is seeded before
every reference call because the reference leaves it unseeded.
"""
import gzip
import json
import math
import os
from collections import OrderedDict
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
VIA = "reference + sympy stand-in for symengine"


def synth_tables(seed, R, V, density, n_true, reads_per_row, dup_pair=None, two_range_frac=0.0, prune_frac=0.0):
    """Seeded integer tables in the shape SURVEY.md 8(d) describes, small enough for the reference."""
    rng = np.random.default_rng(seed)
    mask = rng.random((R, V)) < density
    for r in np.where(mask.sum(1) == 0)[0]:
        mask[r, rng.integers(V)] = True
    A = rng.choice([1, 2, 3], p=[0.85, 0.12, 0.03], size=(R, V)) * mask
    L = rng.integers(200000, 600001, size=V)
    if dup_pair is not None:                      # mutually unidentifiable variants
        a, b = dup_pair
        A[:, b] = A[:, a]
        L[b] = L[a]
    X = rng.integers(1000, 50001, size=R).astype(np.float64)
    theta = np.zeros(V)
    idx = rng.choice(V, n_true, replace=False)
    theta[idx] = rng.dirichlet(0.5 * np.ones(n_true))
    x = theta / float(theta @ L)
    p = X * (A @ x)
    p = p / p.sum()
    n = rng.multinomial(reads_per_row * R, p)
    keep = n > 0
    A, n, X = A[keep], n[keep], X[keep]
    R = A.shape[0]
    # bins: range 0 holds every read path; a fraction of read paths gets a second bin in range 1
    bins = [{"min_len": 1, "max_len": 9999, "min_id": 0, "max_id": 0, "rp_bins": []},
            {"min_len": 10000, "max_len": 29999, "min_id": 0, "max_id": 0, "rp_bins": []}]
    for r in range(R):
        fv = [[int(v), int(A[r, v])] for v in np.where(A[r] > 0)[0]]
        n1 = int(n[r])
        if rng.random() < two_range_frac and n1 >= 2:
            n2 = int(rng.integers(1, n1))
            n1 -= n2
            bins[1]["rp_bins"].append({"row": r, "from_variants": fv, "num_matched": n2,
                                       "num_possible_X": float(X[r] * 0.37 + 3.25).hex()})
        bins[0]["rp_bins"].append({"row": r, "from_variants": fv, "num_matched": n1,
                                   "num_possible_X": float(X[r]).hex()})
    # bins with X < 1 must be pruned by get_like_formula (ModelGenerator.py:144)
    n_prune = int(round(prune_frac * R))
    for r in rng.choice(R, n_prune, replace=False):
        fv = [[int(v), int(A[r, v])] for v in np.where(A[r] > 0)[0]]
        bins[1]["rp_bins"].append({"row": int(r), "from_variants": fv, "num_matched": int(rng.integers(1, 4)),
                                   "num_possible_X": float(rng.random() * 0.9).hex()})
    bins = [b for b in bins if b["rp_bins"]]
    return {"seed": seed, "variant_sizes": [int(v) for v in L], "true_theta": [float(t) for t in theta],
            "n_rows": int(R), "bins_list": bins}


def build_reference_objects(tab):
    """JSON tables -> live reference objects (the constructor arguments of ModelFitMaxLike.py:66-74)."""
    from traversome.utils import Bins, BinInfo, SubPathInfo
    from traversome.ModelGenerator import PathMultinomialModel
    V = len(tab["variant_sizes"])
    R = tab["n_rows"]
    paths = [(("rp%d" % r, True),) for r in range(R)]
    all_sub_paths = OrderedDict()
    rec = 0
    matched = {r: 0 for r in range(R)}
    for bins in tab["bins_list"]:
        for b in bins["rp_bins"]:
            matched[b["row"]] += b["num_matched"]
    fv_of = {}
    for bins in tab["bins_list"]:
        for b in bins["rp_bins"]:
            fv_of.setdefault(b["row"], {int(v): int(c) for v, c in b["from_variants"]})
    for r in range(R):
        spi = SubPathInfo()
        spi.from_variants = fv_of[r]
        spi.mapped_records = list(range(rec, rec + matched[r]))
        rec += matched[r]
        spi.inner_len, spi.full_len = 0, 10 ** 6
        all_sub_paths[paths[r]] = spi
    bins_list = []
    for bins in tab["bins_list"]:
        bo = Bins(min_len=bins["min_len"], max_len=bins["max_len"], min_id=bins["min_id"], max_id=bins["max_id"])
        for b in bins["rp_bins"]:
            bi = BinInfo()
            bi.from_variants = all_sub_paths[paths[b["row"]]].from_variants      # shared dict (traversome.py:1364)
            bi.num_matched = b["num_matched"]
            bi.num_possible_X = float.fromhex(b["num_possible_X"])
            bo.rp_bins.append(bi)
        bins_list.append(bo)
    variant_paths = [(("var%d" % v, True),) for v in range(V)]
    counters = OrderedDict()
    for v in range(V):
        counters[variant_paths[v]] = {paths[r]: spi.from_variants[v]
                                      for r, spi in enumerate(all_sub_paths.values()) if v in spi.from_variants}
    # unidentifiable table, as traversome.py:1224-1243 builds it
    be_unid = OrderedDict()
    for rep in range(V):
        for chk in range(rep, V):
            if chk not in be_unid:
                if chk == rep:
                    be_unid[chk] = chk
                elif counters[variant_paths[chk]] == counters[variant_paths[rep]] and \
                        tab["variant_sizes"][chk] == tab["variant_sizes"][rep]:
                    be_unid[chk] = rep
    repr_to_merged = OrderedDict([(rid, []) for rid in sorted(set(be_unid.values()))])
    for chk, rep in be_unid.items():
        repr_to_merged[rep].append(chk)
    model = PathMultinomialModel(variant_sizes=list(tab["variant_sizes"]), variant_topos=[True] * V,
                                 bins_list=bins_list, all_sub_paths=all_sub_paths)
    sbp_to_sbp_id = {p: i for i, p in enumerate(all_sub_paths)}
    return dict(model=model, variant_paths=variant_paths, variant_subpath_counters=counters,
                sbp_to_sbp_id=sbp_to_sbp_id, repr_to_merged_variants=repr_to_merged, be_unidentifiable_to=be_unid)


def _log0(v):
    return math.log(v) if v > 0 else float("-inf")


def make_synthetic():
    from traversome.ModelFitMaxLike import ModelFitMaxLike
    from traversome.utils import setup_logger
    setup_logger(loglevel="ERROR", timed=False)
    cases = [
        dict(name="k4_all_present", seed=1101, R=300, V=4, density=0.5, n_true=4, reads_per_row=18),
        dict(name="k8_three_absent_dup", seed=1102, R=900, V=8, density=0.3, n_true=5, reads_per_row=22,
             dup_pair=(2, 6), two_range_frac=0.3, prune_frac=0.02),
        dict(name="k12_sparse_truth", seed=1103, R=1500, V=12, density=0.25, n_true=6, reads_per_row=20,
             two_range_frac=0.2),
    ]
    out = {"what": "reference objects from seeded arrays (SURVEY.md 4(ii)); " + VIA, "cases": []}
    for cs in cases:
        kw = {k: v for k, v in cs.items() if k != "name"}
        tab = synth_tables(**kw)
        tab["name"] = cs["name"]
        V = len(tab["variant_sizes"])
        objs = build_reference_objects(tab)
        tab["be_unidentifiable_to"] = [[int(k), int(v)] for k, v in objs["be_unidentifiable_to"].items()]
        tab["repr_to_merged_variants"] = [[int(k), [int(x) for x in v]] for k, v in objs["repr_to_merged_variants"].items()]
        # known-answer LLs straight from the reference (no stand-in involved)
        rng = np.random.default_rng(cs["seed"] + 7)
        ka = []
        reps = sorted(objs["repr_to_merged_variants"])
        for trial in range(4):
            th = rng.dirichlet(np.ones(V))
            subset = None if trial < 2 else set(int(v) for v in rng.choice(V, max(2, V // 2), replace=False))
            m = deepcopy(objs["model"])
            info = m.get_like_formula([float(t) for t in th], _log0, subset)
            ka.append({"theta": [float(t).hex() for t in th], "subset": sorted(subset) if subset else None,
                       "loglike": float(info.loglike_expression).hex(), "variable_size": int(info.variable_size),
                       "sample_size": int(info.sample_size)})
        tab["known_answer_ll"] = ka
        fits = []
        for crit in ("BIC", "AIC"):
            o = build_reference_objects(tab)
            np.random.seed(cs["seed"])
            mf = ModelFitMaxLike(loglevel="ERROR", **o)
            prop, ll, cv = mf.point_estimate(criterion=crit)
            fits.append({"kind": "point_estimate", "criterion": crit, "chosen_ids": None,
                         "prop": [[int(k), float(v)] for k, v in prop.items()], "loglike": float(ll),
                         "criterion_value": float(cv), "via": VIA})
            o = build_reference_objects(tab)
            np.random.seed(cs["seed"] + 1)
            mf = ModelFitMaxLike(loglevel="ERROR", **o)
            prop, ll, cv = mf.reverse_model_selection(n_proc=1, criterion=crit)
            fits.append({"kind": "reverse_model_selection", "criterion": crit, "np_random_seed": cs["seed"] + 1,
                         "prop": [[int(k), float(v)] for k, v in prop.items()], "loglike": float(ll),
                         "criterion_value": float(cv), "via": VIA})
            print(cs["name"], crit, "selected", sorted(prop), "LL %.8f" % ll, "crit %.8f" % cv)
        # a point estimate restricted to a subset (exercises be_unidentifiable_to mapping)
        o = build_reference_objects(tab)
        np.random.seed(cs["seed"] + 2)
        mf = ModelFitMaxLike(loglevel="ERROR", **o)
        sub = set(reps[: max(2, len(reps) // 2)])
        try:
            prop, ll, cv = mf.point_estimate(chosen_ids=set(sub), criterion="BIC")
            fits.append({"kind": "point_estimate", "criterion": "BIC", "chosen_ids": sorted(sub),
                         "prop": [[int(k), float(v)] for k, v in prop.items()], "loglike": float(ll),
                         "criterion_value": float(cv), "via": VIA})
        except Exception as e:   # subset may not cover every read path -> reference fails
            fits.append({"kind": "point_estimate", "criterion": "BIC", "chosen_ids": sorted(sub), "error": str(e)})
        tab["fits"] = fits
        out["cases"].append(tab)
    with gzip.open(os.path.join(GOLDEN, "synthetic_small.json.gz"), "wt") as h:
        json.dump(out, h, separators=(",", ":"))
    print("synthetic: wrote", len(out["cases"]), "cases")
