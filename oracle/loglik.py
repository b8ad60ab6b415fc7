"""TEST INFRASTRUCTURE ONLY -- log-likelihood restatements.

loglik_bins   plain-Python restatement of PathMultinomialModel.get_like_formula evaluated with
              floats (reference: traversome/ModelGenerator.py:110-178), same term order, so it
              reproduces the reference's float bit for bit.  Input is the JSON table form written by
              oracle/make_golden.py (`dump_bins_list`).
loglik_rows   numpy restatement on the row-merged form  LL = C + sum_r w_r log(sum_v A_rv th_v / sum_v L_v th_v)
              (SURVEY.md section 8 "Math contract").
"""
import math

import numpy as np


def _x(b):
    x = b["num_possible_X"]
    return float.fromhex(x) if isinstance(x, str) else float(x)


def loglik_bins(variant_sizes, bins_tables, theta, subset=None, log=None):
    """Returns (loglike, variable_size, sample_size).  ModelGenerator.py:125-176."""
    if log is None:
        def log(v):
            return math.log(v) if v > 0 else float("-inf")
    n_var = len(variant_sizes)
    if not subset or set(subset) == set(range(n_var)):        # :125-126
        subset = None
    total_length = 0
    for v, size in enumerate(variant_sizes):                   # :129-136
        if subset is None or v in subset:
            total_length += theta[v] * float(size)
    loglike = 0
    n_records = 0
    for bins in bins_tables:
        for b in bins["rp_bins"]:
            X = _x(b)
            if X < 1:                                          # :144 pruned bins
                continue
            weight = 0
            for v, occ in b["from_variants"]:                 # :159-166 (dict order)
                if subset is None or v in subset:
                    weight += theta[v] * occ
            loglike += log(weight * X / total_length) * b["num_matched"]   # :167-174
            n_records += b["num_matched"]
    return loglike, (len(subset) if subset else n_var), n_records


def rows_from_bins(variant_sizes, bins_tables):
    """Row-merged integer tables from the JSON bins form: (A dense int64 [R,V], w[R], C, N)."""
    n_var = len(variant_sizes)
    row_of, sigs, w = {}, [], []
    C, N = 0.0, 0
    for bins in bins_tables:
        for b in bins["rp_bins"]:
            X = _x(b)
            if X < 1:
                continue
            sig = tuple(sorted((int(v), int(c)) for v, c in b["from_variants"]))
            r = row_of.get(sig)
            if r is None:
                r = row_of[sig] = len(sigs)
                sigs.append(sig)
                w.append(0)
            w[r] += int(b["num_matched"])
            C += math.log(X) * int(b["num_matched"])
            N += int(b["num_matched"])
    A = np.zeros((len(sigs), n_var), dtype=np.int64)
    for r, sig in enumerate(sigs):
        for v, c in sig:
            A[r, v] = c
    return A, np.asarray(w, dtype=np.int64), C, N


def loglik_rows(A, L, w, C, theta, mask=None):
    """LL for one lane.  A [R,V] (dense ndarray or scipy.sparse), L[V], w[R], theta[V], mask[V] bool."""
    theta = np.asarray(theta, dtype=np.float64)
    if mask is not None:
        theta = np.where(np.asarray(mask, dtype=bool), theta, 0.0)
    p = A @ theta
    denom = float(np.dot(np.asarray(L, dtype=np.float64), theta))
    w = np.asarray(w)
    sel = w > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.sum(w[sel] * np.log(np.asarray(p)[sel]))
    return float(s - w.sum() * math.log(denom) + C)
