"""TEST INFRASTRUCTURE ONLY -- the reference's optimiser driver, restated, as the timed CPU baseline.

`minimize_neg_likelihood_port` restates ModelFitMaxLike.minimize_neg_likelihood
(traversome/ModelFitMaxLike.py:19-58): scipy SLSQP, finite-difference gradient (jac=False),
ftol=1e-5, eps=1e-8, maxiter=1000, equality constraint sum(x)=1, bounds [0,1], uniform random
normalised starts, repeated until 6 successful runs (at most 10,000 starts), best `fun` wins.

`neg_loglike_closure` is the numpy closure of the reference objective
(ModelGenerator.py:153-174) that stands in for the symengine/LLVM-compiled callable
(ModelFitMaxLike.py:534-545).  It omits the reference's per-fit symbolic build + LLVM JIT, so a
baseline timed with it is FASTER than the true reference.
"""
import numpy as np
import scipy.sparse as sp
from scipy import optimize


def neg_loglike_closure(A, L, w, C=0.0, subset=None):
    """x (proportions of the subset's variants, in increasing variant id) -> -LL, scipy style."""
    A = sp.csr_matrix(A, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    ids = np.arange(A.shape[1]) if subset is None else np.asarray(sorted(subset))
    As = sp.csr_matrix(A[:, ids])
    Ls = L[ids]
    sel = w > 0
    As = sp.csr_matrix(As[sel])
    ws = w[sel]
    N = ws.sum()

    def fun(x):
        x = np.asarray(x, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            return -(float(ws @ np.log(As @ x)) - N * np.log(float(Ls @ x)) + C)
    return fun, len(ids)


def minimize_neg_likelihood_port(neg_loglike_func, num_variables, rng=None, n_success=6, max_starts=10000):
    """ModelFitMaxLike.py:19-58.  `rng`: numpy Generator/RandomState-like with .random(n); default = the
    global np.random (unseeded, like the reference)."""
    draw = np.random.random if rng is None else rng.random
    constraints = ({"type": "eq", "fun": lambda x: sum(x) - 1})
    options = {"disp": False, "maxiter": 1000, "ftol": 1.0e-5, "eps": 1.0e-8}
    good = []
    starts = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        while starts < max_starts:
            x0 = draw(num_variables)
            x0 /= sum(x0)
            res = optimize.minimize(fun=neg_loglike_func, x0=x0, jac=False, method="SLSQP",
                                    bounds=[(0., 1.0)] * num_variables, constraints=constraints, options=options)
            if res.success:
                good.append(res)
                if len(good) >= n_success:
                    break
            starts += 1
    if not good:
        return False
    return sorted(good, key=lambda r: r.fun)[0]
