"""TEST INFRASTRUCTURE ONLY -- exact maximiser of the reference's objective on the CPU (fp64).

The reference maximises  LL(theta) = sum_b n_b log(X_b sum_v theta_v f_bv / sum_v theta_v L_v)
subject to theta >= 0, sum theta = 1 with multistart SLSQP at ftol=1e-5 (ModelFitMaxLike.py:19-58),
which stops 1e-7..1e-4 away from the maximiser and is not reproducible run to run.
Parity of the GPU solver is therefore stated against the EXACT maximiser of the same objective
(oracle 4), computed here by a method deliberately different from the GPU's (L-BFGS in
sqrt space): EM fixed-point iterations (phi_v <- phi_v g_v, the fixed point of ModelGenerator.py's
objective in phi_v = theta_v L_v / sum(theta L)) followed by an active-set Newton polish, and
certified by its KKT residuals.

    g_v(phi) = (1/N) sum_r w_r M_rv / (M phi)_r ,  M_rv = A_rv / L_v
    KKT:  g_v = 1 on the support, g_v <= 1 off it  (sum phi = 1 follows).
"""
import numpy as np
import scipy.sparse as sp


def _grad(M, MT, w, N, phi):
    p = M @ phi
    with np.errstate(divide="ignore", invalid="ignore"):
        rr = np.where(w > 0, w / p, 0.0)
    return (MT @ rr) / N, p


def kkt_residuals(A, L, w, phi, mask=None):
    """(complementarity max_v phi_v |g_v - 1|, dual max_v (g_v - 1) over the subset)."""
    M = sp.csr_matrix(A, dtype=np.float64) @ sp.diags(1.0 / np.asarray(L, dtype=np.float64))
    w = np.asarray(w, dtype=np.float64)
    g, _ = _grad(M, sp.csr_matrix(M.T), w, w.sum(), phi)
    m = np.ones(len(phi), bool) if mask is None else np.asarray(mask, bool)
    return float(np.max(np.abs(phi * (g - 1.0))[m])), float(np.max((g - 1.0)[m]))


def exact_fit(A, L, w, mask=None, max_newton=300, verbose=False):
    """Returns dict(theta, phi, loglike_no_C, comp, dual, em_iters, newton_iters).

    A [R,V] ints (dense or scipy.sparse), L[V], w[R] counts, mask[V] bool (candidate subset).
    Raises ValueError when some observed row is covered by no variant of the subset (LL = -inf).

    Method: a few plain EM sweeps from the uniform point (strictly positive iterates), then
    Bertsekas' projected Newton (two-metric projection) on
        F(phi) = -(1/N) sum_r w_r log (M phi)_r + sum_v phi_v ,  phi >= 0,
    whose minimiser is the maximiser of the reference's constrained LL (the multiplier of
    sum(phi)=1 is exactly N).  Dense V x V Hessian: meant for V up to a few hundred."""
    A = sp.csr_matrix(A, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    V = A.shape[1]
    m = np.ones(V, bool) if mask is None else np.asarray(mask, bool)
    M = sp.csr_matrix(A @ sp.diags(1.0 / L))
    MT = sp.csr_matrix(M.T)
    N = w.sum()
    obs = w > 0
    cover = np.asarray((A[:, m] if m.any() else A[:, :0]).sum(axis=1)).ravel()
    if np.any(obs & (cover == 0)):
        raise ValueError("subset does not cover every observed read path: LL = -inf")
    wo = w[obs]

    def F_of(ph):
        pp = (M @ ph)[obs]
        if np.any(pp <= 0):
            return np.inf
        return -float(wo @ np.log(pp)) / N + ph.sum()

    phi = np.where(m, 1.0, 0.0)
    phi /= phi.sum()
    em_iters = 30
    for _ in range(em_iters):
        g, _p = _grad(M, MT, w, N, phi)
        phi = phi * g
    newton = 0
    for newton in range(1, max_newton + 1):
        g, p = _grad(M, MT, w, N, phi)
        grad = np.where(m, 1.0 - g, 0.0)
        proj = phi - np.maximum(phi - grad, 0.0)              # projected-gradient residual
        eps = min(1e-8, float(np.linalg.norm(proj)))
        bound = (~m) | ((phi <= eps) & (grad > 0))          # held at (moved to) zero
        comp = float(np.max(np.abs(phi * grad)))
        dual = float(np.max(-grad[m]))
        if verbose:
            print("newton", newton, "comp", comp, "dual", dual, "free", int((~bound).sum()))
        if comp <= 1e-16 and dual <= 1e-13 and not np.any(bound & (phi > 0)):
            break
        idx = np.where(~bound)[0]
        d = np.zeros(V)
        if idx.size:
            dw = np.where(obs, w / np.where(obs, p * p, 1.0), 0.0) / N
            Mf = M[:, idx]
            H = (Mf.T @ sp.diags(dw) @ Mf).toarray()
            H[np.diag_indices_from(H)] += 1e-15 * max(np.trace(H) / idx.size, 1e-300)
            try:
                d[idx] = np.linalg.solve(H, -grad[idx])
            except np.linalg.LinAlgError:
                d[idx] = -grad[idx]
        d[bound] = -phi[bound]                                 # bound coordinates go to exactly zero
        F0 = F_of(phi)
        t, moved = 1.0, False
        for _ in range(60):
            trial = np.maximum(phi + t * d, 0.0)
            trial[bound] = 0.0
            Ft = F_of(trial)
            # Armijo on the projected arc (Bertsekas 1982)
            decrease = float(grad[idx] @ (phi[idx] - trial[idx])) + float(grad[bound] @ (phi[bound] - trial[bound]))
            if Ft <= F0 - 1e-4 * decrease or (abs(Ft - F0) <= 4e-16 * abs(F0) and decrease >= 0):
                moved = True
                break
            t *= 0.5
        if not moved or np.array_equal(trial, phi):
            break
        phi = trial
    phi = phi / phi.sum()
    g, p = _grad(M, MT, w, N, phi)
    comp = float(np.max(np.abs(phi * (g - 1.0))))
    dual = float(np.max((g - 1.0)[m]))
    x = phi / L
    theta = x / x.sum()
    ll = float(wo @ np.log((A @ theta)[obs])) - N * np.log(float(L @ theta))
    return {"theta": theta, "phi": phi, "loglike_no_C": ll, "comp": comp, "dual": dual,
            "em_iters": em_iters, "newton_iters": newton}
