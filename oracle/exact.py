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


def exact_fit(A, L, w, mask=None, em_tol=1e-9, max_em=200000, verbose=False):
    """Returns dict(theta, phi, loglike_no_C, comp, dual, em_iters, newton_iters).

    A [R,V] ints (dense or scipy.sparse), L[V], w[R] counts, mask[V] bool (candidate subset).
    Raises ValueError when some observed row is covered by no variant of the subset (LL = -inf)."""
    A = sp.csr_matrix(A, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    V = A.shape[1]
    m = np.ones(V, bool) if mask is None else np.asarray(mask, bool)
    M = sp.csr_matrix(A @ sp.diags(1.0 / L))
    MT = sp.csr_matrix(M.T)
    N = w.sum()
    cover = np.asarray((A[:, m] if m.any() else A[:, :0]).sum(axis=1)).ravel()
    if np.any((w > 0) & (cover == 0)):
        raise ValueError("subset does not cover every observed read path: LL = -inf")
    phi = np.where(m, 1.0, 0.0)
    phi /= phi.sum()
    # ---- EM with SQUAREM extrapolation (safeguarded by the log-likelihood)
    def ll_of(p):
        sel = w > 0
        return float(np.sum(w[sel] * np.log(p[sel])))
    it = 0
    g, p = _grad(M, MT, w, N, phi)
    while it < max_em:
        phi1 = phi * g
        g1, p1 = _grad(M, MT, w, N, phi1)
        phi2 = phi1 * g1
        it += 2
        r = phi1 - phi
        v = (phi2 - phi1) - r
        rn, vn = np.sqrt(r @ r), np.sqrt(v @ v)
        cand = phi2
        if vn > 0:
            a = max(1.0, rn / vn)
            if a > 1.0:
                ext = phi + 2 * a * r + a * a * v
                ext = np.where(m, np.maximum(ext, 0.0), 0.0)
                s = ext.sum()
                if s > 0:
                    ext /= s
                    pe = M @ ext
                    if np.all(pe[w > 0] > 0) and ll_of(pe) >= ll_of(M @ phi2):
                        cand = ext
        phi = cand / cand.sum()
        g, p = _grad(M, MT, w, N, phi)
        it += 1
        comp = np.max(np.abs(phi * (g - 1.0)))
        if verbose and it % 300 == 0:
            print("em", it, comp)
        if comp < em_tol:
            break
    # ---- active-set Newton polish on F(phi) = -(1/N) sum w log(M phi) + sum phi, phi >= 0
    newton = 0
    for newton in range(1, 60):
        g, p = _grad(M, MT, w, N, phi)
        grad = 1.0 - g
        free = m & ((phi > 1e-14) | (grad < 0))
        phi = np.where(free, phi, 0.0)
        if np.max(np.abs(phi * grad)) < 1e-15 and np.max(-grad[m & ~free], initial=0.0) <= 0:
            break
        idx = np.where(free)[0]
        d = np.where(w > 0, w / (p * p), 0.0) / N
        Mf = M[:, idx]
        H = (Mf.T @ sp.diags(d) @ Mf).toarray()
        H += 1e-14 * np.trace(H) / max(len(idx), 1) * np.eye(len(idx))
        try:
            step = np.linalg.solve(H, -grad[idx])
        except np.linalg.LinAlgError:
            break
        t = 1.0
        neg = step < 0
        if neg.any():
            t = min(1.0, 0.999999 * np.min(-phi[idx][neg] / step[neg]))
        # backtrack on F
        def F(ph):
            pp = M @ ph
            if np.any(pp[w > 0] <= 0):
                return np.inf
            return -ll_of(pp) / N + ph.sum()
        F0 = F(phi)
        ok = False
        for _ in range(40):
            trial = phi.copy()
            trial[idx] = np.maximum(phi[idx] + t * step, 0.0)
            if F(trial) <= F0 + 1e-15 * abs(F0):
                ok = True
                break
            t *= 0.5
        if not ok:
            break
        # drop numerically-zero coordinates whose gradient says "stay at the bound"
        phi = trial
    phi = np.where(phi < 1e-300, 0.0, phi)
    phi /= phi.sum()
    g, p = _grad(M, MT, w, N, phi)
    comp = float(np.max(np.abs(phi * (g - 1.0))))
    dual = float(np.max((g - 1.0)[m]))
    x = phi / L
    theta = x / x.sum()
    return {"theta": theta, "phi": phi, "loglike_no_C": ll_of(A @ theta) - N * np.log(float(L @ theta)),
            "comp": comp, "dual": dual, "em_iters": it, "newton_iters": newton}
