"""Seeded synthetic read-path x variant workloads of the BASELINE.json shapes.

A in N^{R x V}: Bernoulli(density) pattern x values {1,2,3} w.p. (0.85, 0.12, 0.03), every row forced
to >= 1 non-zero; L_v ~ U{200000..600000}; X_r ~ U{1000..50000} (one bin per read path);
true theta ~ Dirichlet(0.5) on `n_true` of the V variants; observed counts
n ~ Multinomial(reads_per_row * R, p) with p ~ X * A (theta / theta.L); rows with n = 0 removed.
There is no network and the reference ships no data, so every benchmark input is made here.
"""
import math

import numpy as np

CONFIGS = {
    # name: (V, R, density, n_true, lanes)  -- BASELINE.json configs[1..4]
    "cfg2": dict(V=64, R=20000, density=0.10, n_true=None, lanes=1000),
    "cfg3": dict(V=256, R=50000, density=0.10, n_true=32, lanes=256),
    "cfg4": dict(V=2048, R=200000, density=0.05, n_true=None, lanes=10000),
    "cfg5": dict(V=4096, R=100000, density=0.30, n_true=None, lanes=256),
}


class Workload(object):
    __slots__ = ("row_ptr", "col", "val", "L", "X", "n_obs", "theta_true", "V", "R", "nnz", "C", "N", "seed")

    def dense(self):
        a = np.zeros((self.R, self.V), dtype=np.int64)
        rows = np.repeat(np.arange(self.R), np.diff(self.row_ptr))
        a[rows, self.col] = self.val
        return a

    def scipy_csr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.astype(np.float64), self.col, self.row_ptr), shape=(self.R, self.V))


def make_workload(V, R, density, seed, n_true=None, reads_per_row=10, block_rows=8192):
    rng = np.random.default_rng(seed)
    L = rng.integers(200000, 600001, size=V).astype(np.int64)
    theta = np.zeros(V)
    k = V if n_true is None else int(n_true)
    idx = np.sort(rng.choice(V, k, replace=False)) if k < V else np.arange(V)
    theta[idx] = rng.dirichlet(0.5 * np.ones(k))
    x = theta / float(theta @ L)
    cols, vals, counts, Xs, ps = [], [], [], [], []
    for r0 in range(0, R, block_rows):
        nr = min(block_rows, R - r0)
        pat = rng.random((nr, V)) < density
        empty = np.where(pat.sum(1) == 0)[0]
        if empty.size:
            pat[empty, rng.integers(V, size=empty.size)] = True
        vv = rng.choice(np.array([1, 2, 3], dtype=np.int32), p=[0.85, 0.12, 0.03], size=(nr, V))
        vv *= pat
        rr, cc = np.nonzero(vv)                    # row-major, columns increasing within a row
        cols.append(cc.astype(np.int32))
        vals.append(vv[rr, cc].astype(np.int32))
        counts.append(pat.sum(1).astype(np.int64))
        X = rng.integers(1000, 50001, size=nr).astype(np.float64)
        Xs.append(X)
        ps.append(X * (vv @ x))
    p = np.concatenate(ps)
    p /= p.sum()
    n = rng.multinomial(int(reads_per_row) * R, p)
    keep = n > 0
    counts = np.concatenate(counts)
    col = np.concatenate(cols)
    val = np.concatenate(vals)
    ent_keep = np.repeat(keep, counts)
    wl = Workload()
    wl.col = np.ascontiguousarray(col[ent_keep])
    wl.val = np.ascontiguousarray(val[ent_keep])
    kc = counts[keep]
    wl.row_ptr = np.zeros(kc.size + 1, dtype=np.int32)
    np.cumsum(kc, out=wl.row_ptr[1:])
    wl.L = L
    wl.X = np.concatenate(Xs)[keep]
    wl.n_obs = n[keep].astype(np.int32)
    wl.theta_true = theta
    wl.V, wl.R, wl.nnz = int(V), int(kc.size), int(wl.col.size)
    wl.N = int(wl.n_obs.sum())
    wl.C = float(np.sum(wl.n_obs * np.log(wl.X)))
    wl.seed = seed
    return wl


def make_config(name, seed=None, scale=1.0):
    """BASELINE.json config by name; `scale` < 1 shrinks R (parity-test sized variants)."""
    cfg = CONFIGS[name]
    if seed is None:
        seed = 1000 + int(name[-1])
    return make_workload(cfg["V"], max(64, int(cfg["R"] * scale)), cfg["density"], seed, n_true=cfg["n_true"])


def bootstrap_weights(n_obs, B, seed, keep_lane0=True):
    """Host-side multinomial bootstrap lanes w[R,B] int32 (numpy Generator; for tests and e2e inputs)."""
    rng = np.random.default_rng(seed)
    n_obs = np.asarray(n_obs, dtype=np.int64)
    N = int(n_obs.sum())
    w = rng.multinomial(N, n_obs / N, size=B).T.astype(np.int32)
    if keep_lane0:
        w[:, 0] = n_obs
    return np.ascontiguousarray(w)


def algorithmic_bytes_per_pass(R, V, nnz, B):
    """8*nnz + 4*(R+1) + 4*R*B + 16*V*B bytes per solver pass over B lanes."""
    return 8 * nnz + 4 * (R + 1) + 4 * R * B + 16 * V * B


def flops_per_pass(R, nnz, B):
    """4*nnz*B + ~6*R*B fp64 flops (+ R*B logs)."""
    return 4 * nnz * B + 6 * R * B
