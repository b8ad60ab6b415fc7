"""Lane batching: many (replicate, candidate-subset) ML fits -> ONE call of the CUDA solver.

The reference fits one subset of one replicate at a time (ModelFitMaxLike.py:349-370 inside the
loops of :199-322, repeated per replicate by traversome.py:515-556).  Here every such fit is a
"lane" = (weight column, candidate subset); a `LaneContext` owns the read-path x variant CSR on one
GPU and turns a list of lane requests into one `tvs_set_lanes_indexed` + `tvs_fit`.  With a
`LaneSharder` the lanes of a batch are split over the ranks of a torch.distributed job (one process
per GPU) and only the per-lane results are gathered.
"""
import threading

import numpy as np

from . import _cabi
from .csr import ReadPathMatrix, lane_from_bins


class FitResult(object):
    """What the reference reads off scipy's OptimizeResult: ``x`` (proportions of the fitted ids in
    increasing id order), ``fun`` (= -loglike), ``success``."""
    __slots__ = ("x", "fun", "success", "status", "kkt", "nit")

    def __init__(self, x, fun, success, status, kkt, nit):
        self.x, self.fun, self.success, self.status, self.kkt, self.nit = x, fun, success, status, kkt, nit


class LaneRequest(object):
    """One lane: weight column `col` of the context, the theta-free constant C, candidate ids, and optionally a
    starting point `theta0` (proportions of `ids` in increasing id order; None = uniform)."""
    __slots__ = ("col", "C", "ids", "theta0")

    def __init__(self, col, C, ids, theta0=None):
        self.col, self.C, self.ids = int(col), float(C), sorted(int(v) for v in ids)
        self.theta0 = None if theta0 is None else np.asarray(theta0, dtype=np.float64)


class LaneSharder(object):
    """Contiguous split of a batch of lanes over the ranks of a process group + result gather.

    No data-path collective: every rank holds the whole CSR and fits its own slice; one
    all_gather of {theta, loglike, kkt, iters, status} per batch (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    @staticmethod
    def bounds(B, rank, world):
        """Lanes [lo, hi) of `rank`: the first B % world ranks get one extra lane."""
        base, extra = divmod(int(B), int(world))
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)

    def my_slice(self, B):
        return self.bounds(B, self.rank, self.world)

    def gather(self, local, B, V):
        """local: dict of this rank's arrays (theta [V,b], loglike [b], kkt [b,2], iters [b], status [b]);
        returns the same dict for all B lanes, identical on every rank."""
        import torch
        dist = self._dist
        backend = dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        cap = (B + self.world - 1) // self.world
        width = V + 5                                     # theta | loglike | kkt0 | kkt1 | iters | status
        buf = torch.zeros((cap, width), dtype=torch.float64)
        lo, hi = self.my_slice(B)
        n = hi - lo
        if n:
            buf[:n, :V] = torch.from_numpy(np.ascontiguousarray(local["theta"].T))
            buf[:n, V] = torch.from_numpy(np.asarray(local["loglike"], dtype=np.float64))
            buf[:n, V + 1:V + 3] = torch.from_numpy(np.asarray(local["kkt"], dtype=np.float64))
            buf[:n, V + 3] = torch.from_numpy(np.asarray(local["iters"], dtype=np.float64))
            buf[:n, V + 4] = torch.from_numpy(np.asarray(local["status"], dtype=np.float64))
        buf = buf.to(dev)
        out = torch.empty((self.world * cap, width), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.cpu().numpy().reshape(self.world, cap, width)
        rows = []
        for r in range(self.world):
            a, b = self.bounds(B, r, self.world)
            rows.append(out[r, :b - a])
        full = np.concatenate(rows, axis=0)
        return {"theta": np.ascontiguousarray(full[:, :V].T), "loglike": full[:, V].copy(),
                "kkt": full[:, V + 1:V + 3].copy(), "iters": full[:, V + 3].astype(np.int32),
                "status": full[:, V + 4].astype(np.int32)}


class LaneContext(object):
    """The shared CSR (one `tvs_create`) + registered weight columns of every model that is fitted
    against it (the full data and each bootstrap replicate)."""

    # a batch of at least this many local lanes is cut in two halves that run concurrently on two handles (two CUDA
    # streams, two host threads): the narrow tail of one half overlaps the wide phase of the other.  Measured on cfg2:
    # 4000 lanes 40.9 -> 35.7 ms (1.15x), no gain at 1000 lanes (tools/two_handles.py).  0 disables.
    concurrent_min_lanes = 3000

    def __init__(self, n_variants, variant_sizes, device=0, sharder=None, tol=1e-9, max_iter=2000, history=16):
        self.V = int(n_variants)
        self.L = np.asarray(variant_sizes, dtype=np.int64)
        self.device = int(device)
        self.sharder = sharder
        self.tol, self.max_iter, self.history = tol, max_iter, history
        self.matrix = ReadPathMatrix(self.V)
        self._cols = []               # list of (rows, counts): sparse weight columns
        self._engine = None
        self._engine2 = None          # second handle for concurrent halves (created on first use)
        self._engine_R = -1
        self.n_lanes_fitted = 0
        self.n_batches = 0
        self.engine_factory = _cabi.Engine     # tests of the host logic inject a stand-in here

    # ---- models -> weight columns
    def add_model(self, model):
        """Register one model's bins; returns (col, C, N) of its lane (pruning rule of
        ModelGenerator.py:138-151 applied, the model itself is not modified here)."""
        if self.matrix.R == 0 and getattr(model, "all_sub_paths", None):
            self.matrix.add_sub_paths(model.all_sub_paths)
        w_map, C, N = lane_from_bins(model.bins_list, self.matrix)
        self._cols.append(self._sparse_column(w_map))
        return len(self._cols) - 1, float(C), int(N)

    def add_weights(self, w):
        """Register a dense weight column w[R] (synthetic workloads)."""
        w = np.asarray(w)
        rows = np.nonzero(w)[0]
        self._cols.append((rows.astype(np.int64), w[rows].astype(np.int32)))
        return len(self._cols) - 1

    @staticmethod
    def _sparse_column(w_map):
        """{row: count} -> (rows int64[], counts int32[]): columns are scattered with numpy, not Python loops."""
        rows = np.fromiter(w_map.keys(), dtype=np.int64, count=len(w_map))
        vals = np.fromiter(w_map.values(), dtype=np.int32, count=len(w_map))
        return rows, vals

    def engine(self):
        if self._engine is None or self._engine_R != self.matrix.R:
            if self._engine is not None:
                self._engine.close()
            if self._engine2 is not None:
                self._engine2.close()
                self._engine2 = None
            if self.matrix.R == 0:
                raise Exception("Likelihood maximization failed.")        # nothing to fit
            row_ptr, col, val = self.matrix.arrays()
            self._engine = self.engine_factory(row_ptr, col, val, self.L, n_variants=self.V, device=self.device)
            self._engine_R = self.matrix.R
        return self._engine

    def second_engine(self):
        self.engine()
        if self._engine2 is None:
            row_ptr, col, val = self.matrix.arrays()
            self._engine2 = self.engine_factory(row_ptr, col, val, self.L, n_variants=self.V, device=self.device)
        return self._engine2

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        if self._engine2 is not None:
            self._engine2.close()
            self._engine2 = None

    def dense_columns(self, cols):
        w = np.zeros((self.matrix.R, len(cols)), dtype=np.int32)
        for j, c in enumerate(cols):
            rows, vals = self._cols[c]
            w[rows, j] = vals
        return w

    # ---- the batched fit
    def fit(self, requests):
        """requests: list of LaneRequest -> list of FitResult (same order)."""
        B = len(requests)
        if B == 0:
            return []
        lo, hi = (0, B) if self.sharder is None else self.sharder.my_slice(B)
        local = None
        if hi > lo and self.concurrent_min_lanes and hi - lo >= self.concurrent_min_lanes:
            mid = lo + (hi - lo) // 2
            halves = [None, None]

            def work(i, a, b, eng):
                try:
                    halves[i] = self._fit_on(eng, requests[a:b])
                except BaseException as e:                      # re-raised in the calling thread
                    halves[i] = e
            t = threading.Thread(target=work, args=(1, mid, hi, self.second_engine()))
            t.start()
            work(0, lo, mid, self.engine())
            t.join()
            for h in halves:
                if isinstance(h, BaseException):
                    raise h
            local = {k: np.concatenate([halves[0][k], halves[1][k]], axis=(1 if k == "theta" else 0)) for k in halves[0]}
        elif hi > lo:
            local = self._fit_on(self.engine(), requests[lo:hi])
        out = local if self.sharder is None else self.sharder.gather(local, B, self.V)
        self.n_lanes_fitted += B
        self.n_batches += 1
        results = []
        for b, rq in enumerate(requests):
            status = int(out["status"][b])
            ll = float(out["loglike"][b])
            ok = status in (_cabi.CONVERGED, _cabi.STALLED) and np.isfinite(ll)
            results.append(FitResult(x=np.array(out["theta"][rq.ids, b], dtype=np.float64), fun=-ll, success=bool(ok),
                                     status=status, kkt=tuple(out["kkt"][b]), nit=int(out["iters"][b])))
        return results

    def _fit_on(self, eng, mine):
        """Fit the lanes `mine` on one handle; returns the engine's result dict."""
        n = len(mine)
        used = sorted({rq.col for rq in mine})
        pos = {c: j for j, c in enumerate(used)}
        w_cols = self.dense_columns(used)
        col_of_lane = np.array([pos[rq.col] for rq in mine], dtype=np.int32)
        mask = np.zeros((self.V, n), dtype=np.uint8)
        for b, rq in enumerate(mine):
            mask[rq.ids, b] = 1
        eng.set_lanes_indexed(w_cols, col_of_lane, mask=mask, C=np.array([rq.C for rq in mine]))
        theta0 = None
        if any(rq.theta0 is not None for rq in mine):
            theta0 = np.zeros((self.V, n), dtype=np.float64)
            for b, rq in enumerate(mine):
                theta0[rq.ids, b] = rq.theta0 if rq.theta0 is not None else 1.0 / len(rq.ids)
        return eng.fit(theta0=theta0, tol=self.tol, max_iter=self.max_iter, history=self.history)

    def loglik(self, request, theta_of_ids):
        """-LL at given proportions of request.ids (single lane)."""
        theta = np.zeros((self.V, 1), dtype=np.float64)
        theta[request.ids, 0] = theta_of_ids
        mask = np.zeros((self.V, 1), dtype=np.uint8)
        mask[request.ids, 0] = 1
        eng = self.engine()
        eng.set_lanes(self.dense_columns([request.col]), mask=mask, C=np.array([request.C]))
        with np.errstate(invalid="ignore"):
            return -eng.loglik(theta)
