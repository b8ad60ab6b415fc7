"""Lane batching: many (replicate, candidate-subset) ML fits -> ONE call of the CUDA solver.

The reference fits one subset of one replicate at a time (ModelFitMaxLike.py:349-370 inside the
loops of :199-322, repeated per replicate by traversome.py:515-556).  Here every such fit is a
"lane" = (weight column, candidate subset); a `LaneContext` owns the read-path x variant CSR on one
GPU, keeps the weight columns RESIDENT on the device (registered once, named by slot afterwards:
`tvs_columns_*`, `tvs_set_lanes_columns`) and turns a batch of lanes into one `tvs_fit`.  With a
`LaneSharder` the lanes of a batch are split over the ranks of a torch.distributed job (one process
per GPU); every rank fits its slice and ONE all_gather of the per-lane results -- straight from the
device buffers `tvs_fit` wrote -- returns all of them to all ranks.
"""
import sys
import threading
import time
import zlib

import numpy as np

from . import _cabi
from .csr import CsrMatrix, ReadPathMatrix, lane_from_bins


class FitResult(object):
    """What the reference reads off scipy's OptimizeResult: ``x`` (proportions of the fitted ids in
    increasing id order), ``fun`` (= -loglike), ``success``.  ``x`` is cut out of the batch's theta matrix on first
    use: a selection round reads it for the accepted candidate only."""
    __slots__ = ("_x", "_theta", "_ids", "_b", "fun", "success", "status", "kkt", "nit")

    def __init__(self, x, fun, success, status, kkt, nit, theta=None, ids=None, b=0):
        self._x, self._theta, self._ids, self._b = x, theta, ids, b
        self.fun, self.success, self.status, self.kkt, self.nit = fun, success, status, kkt, nit

    @property
    def x(self):
        if self._x is None:
            self._x = np.array(self._theta[self._ids, self._b], dtype=np.float64)
        return self._x


class IdSet(object):
    """A candidate subset as a SORTED int64 array: what the selection loop hands to `lane_requests` (a leave-one-out
    subset is the accepted set with one element deleted -- no per-lane sort, no per-lane Python set)."""
    __slots__ = ("arr",)

    def __init__(self, arr):
        self.arr = arr

    def __iter__(self):
        return iter(self.arr.tolist())

    def __len__(self):
        return int(self.arr.shape[0])

    def __contains__(self, v):
        i = int(np.searchsorted(self.arr, v))
        return i < self.arr.shape[0] and int(self.arr[i]) == int(v)


class LeaveOneOut(object):
    """The candidate subsets of one chunk of a backward-selection round: the accepted set (sorted array) minus ONE of its
    elements each (positions `drop_pos`).  Behaves like a sequence of IdSet, but the lane builder takes it whole: masks and
    starting points of the chunk are built with array operations, not lane by lane."""
    __slots__ = ("chosen", "drop_pos")

    def __init__(self, chosen, drop_pos):
        self.chosen = np.ascontiguousarray(chosen, dtype=np.int64)
        self.drop_pos = np.ascontiguousarray(drop_pos, dtype=np.int64)

    def __len__(self):
        return int(self.drop_pos.shape[0])

    def __getitem__(self, j):
        return IdSet(np.delete(self.chosen, int(self.drop_pos[j])))

    def __iter__(self):
        return (self[j] for j in range(len(self)))


class LaneBlock(object):
    """The lanes of one LeaveOneOut chunk of one weight column: `col`, `C`, the subsets, and the accepted model's optimum
    `warm` (dense [V] proportions, or None for uniform starts)."""
    __slots__ = ("col", "C", "subsets", "warm")

    def __init__(self, col, C, subsets, warm):
        self.col, self.C, self.subsets, self.warm = int(col), float(C), subsets, warm

    def __len__(self):
        return len(self.subsets)

    def arrays(self, V):
        """(cols [n], C [n], mask uint8 [V, n], theta0 float64 [V, n] or None) -- same starts as
        ModelFitMaxLike._start_point: the accepted optimum restricted to the subset, floored at 1e-3 / |subset|."""
        lo = self.subsets
        n, K = len(lo), int(lo.chosen.shape[0]) - 1
        base = np.zeros(V, dtype=np.uint8)
        base[lo.chosen] = 1
        mask = np.repeat(base[:, None], n, axis=1)
        dropped = lo.chosen[lo.drop_pos]
        lane = np.arange(n)
        mask[dropped, lane] = 0
        theta0 = None
        if self.warm is not None and K > 0:
            T = np.repeat(self.warm[:, None], n, axis=1)
            T[dropped, lane] = 0.0
            T *= mask
            tot = T.sum(axis=0)
            bad = ~np.isfinite(tot) | (tot <= 0.0)
            T /= np.where(bad, 1.0, tot)
            T = np.where(mask != 0, np.maximum(T, 1e-3 / K), 0.0)
            T /= T.sum(axis=0)
            if bad.any():
                T[:, bad] = mask[:, bad] / float(K)
            theta0 = T
        return np.full(n, self.col, dtype=np.int64), np.full(n, self.C, dtype=np.float64), mask, theta0


class BlockResult(object):
    """Results of one LaneBlock: `fun` (= -loglike) and `success` per lane as arrays, proportions of lane j on demand."""
    __slots__ = ("fun", "success", "status", "_theta", "_lo", "_subsets")

    def __init__(self, fun, success, status, theta, lo, subsets):
        self.fun, self.success, self.status, self._theta, self._lo, self._subsets = fun, success, status, theta, lo, subsets

    def __len__(self):
        return int(self.fun.shape[0])

    def ids(self, j):
        return np.delete(self._subsets.chosen, int(self._subsets.drop_pos[j]))

    def x(self, j):
        return np.array(self._theta[self.ids(j), self._lo + j], dtype=np.float64)

    def __getitem__(self, j):
        """Lane j as a FitResult (compatibility with the per-lane protocol)."""
        return FitResult(None, float(self.fun[j]), bool(self.success[j]), int(self.status[j]), (0.0, 0.0), 0,
                         theta=self._theta, ids=self.ids(j), b=self._lo + j)


class LaneRequest(object):
    """One lane: weight column `col` of the context, the theta-free constant C, candidate ids, and optionally a
    starting point `theta0` (proportions of `ids` in increasing id order; None = uniform)."""
    __slots__ = ("col", "C", "ids", "theta0")

    def __init__(self, col, C, ids, theta0=None):
        self.col, self.C = int(col), float(C)
        self.ids = ids.arr if isinstance(ids, IdSet) else sorted(int(v) for v in ids)
        self.theta0 = None if theta0 is None else np.asarray(theta0, dtype=np.float64)


class LaneBatch(object):
    """Array form of a list of lanes: `cols` int64[B], `C` float64[B], `mask` uint8[V,B] or None (= every variant in
    every lane), `theta0` float64[V,B] or None (= uniform starts).  Built once when the same batch is fitted again and
    again (bootstrap of a fixed candidate set), otherwise by `from_requests`."""
    __slots__ = ("cols", "C", "mask", "theta0")

    def __init__(self, cols, C=None, mask=None, theta0=None):
        self.cols = np.ascontiguousarray(cols, dtype=np.int64)
        B = self.cols.shape[0]
        self.C = np.zeros(B, dtype=np.float64) if C is None else np.ascontiguousarray(C, dtype=np.float64)
        self.mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.theta0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)

    def __len__(self):
        return int(self.cols.shape[0])

    @classmethod
    def from_requests(cls, requests, V):
        B = len(requests)
        cols = np.fromiter((rq.col for rq in requests), dtype=np.int64, count=B)
        C = np.fromiter((rq.C for rq in requests), dtype=np.float64, count=B)
        sizes = np.fromiter((len(rq.ids) for rq in requests), dtype=np.int64, count=B)
        mask = None
        flat = lane = None
        if B and (sizes != V).any():
            # one scatter for the whole batch instead of a Python loop over lanes
            flat = np.concatenate([np.asarray(rq.ids, dtype=np.int64) for rq in requests]) if sizes.sum() else np.zeros(0, np.int64)
            lane = np.repeat(np.arange(B, dtype=np.int64), sizes)
            mask = np.zeros((V, B), dtype=np.uint8)
            mask[flat, lane] = 1
        theta0 = None
        if any(rq.theta0 is not None for rq in requests):
            if flat is None:
                flat = np.concatenate([np.asarray(rq.ids, dtype=np.int64) for rq in requests])
                lane = np.repeat(np.arange(B, dtype=np.int64), sizes)
            vals = np.concatenate([rq.theta0 if rq.theta0 is not None else np.full(len(rq.ids), 1.0 / max(len(rq.ids), 1))
                                   for rq in requests])
            theta0 = np.zeros((V, B), dtype=np.float64)
            theta0[flat, lane] = vals
        return cls(cols, C, mask, theta0)

    def slice(self, lo, hi):
        return LaneBatch(self.cols[lo:hi], self.C[lo:hi], None if self.mask is None else self.mask[:, lo:hi],
                         None if self.theta0 is None else self.theta0[:, lo:hi])

    def digest(self):
        """32-bit digest of what the lanes ARE (columns, subsets): ranks of a sharded job compare it before they split a
        batch by position."""
        d = zlib.crc32(self.cols.tobytes())
        if self.mask is not None:
            d = zlib.crc32(np.packbits(self.mask, axis=0).tobytes(), d)
        return d & 0x7fffffff


_RESULT_KEYS = ("theta", "loglike", "kkt", "iters", "status")


class ShardedTheta(object):
    """theta [V, B] of a gathered batch kept as the per-rank blocks the collective delivered (no [V, B] copy on the host:
    at 8 ranks x 1,000 lanes x 64 variants that copy is a third of the gather's cost).  Supports what the callers use:
    `theta[ids, b]` / `theta[:, b]` (one lane), `.shape`, and `np.asarray(theta)` / slicing by lanes (materialises)."""
    __slots__ = ("blocks", "starts", "shape", "_full")

    def __init__(self, blocks, starts, V, B):
        self.blocks, self.starts, self.shape, self._full = blocks, np.asarray(starts, dtype=np.int64), (int(V), int(B)), None

    def full(self):
        if self._full is None:
            self._full = np.ascontiguousarray(np.concatenate(self.blocks, axis=1)) if self.blocks else np.zeros(self.shape)
        return self._full

    def __array__(self, dtype=None, copy=None):
        a = self.full()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, key):
        if self._full is None and isinstance(key, tuple) and len(key) == 2 and isinstance(key[1], (int, np.integer)):
            b = int(key[1])
            if b < 0:
                b += self.shape[1]
            r = int(np.searchsorted(self.starts, b, side="right")) - 1
            return self.blocks[r][key[0], b - int(self.starts[r])]
        return self.full()[key]

    def sum(self, *a, **k):
        return self.full().sum(*a, **k)

    def __sub__(self, other):
        return self.full() - np.asarray(other)

    def __rsub__(self, other):
        return np.asarray(other) - self.full()

    def __mul__(self, other):
        return self.full() * np.asarray(other)

    __rmul__ = __mul__

    def __add__(self, other):
        return self.full() + np.asarray(other)

    __radd__ = __add__

    @property
    def T(self):
        return self.full().T

    def copy(self):
        return self.full().copy()


class LaneSharder(object):
    """Contiguous split of a batch of lanes over the ranks of a process group + result gather.

    No data-path collective: every rank holds the whole CSR and fits its own slice; one
    all_gather of {theta, loglike, kkt, iters, status} per batch (NCCL on GPUs, gloo in CPU tests).
    With NCCL the solver writes its results into views of ONE packed device buffer per rank (tvs_fit accepts device
    pointers), the collective runs on that buffer and a single device->host copy of the gathered block follows."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self._torch = torch
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        self._bufs = {}
        self.bytes_gathered = 0          # total bytes received by this rank's all_gathers (diagnostics / bench)
        self.n_gathers = 0
        self.seconds = {"pack": 0.0, "all_gather_to_host": 0.0, "unpack": 0.0}   # host time of this rank inside gather()

    @staticmethod
    def bounds(B, rank, world):
        """Lanes [lo, hi) of `rank`: the first B % world ranks get one extra lane."""
        base, extra = divmod(int(B), int(world))
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)

    def my_slice(self, B):
        return self.bounds(B, self.rank, self.world)

    # ---- packed per-rank result block: [theta V*cap f64][loglike cap f64][kkt 2*cap f64][iters cap i32][status cap i32]
    #      [batch size, batch digest: 2 x i64]
    @staticmethod
    def _layout(cap, V):
        o_ll = V * cap * 8
        o_kkt = o_ll + cap * 8
        o_it = o_kkt + cap * 16
        o_st = o_it + cap * 4
        o_dg = (o_st + cap * 4 + 7) // 8 * 8                     # (batch size, batch digest) of the rank that packed the block
        return o_ll, o_kkt, o_it, o_st, o_dg + 16

    def _buffers(self, cap, V, slot=0):
        key = (cap, V, slot)
        if key not in self._bufs:
            torch = self._torch
            nbytes = self._layout(cap, V)[-1]
            mine = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            every = torch.zeros(self.world * nbytes, dtype=torch.uint8, device=self.device)
            host = torch.zeros(self.world * nbytes, dtype=torch.uint8, pin_memory=True) if self.device.type == "cuda" else None
            if len(self._bufs) > 8:
                self._bufs.clear()
            self._bufs[key] = (mine, every, host)
        return self._bufs[key]

    def views(self, n, B, V, slot=0):
        """Views of this rank's packed block for its `n` lanes of a batch of B, keyed like Engine.fit's result
        (torch tensors; on the CPU their .numpy() share the memory).  `slot` names one of several buffer sets: batches that
        are in flight at the same time (LaneContext.fit_batches) must not share one."""
        torch = self._torch
        cap = (int(B) + self.world - 1) // self.world
        mine = self._buffers(cap, V, slot)[0]
        o_ll, o_kkt, o_it, o_st, _ = self._layout(cap, V)
        return {"theta": mine[0:V * n * 8].view(torch.float64).view(V, n),
                "loglike": mine[o_ll:o_ll + n * 8].view(torch.float64),
                "kkt": mine[o_kkt:o_kkt + n * 16].view(torch.float64).view(n, 2),
                "iters": mine[o_it:o_it + n * 4].view(torch.int32),
                "status": mine[o_st:o_st + n * 4].view(torch.int32)}

    def check_same_batch(self, B, digest):
        """All ranks must hold the same batch before it is split by position (a rank-dependent permutation of the
        candidates would silently pair lanes with the wrong results): one tiny all_gather, raises on mismatch."""
        torch, dist = self._torch, self._dist
        t = torch.tensor([int(B), int(digest)], dtype=torch.int64, device=self.device)
        every = torch.empty(self.world * 2, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(every, t, group=self.group)
        every = every.cpu().view(self.world, 2)
        if not bool((every == every[0]).all()):
            raise RuntimeError("LaneSharder: ranks hold different lane batches (sizes/digests %s): the candidate order must not "
                               "depend on the rank -- seed or broadcast every permutation" % every.tolist())

    def gather(self, local, B, V, slot=0, digest=None):
        """local: this rank's result dict (arrays or the `views` themselves) for its slice of a batch of B lanes;
        returns the dict (numpy) for all B lanes, identical on every rank.  `digest` (of the batch, LaneBatch.digest)
        travels in the block: a rank that fitted a different batch of the same size is caught here."""
        return self.unpack(self.gather_raw(local, B, V, slot, digest), B, V, digest is not None)

    def gather_raw(self, local, B, V, slot=0, digest=None):
        """The collective half of `gather`: pack, all_gather, copy to the host; returns the gathered bytes (a private copy).
        Threads that take turns at the collective call this inside their turn and `unpack` outside it."""
        torch, dist = self._torch, self._dist
        t_0 = time.perf_counter()
        cap = (int(B) + self.world - 1) // self.world
        mine, every, host = self._buffers(cap, V, slot)
        lo, hi = self.my_slice(B)
        n = hi - lo
        if n:
            mv = self.views(n, B, V, slot)
            for k in _RESULT_KEYS:
                src = local[k]
                if isinstance(src, torch.Tensor) and src.data_ptr() == mv[k].data_ptr():
                    continue                                   # the solver wrote straight into the packed block
                mv[k].copy_(torch.as_tensor(np.ascontiguousarray(src)).to(mv[k].dtype))
        nbytes = mine.numel()
        if digest is not None:
            mine[nbytes - 16:].view(torch.int64).copy_(torch.tensor([int(B), int(digest)], dtype=torch.int64))
        t_1 = time.perf_counter()
        dist.all_gather_into_tensor(every, mine, group=self.group)
        if host is not None:
            host.copy_(every, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            flat = host.numpy().copy()                          # the pinned buffer is reused by the next gather of this slot
        else:
            flat = every.numpy().copy()
        t_2 = time.perf_counter()
        self.bytes_gathered += int(every.numel())
        self.n_gathers += 1
        self.seconds["pack"] += t_1 - t_0
        self.seconds["all_gather_to_host"] += t_2 - t_1
        return flat

    def unpack(self, flat, B, V, check_digest=False):
        """The gathered bytes of a batch of B lanes -> the result dict (numpy) of all lanes."""
        t_2 = time.perf_counter()
        cap = (int(B) + self.world - 1) // self.world
        o_ll, o_kkt, o_it, o_st, nbytes = self._layout(cap, V)
        if check_digest:
            tails = np.stack([flat[(r + 1) * nbytes - 16:(r + 1) * nbytes].view(np.int64) for r in range(self.world)])
            if not bool((tails == tails[0]).all()):
                raise RuntimeError("LaneSharder: ranks hold different lane batches (sizes/digests %s): the candidate order must not "
                                   "depend on the rank -- seed or broadcast every permutation" % tails.tolist())
        out = {"loglike": np.empty(B, dtype=np.float64),
               "kkt": np.empty((B, 2), dtype=np.float64), "iters": np.empty(B, dtype=np.int32), "status": np.empty(B, dtype=np.int32)}
        blocks, starts = [], []
        for r in range(self.world):
            a, b = self.bounds(B, r, self.world)
            m = b - a
            if not m:
                continue
            blk = flat[r * nbytes:(r + 1) * nbytes]
            blocks.append(blk[0:V * m * 8].view(np.float64).reshape(V, m))
            starts.append(a)
            out["loglike"][a:b] = blk[o_ll:o_ll + m * 8].view(np.float64)
            out["kkt"][a:b] = blk[o_kkt:o_kkt + m * 16].view(np.float64).reshape(m, 2)
            out["iters"][a:b] = blk[o_it:o_it + m * 4].view(np.int32)
            out["status"][a:b] = blk[o_st:o_st + m * 4].view(np.int32)
        out["theta"] = ShardedTheta(blocks, starts, V, B)
        self.seconds["unpack"] += time.perf_counter() - t_2
        return out

    def broadcast_seed(self):
        """One 63-bit seed chosen by rank 0 for everything that must be random but identical on all ranks."""
        torch, dist = self._torch, self._dist
        t = torch.tensor([int(np.random.SeedSequence().entropy % (1 << 62))], dtype=torch.int64, device=self.device)
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return int(t.item())


class _InOrder(object):
    """Admits the indices 0, 1, 2, .. one at a time and in that order: the collectives of batches that are fitted
    concurrently are issued in batch order on every rank, whatever the thread timing."""

    def __init__(self):
        self.cond = threading.Condition()
        self.next = 0
        self.failed = False

    def enter(self, i):
        with self.cond:
            while self.next != i and not self.failed:
                self.cond.wait()
            if self.failed:
                raise RuntimeError("pipelined fit: another batch failed")

    def leave(self, i):
        with self.cond:
            self.next = i + 1
            self.cond.notify_all()

    def fail(self):
        with self.cond:
            self.failed = True
            self.cond.notify_all()


class _Residency(object):
    """Which weight columns of a LaneContext live in which slot of one engine's column store."""
    __slots__ = ("slot_of", "used", "cap", "last")

    def __init__(self):
        self.slot_of, self.used, self.cap, self.last = {}, 0, 0, None


class LaneContext(object):
    """The shared CSR (one `tvs_create`) + registered weight columns of every model that is fitted
    against it (the full data and each bootstrap replicate)."""

    # a batch of at least this many local lanes is cut in two halves that run concurrently on two handles (two CUDA
    # streams, two host threads): the narrow tail of one half overlaps the wide phase of the other.  Measured on cfg2:
    # 4000 lanes 40.9 -> 35.7 ms (1.15x), no gain at 1000 lanes (tools/two_handles.py).  0 disables.
    concurrent_min_lanes = 3000
    pipeline_batches = True       # fit_batches / fit_weight_batches: `pipeline_depth` batches in flight, one handle each
    pipeline_depth = None         # None: 4 for matrices of up to 8M non-zeros (latency-bound narrow phases), else 2
    selection_groups = 2          # bootstrap.run_selections: replicate groups whose host logic overlaps the other group's solver call
    # STALLED lanes (line search exhausted at fp64 resolution) are accepted only when their KKT residuals are within this
    # factor of the tolerances (complementarity tol, dual 100 tol); otherwise they are reported like MAXITER
    stalled_slack = 1e3

    def __init__(self, n_variants, variant_sizes, device=0, sharder=None, tol=1e-9, max_iter=2000, history=16):
        self.V = int(n_variants)
        self.L = np.asarray(variant_sizes, dtype=np.int64)
        self.device = int(device)
        self.sharder = sharder
        self.tol, self.max_iter, self.history = tol, max_iter, history
        self.matrix = ReadPathMatrix(self.V)
        self._cols = []               # per column: (rows, counts) sparse host copy, or ("resample", n_obs, seed, id)
        self._engine = None
        self._engine2 = None          # second handle for concurrent halves (created on first use)
        self._engines_more = []       # further handles of a deeper batch pipeline (pipeline_depth > 2)
        self._engine_R = -1
        self._resident = {}           # id(engine) -> _Residency
        self.n_lanes_fitted = 0
        self.n_batches = 0
        self.check_consistency = True
        self._checked = None
        self.time_passes = False      # CUDA events around every pass launch (roofline figures; ~1 ms per fit)
        self.collect_stats = False    # sum tvs_last_fit_stats of every solver call into `stats` (bench / diagnostics)
        self.stats = {"fits": 0, "passes": 0, "kernel_launches": 0, "lane_passes": 0, "total_ms": 0.0}
        self._stats_lock = threading.Lock()
        self.engine_factory = _cabi.Engine     # tests of the host logic inject a stand-in here
        # permutations drawn inside a sharded job must not depend on the rank (ModelFitMaxLike.py:203 shuffles the
        # candidates with the process-global np.random): one seed from rank 0
        self.shuffle_seed = sharder.broadcast_seed() if sharder is not None else None

    @classmethod
    def from_csr(cls, n_variants, variant_sizes, row_ptr, col, val, **kw):
        """A context over a matrix given as CSR arrays (rows = read paths in the caller's order)."""
        self = cls(n_variants, variant_sizes, **kw)
        self.matrix = CsrMatrix(n_variants, row_ptr, col, val)
        return self

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
        if w.size and int(w.min()) < 0:
            raise ValueError("LaneContext.add_weights: negative weight")
        rows = np.nonzero(w)[0]
        self._cols.append((rows.astype(np.int64), w[rows].astype(np.int32)))
        return len(self._cols) - 1

    def add_resampled(self, n_obs, n, seed):
        """Register n bootstrap replicates of the observed counts n_obs[R] that are DRAWN ON THE DEVICE when first used
        (tvs_columns_resample: Philox stream = column id, so a column is the same on whichever rank draws it).
        Returns the list of column ids."""
        n_obs = np.ascontiguousarray(n_obs, dtype=np.int32)
        first = len(self._cols)
        for j in range(int(n)):
            self._cols.append(("resample", n_obs, int(seed), first + j))
        return list(range(first, first + int(n)))

    @staticmethod
    def _sparse_column(w_map):
        """{row: count} -> (rows int64[], counts int32[]): columns are scattered with numpy, not Python loops."""
        rows = np.fromiter(w_map.keys(), dtype=np.int64, count=len(w_map))
        vals = np.fromiter(w_map.values(), dtype=np.int32, count=len(w_map))
        if vals.size and int(vals.min()) < 0:
            raise ValueError("LaneContext: negative weight")
        return rows, vals

    def _new_engine(self):
        row_ptr, col, val = self.matrix.arrays()
        return self.engine_factory(row_ptr, col, val, self.L, n_variants=self.V, device=self.device)

    def engine(self):
        if self._engine is None or self._engine_R != self.matrix.R:
            self.close()
            if self.matrix.R == 0:
                raise Exception("Likelihood maximization failed.")        # nothing to fit
            self._engine = self._new_engine()
            self._engine_R = self.matrix.R
        return self._engine

    def second_engine(self):
        self.engine()
        if self._engine2 is None:
            self._engine2 = self._new_engine()
        return self._engine2

    def engines(self, n):
        """The first n handles of this context (created on first use): engine(), second_engine(), then extra ones."""
        if n <= 1:
            return [self.engine()]
        out = [self.engine(), self.second_engine()][:n]
        while len(out) < n:
            k = len(out) - 2
            if k >= len(self._engines_more):
                self._engines_more.append(self._new_engine())
            out.append(self._engines_more[k])
        return out

    def close(self):
        for e in [self._engine, self._engine2] + self._engines_more:
            if e is not None:
                e.close()
        self._engine = self._engine2 = None
        self._engines_more = []
        self._resident = {}

    def dense_columns(self, cols):
        w = np.zeros((self.matrix.R, len(cols)), dtype=np.int32)
        for j, c in enumerate(cols):
            rows, vals = self._cols[c]
            w[rows, j] = vals
        return w

    def _slots(self, eng, cols):
        """Slots of `cols` (array of column ids) in the column store of `eng`; columns seen for the first time are
        uploaded (host copies, one dense block) or drawn on the device (add_resampled) now."""
        res = self._resident.setdefault(id(eng), _Residency())
        key = cols.tobytes()
        if res.last is not None and res.last[0] == key:          # the same lanes as the previous batch (a bootstrap refit)
            return res.last[1]
        uniq = np.unique(cols)
        missing = [int(c) for c in uniq if int(c) not in res.slot_of]
        if missing:
            need = res.used + len(missing)
            if need > res.cap:
                res.cap = max(need, 2 * res.cap, 16)
                eng.columns_reserve(res.cap)
            host = [c for c in missing if not (isinstance(self._cols[c][0], str))]
            if host:
                eng.columns_upload(res.used, self.dense_columns(host))
                for j, c in enumerate(host):
                    res.slot_of[c] = res.used + j
                res.used += len(host)
            drawn = [c for c in missing if isinstance(self._cols[c][0], str)]
            i = 0
            while i < len(drawn):                        # runs of consecutive ids with the same source -> one call each
                j = i
                _, n_obs, seed, cid = self._cols[drawn[i]]
                while (j + 1 < len(drawn) and self._cols[drawn[j + 1]][3] == self._cols[drawn[j]][3] + 1
                       and self._cols[drawn[j + 1]][1] is n_obs and self._cols[drawn[j + 1]][2] == seed):
                    j += 1
                n = j - i + 1
                eng.columns_resample(res.used, n, n_obs, seed, cid)
                for k in range(n):
                    res.slot_of[drawn[i + k]] = res.used + k
                res.used += n
                i = j + 1
        lut = res.slot_of
        slots = np.fromiter((lut[int(c)] for c in cols), dtype=np.int32, count=len(cols))
        res.last = (key, slots)
        return slots

    # ---- the batched fit
    def fit(self, requests):
        """requests: list of LaneRequest -> list of FitResult (same order)."""
        B = len(requests)
        if B == 0:
            return []
        out = self.fit_batch(LaneBatch.from_requests(requests, self.V))
        theta = out["theta"]
        fun = (-out["loglike"]).tolist()
        ok = self.lane_ok(out).tolist()
        status, iters, kkt = out["status"].tolist(), out["iters"].tolist(), out["kkt"].tolist()
        return [FitResult(None, fun[b], ok[b], status[b], tuple(kkt[b]), iters[b], theta=theta, ids=rq.ids, b=b)
                for b, rq in enumerate(requests)]

    def fit_items(self, items, handle=0, gate=None, turn=None):
        """items: a list whose elements are either a list of LaneRequest or a LaneBlock; ONE batch for all of them.
        Returns, per item, a list of FitResult or a BlockResult.  `turn` (a context manager) is held around the solver call
        only -- lane assembly before it and the unpacking of the results after it run outside, so several host threads
        overlap them with each other's solver calls (bootstrap.run_selections).  `handle`, `gate`: see fit_batch."""
        V = self.V
        parts, sizes = [], []
        for it in items:
            if isinstance(it, LaneBlock):
                parts.append(it.arrays(V))
            else:
                b = LaneBatch.from_requests(it, V)
                parts.append((b.cols, b.C, b.mask, b.theta0))
            sizes.append(int(parts[-1][0].shape[0]))
        B = int(sum(sizes))
        if B == 0:
            return [[] for _ in items]
        cols = np.concatenate([p[0] for p in parts])
        C = np.concatenate([p[1] for p in parts])
        mask = None
        if any(p[2] is not None for p in parts):
            mask = np.concatenate([p[2] if p[2] is not None else np.ones((V, n), dtype=np.uint8) for p, n in zip(parts, sizes)], axis=1)
        theta0 = None
        if any(p[3] is not None for p in parts):
            blocks = []
            for p, n in zip(parts, sizes):
                if p[3] is not None:
                    blocks.append(p[3])
                else:                                           # uniform over the lane's subset
                    m = p[2].astype(np.float64) if p[2] is not None else np.ones((V, n))
                    blocks.append(m / np.maximum(m.sum(axis=0), 1.0))
            theta0 = np.concatenate(blocks, axis=1)
        batch = LaneBatch(cols, C, mask, theta0)
        if turn is None:
            out = self.fit_batch(batch, handle=handle, gate=gate)
        else:
            with turn:
                out = self.fit_batch(batch, handle=handle, gate=gate)
        ok = self.lane_ok(out)
        theta = out["theta"]
        fun_all = -out["loglike"]
        results, lo = [], 0
        for it, n in zip(items, sizes):
            if isinstance(it, LaneBlock):
                results.append(BlockResult(fun_all[lo:lo + n], ok[lo:lo + n], out["status"][lo:lo + n], theta, lo, it.subsets))
            else:
                fun, okl = fun_all[lo:lo + n].tolist(), ok[lo:lo + n].tolist()
                st, itr, kkt = out["status"][lo:lo + n].tolist(), out["iters"][lo:lo + n].tolist(), out["kkt"][lo:lo + n].tolist()
                results.append([FitResult(None, fun[j], okl[j], st[j], tuple(kkt[j]), itr[j], theta=theta, ids=rq.ids, b=lo + j)
                                for j, rq in enumerate(it)])
            lo += n
        return results

    def lane_ok(self, out):
        """success flag per lane: CONVERGED, or STALLED with KKT residuals within `stalled_slack` of the tolerances (a lane
        whose line search gave up far from the optimum is a failure, like MAXITER), and a finite log-likelihood."""
        status = np.asarray(out["status"])
        kkt = np.asarray(out["kkt"])
        near = (kkt[:, 0] <= self.stalled_slack * self.tol) & (kkt[:, 1] <= self.stalled_slack * 100.0 * self.tol)
        return ((status == _cabi.CONVERGED) | ((status == _cabi.STALLED) & near)) & np.isfinite(np.asarray(out["loglike"]))

    def fit_batch(self, batch, handle=0, gate=None):
        """LaneBatch -> dict(theta [V,B], loglike [B], kkt [B,2], iters [B], status [B]) for ALL lanes of the batch (with
        a sharder: this rank fits its slice, one all_gather returns the rest).

        `handle` / `gate`: for callers that fit from several host threads (bootstrap.run_selections): the batch runs on the
        context's handle number `handle` (own CUDA stream; a thread must keep to its handle) and `gate`, a context manager,
        is held around the collective only -- the caller's gates must admit the threads in an order that is the same on
        every rank."""
        B = len(batch)
        sh = self.sharder
        digest = batch.digest() if (sh is not None and self.check_consistency) else None
        if digest is not None and gate is None and batch is not self._checked:
            sh.check_same_batch(B, digest)                      # BEFORE the fit, once per batch object (a batch fitted again is the same);
            self._checked = batch                               # threads with a gate rely on the digest inside the gathered block
        lo, hi = (0, B) if sh is None else sh.my_slice(B)
        n = hi - lo
        local = None
        views = None
        if sh is not None and gate is not None and sh.device.type == "cuda":
            sh._torch.cuda.set_device(sh.device)                # the device is a per-thread setting
        if sh is not None and n and sh.device.type == "cuda":
            views = sh.views(n, B, self.V, handle)              # the solver writes into the packed block that is gathered
        if n and handle == 0 and gate is None and self.concurrent_min_lanes and n >= self.concurrent_min_lanes:
            mid = lo + n // 2
            halves = [None, None]

            def work(i, a, b, eng):
                try:
                    halves[i] = self._fit_on(eng, batch.slice(a, b))
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
        elif n:
            local = self._fit_on(self.engines(handle + 1)[handle], batch.slice(lo, hi), out=views)
        if sh is None:
            out = local
        elif gate is None:
            out = sh.gather(local, B, self.V, handle, digest)
        else:
            with gate:
                flat = sh.gather_raw(local, B, self.V, handle, digest)
            out = sh.unpack(flat, B, self.V, digest is not None)
        with self._stats_lock:
            self.n_lanes_fitted += B
            self.n_batches += 1
        return out

    def _pipelined(self, n_batches, total_lanes, fit_local):
        """Batches 0..n-1 over `pipeline_depth` (2) handles (a CUDA stream and a host thread each): handle k fits the batches
        k, k + depth, ..; while
        one handle is in the narrow tail of its batch (few lanes left, latency-bound kernels, host polls) or copies its
        inputs in, the other is in the wide phase of the next batch.  `fit_local(i, eng, views)` fits this rank's lanes of
        batch i on `eng` and returns (local result dict, B_i); the gathers of a sharded context are issued in batch order."""
        sh = self.sharder
        depth = self.pipeline_depth
        if depth is None:
            depth = 4 if getattr(self.engine(), "nnz", 0) <= 8000000 else 2
        depth = max(2, min(int(depth), n_batches))
        engines = self.engines(depth)
        results = [None] * n_batches
        order = _InOrder()
        errors = []

        def work(k):
            try:
                if sh is not None and sh.device.type == "cuda":
                    sh._torch.cuda.set_device(sh.device)        # the device is a per-thread setting
                for i in range(k, n_batches, depth):
                    local, B = fit_local(i, engines[k], k)
                    if sh is None:
                        results[i] = local
                    else:
                        order.enter(i)
                        try:
                            flat = sh.gather_raw(local, B, self.V, slot=k)
                        finally:
                            order.leave(i)
                        results[i] = sh.unpack(flat, B, self.V)
            except BaseException as e:                          # re-raised in the calling thread
                errors.append(e)
                order.fail()
        threads = [threading.Thread(target=work, args=(k,)) for k in range(1, depth)]
        interval = sys.getswitchinterval()
        sys.setswitchinterval(min(interval, 2e-4))              # a thread that returns from the solver gets the interpreter at once
        try:
            for t in threads:
                t.start()
            work(0)
            for t in threads:
                t.join()
        finally:
            sys.setswitchinterval(interval)
        if errors:
            raise errors[0]
        self.n_lanes_fitted += total_lanes
        self.n_batches += n_batches
        return results

    def fit_batches(self, batches):
        """A sequence of LaneBatch -> the list of their result dicts (as fit_batch), with two batches in flight at any time
        (see _pipelined): for callers that hold several independent batches, e.g. a bootstrap of more replicates than one
        batch takes."""
        batches = list(batches)
        if len(batches) < 2 or not self.pipeline_batches:
            return [self.fit_batch(b) for b in batches]
        sh = self.sharder
        if sh is not None and self.check_consistency:           # one collective for the whole sequence, before any thread starts
            sh.check_same_batch(sum(len(b) for b in batches), zlib.crc32(np.array([b.digest() for b in batches], dtype=np.int64).tobytes()))

        def fit_local(i, eng, slot):
            batch = batches[i]
            B = len(batch)
            lo, hi = (0, B) if sh is None else sh.my_slice(B)
            if hi == lo:
                return None, B
            views = sh.views(hi - lo, B, self.V, slot) if (sh is not None and sh.device.type == "cuda") else None
            return self._fit_on(eng, batch.slice(lo, hi), out=views), B
        return self._pipelined(len(batches), sum(len(b) for b in batches), fit_local)

    def fit_weight_batches(self, w_locals, B=None, C=None):
        """One-shot lanes (see fit_weights) for a sequence of weight blocks, two in flight: the host->device copy of one
        block runs while the other handle fits.  `w_locals[i]` [R, n_i] are THIS rank's lanes of batch i; B (int or
        list) the batch sizes over all ranks (default n_i), C (list of arrays or None) the constants."""
        w_locals = list(w_locals)
        sh = self.sharder
        n_b = len(w_locals)
        sizes = [int(w.shape[1]) for w in w_locals]
        totals = sizes if B is None else ([int(B)] * n_b if np.isscalar(B) else [int(x) for x in B])
        if n_b < 2 or not self.pipeline_batches:
            return [self.fit_weights(w, B=t, C=(None if C is None else C[i])) for i, (w, t) in enumerate(zip(w_locals, totals))]
        if sh is not None:
            for n, t in zip(sizes, totals):
                lo, hi = sh.my_slice(t)
                if hi - lo != n:
                    raise ValueError("fit_weight_batches: this rank owns %d lanes of a batch of %d, got %d" % (hi - lo, t, n))

        def fit_local(i, eng, slot):
            n, t = sizes[i], totals[i]
            if not n:
                return None, t
            views = sh.views(n, t, self.V, slot) if (sh is not None and sh.device.type == "cuda") else None
            eng.set_lanes(w_locals[i], C=(None if C is None else C[i]), B=n)
            res = eng.fit(tol=self.tol, max_iter=self.max_iter, history=self.history, out=views)
            self._note_stats(eng)
            return res, t
        return self._pipelined(n_b, sum(totals), fit_local)

    def _fit_on(self, eng, mine, out=None):
        """Fit the lanes `mine` (LaneBatch) on one handle; returns the engine's result dict."""
        slots = self._slots(eng, np.ascontiguousarray(mine.cols))
        eng.set_lanes_columns(slots, mask=mine.mask, C=mine.C)
        res = eng.fit(theta0=mine.theta0, tol=self.tol, max_iter=self.max_iter, history=self.history, out=out,
                      time_passes=self.time_passes)
        self._note_stats(eng)
        return res

    def _note_stats(self, eng):
        if self.collect_stats:
            st = eng.last_fit_stats()
            with self._stats_lock:
                self.stats["fits"] += 1
                for k in ("passes", "kernel_launches", "lane_passes", "total_ms"):
                    self.stats[k] += st[k]

    def make_resident(self, cols):
        """Upload / draw the given columns on this rank's device now (otherwise it happens at their first fit)."""
        self._slots(self.engine(), np.asarray(cols, dtype=np.int64))

    def fit_weights(self, w_local, B=None, C=None):
        """One-shot lanes: the full-candidate-set ML fit of replicates whose weights are used once (a plain bootstrap):
        `w_local` [R, n] (uint16 or int32; numpy or pinned torch) are the weights of THIS rank's lanes of a batch of B
        (default n: single process), uploaded as lanes directly -- no column registration.  Returns the result dict of
        all B lanes (gathered when the context has a sharder)."""
        n = int(w_local.shape[1])
        sh = self.sharder
        B = n if B is None else int(B)
        if sh is not None:
            lo, hi = sh.my_slice(B)
            if hi - lo != n:
                raise ValueError("fit_weights: this rank owns %d lanes of a batch of %d, got %d" % (hi - lo, B, n))
        eng = self.engine()
        views = sh.views(n, B, self.V) if (sh is not None and n and sh.device.type == "cuda") else None
        local = None
        if n:
            eng.set_lanes(w_local, C=C, B=n)
            local = eng.fit(tol=self.tol, max_iter=self.max_iter, history=self.history, out=views)
            self._note_stats(eng)
        out = local if sh is None else sh.gather(local, B, self.V)
        self.n_lanes_fitted += B
        self.n_batches += 1
        return out

    def loglik(self, request, theta_of_ids):
        """-LL at given proportions of request.ids (single lane)."""
        theta = np.zeros((self.V, 1), dtype=np.float64)
        theta[request.ids, 0] = theta_of_ids
        mask = np.zeros((self.V, 1), dtype=np.uint8)
        mask[request.ids, 0] = 1
        eng = self.engine()
        eng.set_lanes_columns(self._slots(eng, np.array([request.col])), mask=mask, C=np.array([request.C]))
        with np.errstate(invalid="ignore"):
            return -eng.loglik(theta)
