"""Shared helpers of the test-suite: golden tables -> objects of this package."""
from collections import OrderedDict

import numpy as np

from traversome_b200.ModelGenerator import PathMultinomialModel
from traversome_b200.utils import BinInfo, Bins, SubPathInfo


def hexf(x):
    return float.fromhex(x) if isinstance(x, str) else float(x)


def model_from_tables(variant_sizes, bins_tables, n_rows=None):
    """Golden JSON tables -> PathMultinomialModel of THIS package (bins of one read path share one
    from_variants dict, like traversome.py:1364)."""
    shared = {}
    all_sub_paths = OrderedDict()
    rec = [0]

    def fv_of(b):
        key = b.get("row", tuple(map(tuple, b["from_variants"])))
        if key not in shared:
            shared[key] = {int(v): int(c) for v, c in b["from_variants"]}
            spi = SubPathInfo()
            spi.from_variants = shared[key]
            spi.inner_len, spi.full_len = 0, 10 ** 6
            all_sub_paths[(("rp%s" % (len(all_sub_paths)), True),)] = spi
        return shared[key]
    bins_list = []
    for bins in bins_tables:
        bo = Bins(min_len=bins["min_len"], max_len=bins["max_len"], min_id=bins["min_id"], max_id=bins["max_id"])
        for b in bins["rp_bins"]:
            fv = fv_of(b)
            bo.rp_bins.append(BinInfo(from_variants=fv, num_possible_X=hexf(b["num_possible_X"]),
                                      num_matched=int(b["num_matched"])))
        bins_list.append(bo)
    # records: give every read path as many record ids as it has matched reads
    for bins in bins_list:
        for rp in bins.rp_bins:
            for spi in all_sub_paths.values():
                if spi.from_variants is rp.from_variants:
                    spi.mapped_records = spi.mapped_records + list(range(rec[0], rec[0] + rp.num_matched))
                    rec[0] += rp.num_matched
                    break
    return PathMultinomialModel(variant_sizes=list(variant_sizes), variant_topos=[True] * len(variant_sizes),
                                bins_list=bins_list, all_sub_paths=all_sub_paths)


def fit_context(model, be_unidentifiable_to=None, repr_to_merged_variants=None):
    """The remaining constructor arguments of ModelFitMaxLike (ModelFitMaxLike.py:66-74) for a model
    built by model_from_tables."""
    V = model.num_put_variants
    variant_paths = [(("var%d" % v, True),) for v in range(V)]
    counters = OrderedDict()
    for v in range(V):
        counters[variant_paths[v]] = {p: spi.from_variants[v] for p, spi in model.all_sub_paths.items()
                                      if v in spi.from_variants}
    if be_unidentifiable_to is None:
        be_unidentifiable_to = OrderedDict((v, v) for v in range(V))
        repr_to_merged_variants = OrderedDict((v, [v]) for v in range(V))
    else:
        be_unidentifiable_to = OrderedDict((int(k), int(v)) for k, v in be_unidentifiable_to)
        repr_to_merged_variants = OrderedDict((int(k), [int(x) for x in v]) for k, v in repr_to_merged_variants)
    return dict(model=model, variant_paths=variant_paths, variant_subpath_counters=counters,
                sbp_to_sbp_id={p: i for i, p in enumerate(model.all_sub_paths)},
                repr_to_merged_variants=repr_to_merged_variants, be_unidentifiable_to=be_unidentifiable_to)


def philox4x32_10(c, k):
    """numpy Philox4x32-10 (Salmon et al. 2011); c [n,4] uint32 counters, k [2] uint32 key -> [n,4] uint32."""
    c = np.array(c, dtype=np.uint64)
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def resample_reference(n_obs, B, seed, keep_lane0):
    """numpy restatement of tvs_resample's documented stream (include/tvs.h)."""
    n_obs = np.asarray(n_obs, dtype=np.int64)
    R, N = n_obs.size, int(n_obs.sum())
    cum = np.cumsum(n_obs)
    w = np.zeros((R, B), dtype=np.int32)
    i = np.arange(N, dtype=np.uint64)
    for b in range(B):
        if keep_lane0 and b == 0:
            w[:, 0] = n_obs
            continue
        ctr = np.stack([i & np.uint64(0xFFFFFFFF), i >> np.uint64(32), np.full(N, b, np.uint64),
                        np.full(N, 0x54565321, np.uint64)], axis=1)
        o = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
        u = (o[:, 1].astype(np.uint64) << np.uint64(32)) | o[:, 0].astype(np.uint64)
        idx = np.array([(int(x) * N) >> 64 for x in u], dtype=np.int64)
        rows = np.searchsorted(cum, idx, side="right")
        w[:, b] = np.bincount(rows, minlength=R)
    return w


class OracleEngine(object):
    """CPU stand-in for traversome_b200._cabi.Engine, backed by the oracle's exact solver.  TEST ONLY:
    lets the host-side logic above the C ABI (lane batching, lock-step selection, sharding + gather)
    run in `-m "not gpu"` tests; injected through LaneContext.engine_factory."""

    def __init__(self, row_ptr, col, val, variant_sizes, n_variants=None, device=0):
        import scipy.sparse as sp
        self.R = len(row_ptr) - 1
        self.V = int(n_variants if n_variants is not None else len(variant_sizes))
        self.A = sp.csr_matrix((np.asarray(val, np.float64), np.asarray(col), np.asarray(row_ptr)), shape=(self.R, self.V))
        self.L = np.asarray(variant_sizes, np.float64)
        self.calls = 0

    def set_lanes_indexed(self, w_cols, col_of_lane, mask=None, C=None):
        self.w = np.asarray(w_cols)[:, np.asarray(col_of_lane)]
        self.mask = np.ones((self.V, self.w.shape[1]), np.uint8) if mask is None else np.asarray(mask)
        self.C = np.zeros(self.w.shape[1]) if C is None else np.asarray(C)
        self.B = self.w.shape[1]

    def set_lanes(self, w, mask=None, C=None, B=None):
        w = np.asarray(w)
        self.set_lanes_indexed(w, np.arange(w.shape[1]), mask=mask, C=C)

    # device-resident column store of the real engine (tvs_columns_*), host arrays here
    def columns_reserve(self, n_cols):
        old = getattr(self, "cols", None)
        self.cols = np.zeros((self.R, int(n_cols)), dtype=np.int64)
        if old is not None:
            self.cols[:, :old.shape[1]] = old

    def columns_upload(self, first, w_cols):
        w_cols = np.asarray(w_cols)
        self.cols[:, first:first + w_cols.shape[1]] = w_cols

    def set_lanes_columns(self, col_of_lane, mask=None, C=None):
        self.set_lanes_indexed(self.cols, col_of_lane, mask=mask, C=C)

    def fit(self, tol=0.0, max_iter=0, history=0, **_kw):
        from oracle import exact
        self.calls += 1
        B = self.B
        out = {"theta": np.zeros((self.V, B)), "loglike": np.zeros(B), "kkt": np.zeros((B, 2)),
               "iters": np.zeros(B, np.int32), "status": np.zeros(B, np.int32)}
        for b in range(B):
            try:
                r = exact.exact_fit(self.A, self.L, self.w[:, b], mask=self.mask[:, b].astype(bool))
            except ValueError:
                out["status"][b] = 3
                out["loglike"][b] = -np.inf
                out["theta"][:, b] = np.nan
                continue
            out["theta"][:, b] = r["theta"]
            out["loglike"][b] = r["loglike_no_C"] + self.C[b]
            out["kkt"][b] = (r["comp"], r["dual"])
        return out

    def close(self):
        pass


def model_from_workload(wl, n_bins=3, shared_from_variants=None):
    """Reference-style model objects (Bins/BinInfo/SubPathInfo/PathMultinomialModel) from a synthetic Workload:
    every read path gets one BinInfo (X = wl.X[r], n = wl.n_obs[r]) in one of `n_bins` length ranges.
    `shared_from_variants`: list that receives / supplies the per-read-path from_variants dicts, so that replicate models
    share those objects with the whole-data model exactly like Traversome.sample_sub_paths does (traversome.py:1671)."""
    all_sub_paths = OrderedDict()
    bins_list = [Bins(min_len=1000 * (i + 1), max_len=1000 * (i + 2) - 1, min_id=0, max_id=0) for i in range(n_bins)]
    rec = 0
    for r in range(wl.R):
        if shared_from_variants is not None and len(shared_from_variants) > r:
            fv = shared_from_variants[r]
        else:
            fv = {int(c): int(v) for c, v in zip(wl.col[wl.row_ptr[r]:wl.row_ptr[r + 1]], wl.val[wl.row_ptr[r]:wl.row_ptr[r + 1]])}
            if shared_from_variants is not None:
                shared_from_variants.append(fv)
        spi = SubPathInfo()
        spi.from_variants = fv
        spi.inner_len, spi.full_len = 0, 10 ** 6
        n = int(wl.n_obs[r])
        spi.mapped_records = list(range(rec, rec + n))
        rec += n
        all_sub_paths[(("rp%d" % r, True),)] = spi
        bins_list[r % n_bins].rp_bins.append(BinInfo(from_variants=fv, num_possible_X=float(wl.X[r]), num_matched=n))
    return PathMultinomialModel(variant_sizes=[int(x) for x in wl.L], variant_topos=[True] * wl.V, bins_list=bins_list,
                                all_sub_paths=all_sub_paths)
