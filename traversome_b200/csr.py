"""Integer CSR build of the read-path x variant count matrix.

Replaces the symbolic expression build of ``PathMultinomialModel.get_like_formula``
(ModelGenerator.py:153-174) + ``ModelFitMaxLike.get_neg_likelihood_of_var_freq``
(ModelFitMaxLike.py:515-558).  Everything here is integer bookkeeping on the host and is
bit-exact against the reference's tables:

  A[r, v]  = ``from_variants[v]`` of read path r                 (traversome.py:1250-1264)
  w[r]     = sum of ``num_matched`` over the bins of read path r  (traversome.py:1354-1362),
             bins with ``num_possible_X < 1`` pruned first        (ModelGenerator.py:138-151)
  L[v]     = ``variant_sizes[v]``
  C        = sum_bins num_matched * log(num_possible_X)           (the theta-free part of the LL)
  N        = sum_bins num_matched                                 (``sample_size``)

Bins of one read path share one ``from_variants`` dict (traversome.py:1364), so bins merge into
rows by the CONTENT of that dict; two read paths with identical content have identical
probabilities for every theta and merge too (exact, it only shortens the matrix).
"""
import math

import numpy as np


def _signature(from_variants):
    return tuple(sorted((int(v), int(c)) for v, c in from_variants.items()))


class ReadPathMatrix(object):
    """Row registry: distinct ``from_variants`` signatures in first-seen order -> CSR arrays."""

    def __init__(self, n_variants):
        self.V = int(n_variants)
        self._row_of = {}
        self._sigs = []
        self._by_id = {}     # id(from_variants dict) -> row, a cache for the shared dict objects

    @property
    def R(self):
        return len(self._sigs)

    def row_index(self, from_variants, add=True):
        key = id(from_variants)
        hit = self._by_id.get(key)
        if hit is not None and hit[1] is from_variants:
            return hit[0]
        sig = _signature(from_variants)
        row = self._row_of.get(sig)
        if row is None:
            if not add:
                raise KeyError("read path signature %r not in the matrix" % (sig,))
            if not sig:
                raise ValueError("a bin without any variant cannot be modelled (empty from_variants)")
            if sig[0][0] < 0 or sig[-1][0] >= self.V:
                raise ValueError("variant id out of range in from_variants: %r" % (sig,))
            row = len(self._sigs)
            self._row_of[sig] = row
            self._sigs.append(sig)
        self._by_id[key] = (row, from_variants)
        return row

    def row_index_distinct(self, from_variants):
        """Append a row for this read path even when another read path has the same from_variants content (array-form
        replicates keep one row per read path of ``all_sub_paths``, in that order)."""
        sig = _signature(from_variants)
        if not sig:
            raise ValueError("a read path without any variant cannot be modelled (empty from_variants)")
        if sig[0][0] < 0 or sig[-1][0] >= self.V:
            raise ValueError("variant id out of range in from_variants: %r" % (sig,))
        self._sigs.append(sig)
        self._row_of.setdefault(sig, len(self._sigs) - 1)
        return len(self._sigs) - 1

    def add_sub_paths(self, all_sub_paths):
        """Register every read path of ``model.all_sub_paths`` (rows then follow that order)."""
        for info in all_sub_paths.values():
            if info.from_variants:
                self.row_index(info.from_variants)

    def arrays(self):
        """(row_ptr[R+1], col[nnz], val[nnz]) int32; columns strictly increasing inside a row."""
        counts = np.fromiter((len(s) for s in self._sigs), dtype=np.int64, count=len(self._sigs))
        row_ptr = np.zeros(len(self._sigs) + 1, dtype=np.int64)
        np.cumsum(counts, out=row_ptr[1:])
        nnz = int(row_ptr[-1])
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz, dtype=np.int32)
        k = 0
        for sig in self._sigs:
            for v, c in sig:
                col[k] = v
                val[k] = c
                k += 1
        return row_ptr.astype(np.int32), col, val

    def dense(self):
        a = np.zeros((self.R, self.V), dtype=np.int64)
        for r, sig in enumerate(self._sigs):
            for v, c in sig:
                a[r, v] = c
        return a

    def covers(self, variant_ids, observed_rows):
        """True iff every observed row has >=1 variant of ``variant_ids`` (row A8: the matrix form of
        ModelFitMaxLike.cover_all_observed_subpaths, ModelFitMaxLike.py:568-580)."""
        ids = set(variant_ids)
        for r in observed_rows:
            if not any(v in ids for v, _ in self._sigs[r]):
                return False
        return True


class CsrMatrix(object):
    """A read-path x variant matrix given directly as CSR arrays (synthetic workloads, array-form replicates): the
    same `R` / `arrays()` surface as ReadPathMatrix without one Python tuple per non-zero."""

    def __init__(self, n_variants, row_ptr, col, val):
        self.V = int(n_variants)
        self._row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self._col = np.ascontiguousarray(col, dtype=np.int32)
        self._val = np.ascontiguousarray(val, dtype=np.int32)

    @property
    def R(self):
        return int(self._row_ptr.shape[0] - 1)

    def arrays(self):
        return self._row_ptr, self._col, self._val


def lane_from_bins(bins_list, matrix, add_rows=True):
    """One lane (= one model / replicate) -> (w {row: count}, C, N).

    Applies the reference's pruning rule (num_possible_X < 1 dropped, ModelGenerator.py:144) WITHOUT
    mutating ``bins_list``; iterates bins in the reference's order."""
    w = {}
    C = 0.0
    N = 0
    for bins in bins_list:
        for rp_bin in bins.rp_bins:
            X = rp_bin.num_possible_X
            if X < 1:
                continue
            n = int(rp_bin.num_matched)
            row = matrix.row_index(rp_bin.from_variants, add=add_rows)
            w[row] = w.get(row, 0) + n
            C += math.log(X) * n
            N += n
    return w, C, N


def lanes_from_models(models, matrix=None, n_variants=None):
    """Batch of models sharing one variant set -> (matrix, w[R,B] int32, C[B], N[B] int64)."""
    if matrix is None:
        matrix = ReadPathMatrix(n_variants if n_variants is not None else models[0].num_put_variants)
    per_lane = [lane_from_bins(m.bins_list, matrix) for m in models]
    R, B = matrix.R, len(models)
    w = np.zeros((R, B), dtype=np.int32)
    C = np.zeros(B, dtype=np.float64)
    N = np.zeros(B, dtype=np.int64)
    for b, (wb, Cb, Nb) in enumerate(per_lane):
        for r, n in wb.items():
            w[r, b] = n
        C[b] = Cb
        N[b] = Nb
    return matrix, w, C, N


def subset_masks(n_variants, subsets):
    """[V,B] uint8 masks from an iterable of variant-id collections (None = all variants)."""
    subsets = list(subsets)
    mask = np.zeros((n_variants, len(subsets)), dtype=np.uint8)
    for b, ids in enumerate(subsets):
        if ids is None:
            mask[:, b] = 1
        else:
            mask[list(ids), b] = 1
    return mask
