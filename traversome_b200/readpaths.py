"""Read-path construction from alignment records in array form (the part after parsing).

Mirror of ``Traversome.generate_read_paths`` (traversome/traversome.py:869-966): every alignment record's path is
standardised (``Assembly.get_standardized_path``), records are grouped by path in order of first appearance, paths the
graph does not contain are dropped, and with ``min_alignment_counts > 1`` a path supported by fewer records survives only
if, counted together with its occurrences as a window of LONGER read paths, it reaches the threshold (:892-927).

The reference does all of it one record (and one window) at a time in Python.  Here the record paths are tokenised in
one pass and standardised, grouped and counted with array operations (hashes are only a lookup device: every match is
confirmed token by token).  Host-side integer work: the result equals the reference's, including the order of
``read_paths``.  (A first version counted the window support with ``tvs_subpaths_count``; its dense [paths x candidates]
scratch made that slower than this at 10^5 read paths -- tools/readpaths_bench.py.)

GAF parsing and the per-read filters of ``GraphAlignRecords`` (GraphAlignRecords.py:151-171, :936-1025) stay with the
reference: `raw_records` is any sequence of objects with ``.path`` and ``.p_align_len``.
"""
from collections import OrderedDict
from itertools import chain

import numpy as np

from .subpaths import GraphTables


class ReadPathTable(object):
    """What generate_read_paths leaves on the Traversome object (traversome.py:869-966)."""

    def __init__(self):
        self.read_paths = OrderedDict()          # standardised path -> record ids (ascending)
        self.align_len_at_path_map = {}          # record id -> p_align_len
        self.num_valid_records = 0
        self.min_alignment_length = None
        self.max_alignment_length = None
        self.max_read_path_size = None


_HASH_ADD = np.uint64(0x9E3779B97F4A7C15)
_HASH_MUL = 0x100000001B3


def _window_support(flat_ptr, tok, lens, cand, is_pal):
    """support[c] = number of windows of STRICTLY longer kept paths that standardise onto candidate path cand[c]
    (traversome.py:905-921).  A window W of size k stands for candidate r iff W == r or rc(W) == r (r is standard).
    Windows and candidates of one size are compared through a 64-bit polynomial hash; every hash match is confirmed
    token by token, and two candidates of one size with the same hash (never seen) are told apart exactly."""
    support = np.zeros(cand.size, dtype=np.int64)
    tok = tok.astype(np.int64)
    with np.errstate(over="ignore"):
        for k in np.unique(lens[cand]).tolist():
            c_k = cand[lens[cand] == k]
            hosts = np.flatnonzero(lens > k)
            if hosts.size == 0:
                continue
            n_win = lens[hosts] - k + 1
            w_ptr = np.zeros(hosts.size + 1, dtype=np.int64)
            np.cumsum(n_win, out=w_ptr[1:])
            start = np.repeat(flat_ptr[hosts] - w_ptr[:-1], n_win) + np.arange(int(w_ptr[-1]), dtype=np.int64)
            pw = np.ones(k, dtype=np.uint64)
            for j in range(1, k):
                pw[j] = pw[j - 1] * np.uint64(_HASH_MUL)
            h_fwd = np.zeros(start.size, dtype=np.uint64)
            h_rev = np.zeros(start.size, dtype=np.uint64)
            self_rc = np.ones(start.size, dtype=bool)
            for j in range(k):
                f = tok[start + j]
                r = tok[start + k - 1 - j] ^ 1
                r = np.where(is_pal[r >> 1], r | 1, r)
                h_fwd += (f.astype(np.uint64) + _HASH_ADD) * pw[j]
                h_rev += (r.astype(np.uint64) + _HASH_ADD) * pw[j]
                self_rc &= f == r
            h_c = np.zeros(c_k.size, dtype=np.uint64)
            for j in range(k):
                h_c += (tok[flat_ptr[c_k] + j].astype(np.uint64) + _HASH_ADD) * pw[j]
            order = np.argsort(h_c, kind="stable")
            h_sorted = h_c[order]
            if np.any(h_sorted[1:] == h_sorted[:-1]):
                # two candidates of one size with the same 64-bit hash: count this size exactly, window by window
                index = {tuple(tok[flat_ptr[c]:flat_ptr[c] + k].tolist()): int(np.searchsorted(cand, c)) for c in c_k.tolist()}
                for s0 in start.tolist():
                    f = tuple(tok[s0:s0 + k].tolist())
                    r = tok[s0:s0 + k][::-1] ^ 1
                    r = tuple(np.where(is_pal[r >> 1], r | 1, r).tolist())
                    hit = index.get(f, index.get(r))
                    if hit is not None:
                        support[hit] += 1
                continue
            for h_w, mask, reverse in ((h_fwd, None, False), (h_rev, ~self_rc, True)):
                pos = np.minimum(np.searchsorted(h_sorted, h_w), h_sorted.size - 1)
                hit = h_sorted[pos] == h_w
                if mask is not None:
                    hit &= mask                                    # a window equal to its own reverse complement counts once
                w_idx = np.flatnonzero(hit)
                if w_idx.size == 0:
                    continue
                c_idx = order[pos[w_idx]]
                ok = np.ones(w_idx.size, dtype=bool)
                for j in range(k):                                 # confirm the match token by token
                    wt = tok[start[w_idx] + (k - 1 - j if reverse else j)]
                    if reverse:
                        wt = wt ^ 1
                        wt = np.where(is_pal[wt >> 1], wt | 1, wt)
                    ok &= wt == tok[flat_ptr[c_k[c_idx]] + j]
                support += np.bincount(np.searchsorted(cand, c_k[c_idx[ok]]), minlength=cand.size)
    return support


def _hash_segments(tok, ptr, mult):
    """Wrap-around polynomial hash of every segment tok[ptr[i]:ptr[i+1]] (uint64)."""
    lens = np.diff(ptr)
    pos = np.arange(tok.size, dtype=np.int64) - np.repeat(ptr[:-1], lens)
    pw = np.ones(int(lens.max()) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(1, pw.size):
            pw[i] = pw[i - 1] * np.uint64(mult)
        terms = (tok.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) * pw[pos]
        return np.add.reduceat(terms, ptr[:-1])


def generate_read_paths(graph, raw_records, filter_by_graph=True, min_alignment_counts=1, tables=None):
    """Array form of Traversome.generate_read_paths.  Raises SystemExit(0) like the reference when nothing remains.
    With `filter_by_graph=False` an empty alignment path is rejected (the reference would file it under the empty tuple)."""
    tables = tables if tables is not None else GraphTables(graph)
    out = ReadPathTable()
    n_rec = len(raw_records)
    paths = [rec.path for rec in raw_records]
    lens_all = np.fromiter((len(p) for p in paths), dtype=np.int64, count=n_rec)
    rec_ids = np.flatnonzero(lens_all > 0)                    # "Record i is empty" (:877-879); an empty path is also falsy
    if rec_ids.size != n_rec:
        if not filter_by_graph:
            raise ValueError("empty alignment path")
        paths = [paths[i] for i in rec_ids.tolist()]
    lens = lens_all[rec_ids]
    ptr = np.zeros(rec_ids.size + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    n = rec_ids.size
    if n == 0:
        raise SystemExit(0)
    # contigs the graph does not have still take part in the standardisation and the grouping (by name, like the
    # reference's tuples); they get ids beyond the graph's and fail the containment test below
    names = list(tables.names)
    n_graph = len(names)
    tok_of = dict(tables._tok_of)

    def token(x):
        t = tok_of.get(x)
        if t is None:
            name = x[0]
            if (name, True) not in tok_of:
                tok_of[(name, True)], tok_of[(name, False)] = 2 * len(names) + 1, 2 * len(names)
                names.append(name)
            t = tok_of[(name, bool(x[1]))]
        return t
    try:                                                      # C-level lookups when every contig is in the graph
        raw = np.fromiter(map(tables._tok_of.__getitem__, chain.from_iterable(paths)), dtype=np.int64, count=int(ptr[-1]))
    except KeyError:
        raw = np.fromiter(map(token, chain.from_iterable(paths)), dtype=np.int64, count=int(ptr[-1]))
    is_pal = np.concatenate([tables.is_pal, np.zeros(len(names) - n_graph, dtype=bool)])
    known = raw < 2 * n_graph
    seg = np.repeat(np.arange(n, dtype=np.int64), lens)
    path_known = np.logical_and.reduceat(known, ptr[:-1])
    # ---- standardise: strand correction of palindromic contigs, reverse complement, smaller of the two (:876 -> Assembly.py:1174-1182)
    fwd = np.where(is_pal[raw >> 1], raw | 1, raw)
    idx = np.arange(raw.size, dtype=np.int64)
    mirror = ptr[seg] + ptr[seg + 1] - 1 - idx
    rev = fwd[mirror] ^ 1
    rev = np.where(is_pal[rev >> 1], rev | 1, rev)
    rank = np.empty(len(names), dtype=np.int64)
    rank[np.argsort(np.asarray(names, dtype=object), kind="stable")] = np.arange(len(names))
    ord_fwd = 2 * rank[fwd >> 1] + (fwd & 1)
    ord_rev = 2 * rank[rev >> 1] + (rev & 1)
    first_diff = np.minimum.reduceat(np.where(ord_fwd != ord_rev, idx, raw.size), ptr[:-1])
    at = np.minimum(first_diff, raw.size - 1)
    use_rev = (first_diff < raw.size) & (ord_rev[at] < ord_fwd[at])
    std = np.where(np.repeat(use_rev, lens), rev, fwd)
    # ---- group equal paths (64-bit hash + length as the sort key, then verified token by token)
    with np.errstate(over="ignore"):
        key = _hash_segments(std, ptr, _HASH_MUL) * np.uint64(31) + lens.astype(np.uint64)
    _, first_rec, inverse = np.unique(key, return_index=True, return_inverse=True)
    rep = first_rec[inverse]                                   # first record of each record's group
    same = std == std[np.repeat(ptr[rep], lens) + (idx - np.repeat(ptr[:-1], lens))]
    if not same.all():                                         # a hash collision: group exactly instead
        seen = {}
        rep = np.fromiter((seen.setdefault(tuple(std[ptr[i]:ptr[i + 1]].tolist()), i) for i in range(n)), dtype=np.int64, count=n)
    groups = np.unique(rep)                                    # ascending first record = order of first appearance
    # ---- filter 1: the graph must contain the path (:880-882 -> Assembly.contain_path, Assembly.py:1135-1150)
    g_ptr0, g_len = ptr[groups], lens[groups]
    keep = np.ones(groups.size, dtype=bool)
    if filter_by_graph:
        keep &= path_known[groups]
        inner = np.ones(raw.size, dtype=bool)
        inner[ptr[:-1]] = False                                # positions that are entered over a link
        prev = std[np.maximum(idx - 1, 0)]
        both_known = (prev < 2 * n_graph) & (std < 2 * n_graph)
        found, _ = tables.edge_overlap(np.where(both_known, prev, 0), np.where(both_known, std, 0))
        link_ok = np.logical_and.reduceat((found & both_known) | ~inner, ptr[:-1])
        keep &= link_ok[groups]
    groups, g_ptr0, g_len = groups[keep], g_ptr0[keep], g_len[keep]
    group_of_rep = -np.ones(n, dtype=np.int64)
    group_of_rep[groups] = np.arange(groups.size)
    rec_group = group_of_rep[rep]                              # -1: dropped
    order = np.argsort(rec_group, kind="stable")
    order = order[rec_group[order] >= 0]
    counts = np.bincount(rec_group[rec_group >= 0], minlength=groups.size)
    g_rec_ptr = np.zeros(groups.size + 1, dtype=np.int64)
    np.cumsum(counts, out=g_rec_ptr[1:])
    # ---- filter 2: shallow paths need support from longer read paths (:892-927)
    alive = np.ones(groups.size, dtype=bool)
    if min_alignment_counts > 1 and groups.size:
        shallow = np.flatnonzero(counts < min_alignment_counts)
        if shallow.size:
            flat_ptr = np.zeros(groups.size + 1, dtype=np.int64)
            np.cumsum(g_len, out=flat_ptr[1:])
            take = np.repeat(g_ptr0 - flat_ptr[:-1], g_len) + np.arange(int(flat_ptr[-1]), dtype=np.int64)
            support = _window_support(flat_ptr, std[take], g_len, shallow, is_pal)
            alive[shallow[counts[shallow] + support < min_alignment_counts]] = False
    if not alive.any():
        raise SystemExit(0)                                    # "No valid alignment records remains after filtering!"
    # ---- outputs in the reference's shapes
    align_len = np.fromiter((rec.p_align_len for rec in raw_records), dtype=np.int64, count=n_rec)
    tuple_of = [(names[t >> 1], bool(t & 1)) for t in range(2 * len(names))]
    std_list = std.tolist()
    alive_groups = np.flatnonzero(alive)
    rec_sorted = rec_ids[order]                                # record ids grouped by path, ascending inside a group
    starts, ends = g_rec_ptr[alive_groups].tolist(), g_rec_ptr[alive_groups + 1].tolist()
    rec_list = rec_sorted.tolist()
    p0s, lns = g_ptr0[alive_groups].tolist(), g_len[alive_groups].tolist()
    for p0, ln, a0, a1 in zip(p0s, lns, starts, ends):
        out.read_paths[tuple(map(tuple_of.__getitem__, std_list[p0:p0 + ln]))] = rec_list[a0:a1]
    kept = rec_sorted[np.repeat(alive, counts)]              # rec_sorted runs through the groups in order
    out.align_len_at_path_map = dict(zip(kept.tolist(), align_len[kept].tolist()))
    out.num_valid_records = len(out.align_len_at_path_map)
    lengths = sorted(out.align_len_at_path_map.values())
    out.min_alignment_length, out.max_alignment_length = lengths[0], lengths[-1]
    out.max_read_path_size = max(len(p) for p in out.read_paths)
    return out


def install_on(traversome_obj, table):
    """Copy a ReadPathTable onto a Traversome object: the attributes generate_read_paths sets (traversome.py:869-966)."""
    traversome_obj.read_paths = table.read_paths
    traversome_obj.align_len_at_path_map = table.align_len_at_path_map
    traversome_obj.num_valid_records = table.num_valid_records
    traversome_obj.min_alignment_length = table.min_alignment_length
    traversome_obj.max_alignment_length = table.max_alignment_length
    traversome_obj.max_read_path_size = table.max_read_path_size
