"""Read-path x variant occurrence counting on the GPU.

Host mirror of ``VariantSubPathsGenerator`` (traversome/utils.py:416-501) and of the counting part of
``Traversome.gen_all_informative_sub_paths`` (traversome/traversome.py:1210-1263).  The reference walks, in Python,
every start position of every variant path, grows the longest window the alignment length allows, standardises each
prefix and counts those that are observed read paths.  Here the graph is turned into integer tables once, all variants
go to ``tvs_subpaths_count`` (include/tvs_subpaths.h) in one call, and the integer matrix A[read path, variant] comes
back as CSR together with the rank of each first hit, which reproduces the reference's dict orders (and therefore the
row order of ``all_sub_paths``).  Integer work: results are identical to the reference's, not approximately equal.

``graph`` is duck-typed on what the reference's ``Assembly`` offers: ``vertex_info[name].len``,
``vertex_info[name].connections[end] -> {(name2, end2): overlap}`` and ``palindromic_repeats`` (a set, possibly empty;
call ``graph.detect_palindromic_repeats()`` first on a reference graph, Assembly.py:613-641).
"""
from collections import OrderedDict
from itertools import chain

import numpy as np

from . import _cabi
from .utils import SubPathInfo


# ---------------------------------------------------------------------------------------------- graph tables
class GraphTables(object):
    """Integer view of the assembly graph.  token = 2 * contig id + strand (True = 1)."""

    def __init__(self, graph):
        self.graph = graph
        pal = getattr(graph, "palindromic_repeats", None)
        if pal is None and hasattr(graph, "detect_palindromic_repeats"):
            graph.detect_palindromic_repeats()          # Assembly.py:613-641 (prunes the graph: must precede the edge table)
            pal = graph.palindromic_repeats
        self.palindromic = set(pal or ())
        self.names = list(graph.vertex_info)
        self.id_of = {n: i for i, n in enumerate(self.names)}
        self._tok_of = {}
        for i, n in enumerate(self.names):
            self._tok_of[(n, True)] = 2 * i + 1
            self._tok_of[(n, False)] = 2 * i
        n = len(self.names)
        self.length = np.fromiter((int(graph.vertex_info[v].len) for v in self.names), dtype=np.int64, count=n)
        self.is_pal = np.zeros(n, dtype=bool)
        for v in self.palindromic:
            if v in self.id_of:
                self.is_pal[self.id_of[v]] = True
        src, dst, ov = [], [], []
        for name in self.names:
            conn = graph.vertex_info[name].connections
            i = self.id_of[name]
            for end in (True, False):
                for (n2, e2), overlap in conn[end].items():
                    src.append(2 * i + int(end))
                    dst.append(2 * self.id_of[n2] + int(not e2))
                    ov.append(int(overlap))
        self._stride = 2 * n + 1
        key = np.asarray(src, dtype=np.int64) * self._stride + np.asarray(dst, dtype=np.int64)
        order = np.argsort(key, kind="stable")
        self._edge_key = key[order]
        self._edge_ov = np.asarray(ov, dtype=np.int64)[order]

    # path = sequence of (name, strand)
    def tokens(self, paths):
        """(ptr int64[n+1], raw tokens int32[total]) of a list of paths."""
        ptr = np.zeros(len(paths) + 1, dtype=np.int64)
        np.cumsum([len(p) for p in paths], out=ptr[1:])
        # one C-level pass: (name, strand) tuples are looked up whole, no per-token Python bytecode
        tok = np.fromiter(map(self._tok_of.__getitem__, chain.from_iterable(paths)), dtype=np.int32, count=int(ptr[-1]))
        return ptr, tok

    def corrected(self, tok):
        """Assembly.correct_path_with_palindromic_repeats (Assembly.py:643-653) on tokens."""
        if not self.palindromic:
            return tok
        return np.where(self.is_pal[tok >> 1], tok | 1, tok).astype(np.int32)

    def edge_overlap(self, tok_a, tok_b):
        """(found bool[], overlap int64[]) of the links tok_a -> tok_b."""
        key = tok_a.astype(np.int64) * self._stride + tok_b.astype(np.int64)
        if self._edge_key.size == 0:
            return np.zeros(key.shape, dtype=bool), np.zeros(key.shape, dtype=np.int64)
        pos = np.minimum(np.searchsorted(self._edge_key, key), self._edge_key.size - 1)
        found = self._edge_key[pos] == key
        return found, np.where(found, self._edge_ov[pos], 0)

    # ---- the reference's path functions, on (name, strand) tuples
    def correct(self, path):
        if not self.palindromic:
            return tuple(path)
        pal = self.palindromic
        return tuple((v, True) if v in pal else (v, e) for v, e in path)

    def reverse(self, path):
        """Assembly.reverse_path (Assembly.py:1108-1133)."""
        pal = self.palindromic
        if pal:
            return tuple((v, True) if v in pal else (v, not e) for v, e in path[::-1])
        return tuple((v, not e) for v, e in path[::-1])

    def is_circular(self, path):
        """Assembly.is_circular_path (Assembly.py:991-993)."""
        (n0, e0), (n1, e1) = path[0], path[-1]
        return (n0, not e0) in self.graph.vertex_info[n1].connections[e1]

    def standardized(self, path):
        """Assembly.get_standardized_path (Assembly.py:1174-1182)."""
        fwd = self.correct(path)
        rev = self.reverse(fwd)
        return fwd if fwd <= rev else rev

    def standardized_circ(self, path):
        """Assembly.get_standardized_path_circ (Assembly.py:1184-1202)."""
        fwd = self.correct(path)
        rev = self.reverse(fwd)
        if not self.is_circular(fwd):
            return fwd if fwd <= rev else rev
        best = min(fwd, rev)
        for k in range(1, len(fwd)):
            best = min(best, fwd[k:] + fwd[:k], rev[k:] + rev[:k])
        return best

    def may_close(self, path):
        """True when some window of a variant could be a proper rotation of `path`: its last contig links to its first
        one, under either strand of a palindromic end."""
        pal = self.palindromic
        (n0, e0), (n1, e1) = path[0], path[-1]
        conn = self.graph.vertex_info[n1].connections
        for a in ((e1, not e1) if n1 in pal else (e1,)):
            for b in ((e0, not e0) if n0 in pal else (e0,)):
                if (n0, not b) in conn[a]:
                    return True
        return False

    def path_lengths(self, path):
        """(inner_len, full_len) of a read path: Assembly.get_path_internal_length (Assembly.py:1023-1047) and
        get_path_length(check_valid=False, adjust_for_cyclic=False) (Assembly.py:1003-1020)."""
        vi = self.graph.vertex_info
        full = 0
        for (n1, e1), (n2, e2) in zip(path[:-1], path[1:]):
            full += vi[n1].len - vi[n1].connections[e1][(n2, not e2)]
        full += vi[path[-1][0]].len
        if len(path) < 3:
            return 0, full
        inner = sum(vi[n].len for n, _ in path[1:-1])
        inner -= sum(vi[n1].connections[e1][(n2, not e2)] for (n1, e1), (n2, e2) in zip(path[1:-2], path[2:-1]))
        return max(inner, 0), full


# ---------------------------------------------------------------------------------------------- encoders
def encode_variants(tables, variant_paths):
    """(var_ptr, var_tok (strand-corrected), var_step, var_circular) of the variant paths (utils.py:441-486)."""
    ptr, raw = tables.tokens(variant_paths)
    n_var = len(variant_paths)
    if n_var == 0:
        return ptr, raw, raw.copy(), np.zeros(0, dtype=np.uint8)
    first, last = ptr[:-1], ptr[1:] - 1
    if np.any(last < first):
        raise ValueError("empty variant path")
    prev_index = np.arange(raw.size, dtype=np.int64) - 1
    prev_index[first] = last                                     # position 0 is entered from the last contig
    found, overlap = tables.edge_overlap(raw[prev_index], raw)
    circular = found[first]                                      # Assembly.is_circular_path on the raw variant path
    is_first = np.zeros(raw.size, dtype=bool)
    is_first[first] = True
    needed = ~is_first | np.repeat(circular, np.diff(ptr))
    if np.any(needed & ~found):
        bad = int(np.flatnonzero(needed & ~found)[0])
        v = int(np.searchsorted(ptr, bad, side="right") - 1)
        raise KeyError("variant %d steps over a link that is not in the graph (position %d)" % (v, bad - int(ptr[v])))
    step = np.where(needed, tables.length[raw >> 1] - overlap, 0)
    if np.any(np.abs(step) > 2 ** 31 - 1):
        raise OverflowError("contig length beyond int32")
    return ptr, tables.corrected(raw), step.astype(np.int32), circular.astype(np.uint8)


def _encode_read_paths_general(tables, read_paths, rows=None):
    """Keys of the read paths (`rows` = their row numbers, default 0..n-1), one path at a time with the standardisers
    themselves: (key_ptr, key_tok, key_row_circular, key_row_linear), see include/tvs_subpaths.h.

    Windows of circular variants are standardised by get_standardized_path (utils.py:461): a window W stands for read
    path r iff r is min(W, rc W), i.e. W is r or rc(r) and r is standard.  Windows of linear variants go through
    get_standardized_path_circ (utils.py:490): the same unless W is itself a circular path, in which case all rotations
    of both strands compete; every candidate sequence is pushed through that standardiser and kept iff it lands on r.
    The two rules can send one sequence to two different read paths, hence two rows per key."""
    circ_row, lin_row = {}, {}

    def put(table, seq, row):
        if table.setdefault(seq, row) != row:
            raise ValueError("two read paths claim the same window: rows %d and %d" % (table[seq], row))

    for row, r in (enumerate(read_paths) if rows is None else zip(rows, read_paths)):
        r = tuple(r)
        if not r or tables.correct(r) != r:
            continue                                        # a window is strand-corrected first: r can never be hit
        rev = tables.reverse(r)
        if r <= rev:                                        # r is standard
            put(circ_row, r, row)
            put(circ_row, rev, row)
        cands = {r, rev}
        if len(r) > 1 and (tables.may_close(r) or tables.may_close(rev)):
            for k in range(1, len(r)):
                cands.add(r[k:] + r[:k])
                cands.add(rev[k:] + rev[:k])
        for cand in cands:
            if tables.standardized_circ(cand) == r:
                put(lin_row, cand, row)
    seqs = list(OrderedDict.fromkeys(list(circ_row) + list(lin_row)))
    key_ptr, raw = tables.tokens(seqs)
    key_row_circ = np.fromiter((circ_row.get(s, -1) for s in seqs), dtype=np.int32, count=len(seqs))
    key_row_lin = np.fromiter((lin_row.get(s, -1) for s in seqs), dtype=np.int32, count=len(seqs))
    return key_ptr, raw, key_row_circ, key_row_lin


def encode_read_paths(tables, read_paths):
    """Keys of all read paths.  Without palindromic contigs a read path that is not itself a closed walk has the keys
    {r, rc(r)} under both standardisers, provided r <= rc(r) in the reference's ordering of (name, strand) tuples; that
    common case is done on the token arrays.  Closed walks (and everything on a graph with palindromic contigs) go
    through `_encode_read_paths_general`."""
    read_paths = [tuple(p) for p in read_paths]
    if tables.palindromic or not read_paths:
        return _encode_read_paths_general(tables, read_paths)
    ptr, tok = tables.tokens(read_paths)
    n = len(read_paths)
    lens = np.diff(ptr)
    if np.any(lens == 0):
        raise ValueError("empty read path")
    seg = np.repeat(np.arange(n, dtype=np.int64), lens)
    idx = np.arange(tok.size, dtype=np.int64)
    mirror = ptr[seg] + ptr[seg + 1] - 1 - idx
    rev = tok[mirror] ^ 1
    # order of (name, strand) tuples: by name as a string, then False < True
    rank = np.empty(len(tables.names), dtype=np.int64)
    rank[np.argsort(np.asarray(tables.names, dtype=object), kind="stable")] = np.arange(len(tables.names))
    ord_fwd = 2 * rank[tok >> 1] + (tok & 1)
    ord_rev = 2 * rank[rev >> 1] + (rev & 1)
    differ = ord_fwd != ord_rev
    first_diff = np.minimum.reduceat(np.where(differ, idx, tok.size), ptr[:-1])
    self_rc = first_diff == tok.size
    at = np.minimum(first_diff, tok.size - 1)
    standard = self_rc | (ord_fwd[at] < ord_rev[at])
    closes, _ = tables.edge_overlap(tok[ptr[1:] - 1], tok[ptr[:-1]])
    slow = closes & (lens > 1)
    fast_rows = np.flatnonzero(standard & ~slow)
    both = fast_rows[~self_rc[fast_rows]]

    def gather(rows_, source):
        ln = lens[rows_]
        out_ptr = np.zeros(rows_.size + 1, dtype=np.int64)
        np.cumsum(ln, out=out_ptr[1:])
        take = np.repeat(ptr[rows_] - out_ptr[:-1], ln) + np.arange(int(out_ptr[-1]), dtype=np.int64)
        return out_ptr, source[take]

    parts = [gather(fast_rows, tok) + (fast_rows, fast_rows), gather(both, rev) + (both, both)]
    slow_rows = np.flatnonzero(slow)
    if slow_rows.size:
        parts.append(_encode_read_paths_general(tables, [read_paths[i] for i in slow_rows.tolist()], slow_rows.tolist()))
    key_ptr = np.zeros(1, dtype=np.int64)
    for part_ptr, _, _, _ in parts:
        key_ptr = np.concatenate([key_ptr, part_ptr[1:] + key_ptr[-1]])
    return (key_ptr, np.concatenate([p[1] for p in parts]).astype(np.int32),
            np.concatenate([p[2] for p in parts]).astype(np.int32), np.concatenate([p[3] for p in parts]).astype(np.int32))


# ---------------------------------------------------------------------------------------------- result
class SubPathCounts(object):
    """A[read path, variant] as CSR over the rows handed to the generator, plus first-hit ranks."""

    def __init__(self, read_paths, n_variants, row_ptr, col, val, first, stats):
        self.read_paths = read_paths
        self.n_variants = int(n_variants)
        self.row_ptr, self.col, self.val, self.first = row_ptr, col, val, first
        self.stats = stats

    def dense(self):
        a = np.zeros((len(self.read_paths), self.n_variants), dtype=np.int64)
        rows = np.repeat(np.arange(len(self.read_paths)), np.diff(self.row_ptr))
        a[rows, self.col] = self.val
        return a

    def counters(self):
        """Per variant, the reference's ``these_sub_paths`` dict in its insertion order (utils.py:438-500)."""
        rows = np.repeat(np.arange(len(self.read_paths), dtype=np.int64), np.diff(self.row_ptr))
        order = np.lexsort((self.first, self.col))
        out = [dict() for _ in range(self.n_variants)]
        rp = self.read_paths
        for v, r, c in zip(self.col[order].tolist(), rows[order].tolist(), self.val[order].tolist()):
            out[v][rp[r]] = c
        return out

    def row_order(self):
        """Rows in ``all_sub_paths`` order: first variant containing the read path, then that variant's dict order
        (traversome.py:1246-1263).  Read paths no variant contains are absent."""
        has = np.flatnonzero(np.diff(self.row_ptr) > 0)
        k0 = self.row_ptr[has]                                # columns increase inside a row: k0 is the first variant
        return has[np.lexsort((self.first[k0], self.col[k0]))]

    def all_sub_paths(self, tables):
        """OrderedDict read path -> SubPathInfo(inner_len, full_len, from_variants) like traversome.py:1246-1263."""
        out = OrderedDict()
        for r in self.row_order().tolist():
            info = SubPathInfo()
            info.inner_len, info.full_len = tables.path_lengths(self.read_paths[r])
            lo, hi = int(self.row_ptr[r]), int(self.row_ptr[r + 1])
            info.from_variants = dict(zip(self.col[lo:hi].tolist(), self.val[lo:hi].tolist()))
            out[self.read_paths[r]] = info
        return out

    def csr_in_reference_order(self):
        """(row_ptr int32, col, val, rows) of A with rows in ``all_sub_paths`` order: the arrays tvs_create takes."""
        rows = self.row_order()
        counts = (self.row_ptr[rows + 1] - self.row_ptr[rows]).astype(np.int64)
        row_ptr = np.zeros(rows.size + 1, dtype=np.int64)
        np.cumsum(counts, out=row_ptr[1:])
        take = np.repeat(self.row_ptr[rows] - row_ptr[:-1], counts) + np.arange(int(row_ptr[-1]), dtype=np.int64)
        return row_ptr.astype(np.int32), self.col[take], self.val[take], rows

    def unidentifiable(self, variant_sizes):
        """(be_unidentifiable_to, repr_to_merged_variants) of traversome.py:1222-1243: a variant is represented by the
        first variant with the same counter and the same size."""
        V = self.n_variants
        rows = np.repeat(np.arange(len(self.read_paths), dtype=np.int64), np.diff(self.row_ptr))
        order = np.lexsort((rows, self.col))
        col_s, row_s, val_s = self.col[order], rows[order], self.val[order]
        cptr = np.zeros(V + 1, dtype=np.int64)
        np.cumsum(np.bincount(col_s, minlength=V), out=cptr[1:])
        groups = {}
        rep = {}
        for v in range(V):
            lo, hi = cptr[v], cptr[v + 1]
            sig = (int(variant_sizes[v]), row_s[lo:hi].tobytes(), val_s[lo:hi].tobytes())
            rep[v] = groups.setdefault(sig, v)
        members = OrderedDict()
        for v in range(V):
            members.setdefault(rep[v], []).append(v)
        be = OrderedDict()
        for r in range(V):                                      # the reference's insertion order
            if r in be:
                continue
            for v in members[r]:
                be[v] = r
        merged = OrderedDict((r, []) for r in sorted(set(be.values())))
        for v, r in be.items():
            merged[r].append(v)
        return be, merged


# ---------------------------------------------------------------------------------------------- generator
class VariantSubPathsGenerator(object):
    """Same constructor and ``gen_subpaths`` as the reference class (utils.py:416-501); ``count`` is the batched form.

    ``read_paths_hashed`` may be any collection of standardised read paths; rows of the result follow its iteration
    order (pass a list / the keys of ``Traversome.read_paths`` for a reproducible order)."""

    def __init__(self, graph, min_alignment_len, max_alignment_len, read_paths_hashed, device=0):
        self.graph = graph
        self.min_alignment_len = min_alignment_len
        self.max_alignment_len = max_alignment_len
        self.read_paths_hashed = read_paths_hashed
        self.variant_subpath_counters = {}
        self.device = device
        self._tables = None
        self._keys = None
        self._rows = None

    @property
    def tables(self):
        if self._tables is None:
            self._tables = GraphTables(self.graph)
        return self._tables

    def _prepare(self):
        if self._keys is None:
            self._rows = [tuple(p) for p in self.read_paths_hashed]
            if len(set(self._rows)) != len(self._rows):
                raise ValueError("read_paths_hashed holds the same read path twice")
            self._keys = encode_read_paths(self.tables, self._rows)
        return self._rows, self._keys

    def count(self, variant_paths):
        """All variants in one device call -> SubPathCounts."""
        rows, (key_ptr, key_tok, key_row_circ, key_row_lin) = self._prepare()
        variant_paths = [tuple(p) for p in variant_paths]
        var_ptr, var_tok, var_step, var_circ = encode_variants(self.tables, variant_paths)
        row_ptr, col, val, first, stats = _cabi.count_subpaths(
            var_ptr, var_tok, var_step, var_circ, len(rows), key_ptr, key_tok, key_row_circ, key_row_lin,
            int(self.max_alignment_len) - 2, device=self.device)
        return SubPathCounts(rows, len(variant_paths), row_ptr, col, val, first, stats)

    def gen_subpaths(self, variant_path):
        variant_path = tuple(variant_path)
        hit = self.variant_subpath_counters.get(variant_path)
        if hit is None:
            hit = self.variant_subpath_counters[variant_path] = self.count([variant_path]).counters()[0]
        return hit

    def gen_all(self, variant_paths):
        """Fill ``variant_subpath_counters`` for every variant (the loop of traversome.py:1217-1219) in one call."""
        counts = self.count(variant_paths)
        for path, counter in zip(variant_paths, counts.counters()):
            self.variant_subpath_counters.setdefault(tuple(path), counter)
        return counts


def gen_all_informative_sub_paths(generator, variant_paths, variant_sizes):
    """Array form of Traversome.gen_all_informative_sub_paths (traversome.py:1210-1263).

    Returns (variant_subpath_counters, be_unidentifiable_to, repr_to_merged_variants, all_sub_paths, counts)."""
    counts = generator.gen_all(variant_paths)
    counters = OrderedDict((tuple(p), generator.variant_subpath_counters[tuple(p)]) for p in variant_paths)
    be, merged = counts.unidentifiable(variant_sizes)
    return counters, be, merged, counts.all_sub_paths(generator.tables), counts
