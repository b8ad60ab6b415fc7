"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's read-path x variant occurrence counting.

Follows ``VariantSubPathsGenerator.gen_subpaths`` (traversome/utils.py:433-501) and the table construction of
``Traversome.gen_all_informative_sub_paths`` (traversome/traversome.py:1210-1263), with the path standardisers of
``Assembly`` (traversome/Assembly.py:643-653, :991-993, :1108-1133, :1174-1202) restated on a plain graph
description.  Pinned against outputs of the unmodified reference: ``tests/golden/subpaths.json.gz`` (written by
``oracle/make_golden_subpaths.py``), checked in ``tests/test_subpaths_oracle.py``.

Only ``tests/`` and the developer tools may import this module; the product (``traversome_b200/``) never does.
"""
from collections import OrderedDict


class PlainVertex(object):
    __slots__ = ("len", "connections")

    def __init__(self, length):
        self.len = int(length)
        self.connections = {True: OrderedDict(), False: OrderedDict()}


class PlainGraph(object):
    """The part of the reference's ``Assembly`` the counting reads: vertex lengths, links with overlaps, and the set
    of palindromic contigs.  Built from the JSON description the golden fixtures carry."""

    def __init__(self, vertices, links, palindromic=()):
        self.vertex_info = OrderedDict((name, PlainVertex(length)) for name, length in vertices)
        for n1, e1, n2, e2, overlap in links:               # one direction of connections[...] per entry, as dumped
            self.vertex_info[n1].connections[bool(e1)][(n2, bool(e2))] = int(overlap)
        self.palindromic_repeats = set(palindromic)

    @classmethod
    def from_json(cls, d):
        return cls(d["vertices"], d["links"], d.get("palindromic", ()))


def as_path(p):
    return tuple((str(n), bool(e)) for n, e in p)


def _fix_strands(graph, path):                               # Assembly.py:643-653
    pal = graph.palindromic_repeats
    return tuple((n, True) if n in pal else (n, e) for n, e in path)


def _revcomp(graph, path):                                   # Assembly.py:1108-1133
    pal = graph.palindromic_repeats
    return tuple((n, True) if n in pal else (n, not e) for n, e in reversed(path))


def closes(graph, path):                                     # Assembly.py:991-993
    (n_first, e_first), (n_last, e_last) = path[0], path[-1]
    return (n_first, not e_first) in graph.vertex_info[n_last].connections[e_last]


def canonical(graph, path, rotations):
    """rotations=False: Assembly.py:1174-1182; rotations=True: Assembly.py:1184-1202."""
    fwd = _fix_strands(graph, path)
    rev = _revcomp(graph, fwd)
    forms = [fwd, rev]
    if rotations and closes(graph, fwd):
        for k in range(1, len(fwd)):
            forms.append(fwd[k:] + fwd[:k])
            forms.append(rev[k:] + rev[:k])
    return min(forms)


def count_read_paths_in_variant(graph, variant_path, read_paths, max_alignment_len):
    """utils.py:433-501 for one variant: OrderedDict read path -> occurrences, in first-hit order."""
    found = OrderedDict()
    n = len(variant_path)
    wraps = closes(graph, variant_path)
    limit = max_alignment_len - 2
    for start in range(n):
        window = [variant_path[start]]
        grown = 0
        nxt = start + 1
        while grown <= limit:
            if nxt >= n:
                if not wraps:
                    break
                nxt -= n
            name, strand = variant_path[nxt]
            vertex = graph.vertex_info[name]
            grown += vertex.len - vertex.connections[not strand][window[-1]]
            window.append((name, strand))
            nxt += 1
        for size in range(len(window), 0, -1):
            # windows of circular variants use the plain standardiser, those of linear variants the rotating one
            # (utils.py:461 versus :490)
            key = canonical(graph, window[:size], rotations=not wraps)
            if key in read_paths:
                found[key] = found.get(key, 0) + 1
    return found


def informative_sub_paths(graph, variant_paths, variant_sizes, read_paths, max_alignment_len):
    """traversome.py:1210-1263: (counters per variant, be_unidentifiable_to, repr_to_merged_variants,
    all_sub_paths as OrderedDict path -> {"inner_len", "full_len", "from_variants"})."""
    read_paths = set(read_paths)
    counters = [count_read_paths_in_variant(graph, tuple(p), read_paths, max_alignment_len) for p in variant_paths]
    be = OrderedDict()
    V = len(variant_paths)
    for rep in range(V):
        for other in range(rep, V):
            if other in be:
                continue
            if other == rep or (dict(counters[other]) == dict(counters[rep]) and
                                variant_sizes[other] == variant_sizes[rep]):   # plain dicts: order-insensitive
                be[other] = rep
    merged = OrderedDict((r, []) for r in sorted(set(be.values())))
    for v, r in be.items():
        merged[r].append(v)
    table = OrderedDict()
    for v, counter in enumerate(counters):
        for path, c in counter.items():
            if path not in table:
                table[path] = {"inner_len": inner_length(graph, path), "full_len": full_length(graph, path),
                               "from_variants": OrderedDict()}
            table[path]["from_variants"][v] = c
    return counters, be, merged, table


def full_length(graph, path):                                # Assembly.py:1003-1020, adjust_for_cyclic=False
    vi = graph.vertex_info
    total = vi[path[-1][0]].len
    for (n1, e1), (n2, e2) in zip(path[:-1], path[1:]):
        total += vi[n1].len - vi[n1].connections[e1][(n2, not e2)]
    return total


def inner_length(graph, path):                               # Assembly.py:1023-1047, keep_terminal_overlaps=True
    if len(path) < 3:
        return 0
    vi = graph.vertex_info
    total = sum(vi[n].len for n, _ in path[1:-1])
    for (n1, e1), (n2, e2) in zip(path[1:-2], path[2:-1]):
        total -= vi[n1].connections[e1][(n2, not e2)]
    return max(total, 0)
