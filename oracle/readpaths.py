"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``Traversome.generate_read_paths`` (traversome/traversome.py:869-966)
on the plain graph description of ``oracle/subpaths.py``.  Pinned against outputs of the unmodified reference:
``tests/golden/readpaths.json.gz`` (``oracle/make_golden_readpaths.py``), checked in ``tests/test_readpaths.py``.
Only ``tests/`` and developer tools may import this module.
"""
from collections import OrderedDict

from .subpaths import canonical


def graph_has_path(graph, path):                             # Assembly.contain_path, Assembly.py:1135-1150
    previous = None
    for name, strand in path:
        if name not in graph.vertex_info:
            return False
        if previous is not None and (name, not strand) not in graph.vertex_info[previous[0]].connections[previous[1]]:
            return False
        previous = (name, strand)
    return True


def read_paths_from_records(graph, record_paths, record_lengths, filter_by_graph=True, min_alignment_counts=1):
    """-> (OrderedDict path -> record ids, {record id: alignment length})."""
    table = OrderedDict()
    rejected = set()
    for rid, raw in enumerate(record_paths):
        path = canonical(graph, tuple(raw), rotations=False) if raw else ()
        if filter_by_graph:
            if not path or path in rejected:
                continue
            if path not in table and not graph_has_path(graph, path):
                rejected.add(path)
                continue
        table.setdefault(path, []).append(rid)
    if min_alignment_counts > 1:
        support = {p: len(ids) for p, ids in table.items() if len(ids) < min_alignment_counts}
        for longer in table:
            for size in range(1, len(longer)):
                for start in range(len(longer) - size + 1):
                    window = canonical(graph, longer[start:start + size], rotations=False)
                    if window in support:
                        support[window] += 1
        for path, total in support.items():
            if total < min_alignment_counts:
                del table[path]
    lengths = {rid: record_lengths[rid] for ids in table.values() for rid in ids}
    return table, lengths
