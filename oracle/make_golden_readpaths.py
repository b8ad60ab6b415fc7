#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/readpaths.json.gz.

Runs the UNMODIFIED ``Traversome.generate_read_paths`` (traversome/traversome.py:869-966) of the reference on seeded
random graphs and synthetic alignment records (objects with ``.path`` and ``.p_align_len``) and records its outputs:
``read_paths`` (ordered, with record ids), ``align_len_at_path_map``, ``num_valid_records``, min / max alignment length,
``max_read_path_size``.  Run HERE, once (needs /root/reference); the output is committed.
"""
import gzip
import json
import os
import random
import sys
import tempfile
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = "/root/reference"


class Rec(object):
    def __init__(self, path, p_align_len):
        self.path, self.p_align_len = tuple(path), int(p_align_len)


class Ali(object):
    def __init__(self, raw_records):
        self.raw_records = raw_records


def main():
    sys.path.insert(0, HERE)
    sys.path.insert(0, REFERENCE)
    import symengine_standin
    symengine_standin.install()
    import traversome.traversome as tvm
    from traversome.Assembly import Assembly
    import make_golden_subpaths as mg

    specs = [
        ("circular graph", dict(seed=31, n_contigs=14, n_variants=6, var_len=(10, 24), circular=True, max_alignment_len=500,
                                n_read_paths=0, max_window=6), 600),
        ("linear graph with repeat units", dict(seed=32, n_contigs=8, n_variants=6, var_len=(8, 14), circular=False,
                                                max_alignment_len=500, n_read_paths=0, max_window=6, repeat_units=True), 500),
        ("palindromic contig", dict(seed=33, n_contigs=10, n_variants=6, var_len=(8, 16), circular=True, max_alignment_len=500,
                                    n_read_paths=0, max_window=5, palindromic=True), 500),
    ]
    cases = []
    for label, kw, n_records in specs:
        rnd, names, seqs, variants, links, _mal, _nrp, max_window = mg.random_case(**kw)
        tmp = tempfile.mkdtemp(prefix="tvs_rp_")
        gfa = os.path.join(tmp, "g.gfa")
        mg.write_gfa(gfa, seqs, links)
        graph = Assembly(gfa)
        graph.detect_palindromic_repeats()
        variants = [v for v in variants if graph.contain_path(v)]
        records = []
        for go in range(n_records):
            v = variants[rnd.randrange(len(variants))]
            n = len(v)
            # skewed window choice so that some paths are deep and many are shallow
            s = int(n * rnd.random() ** 2) % n
            k = rnd.randint(1, max_window)
            w = [v[(s + i) % n] for i in range(k)] if kw["circular"] else list(v[s:s + k])
            if rnd.random() < 0.5:                                       # aligners report either strand
                w = [(nm, not e) for nm, e in reversed(w)]
            u = rnd.random()
            if u < 0.04:
                w = w + [(rnd.choice(names), rnd.random() < 0.5)]       # probably a link the graph does not have
            elif u < 0.06:
                w = [("ghost%d" % rnd.randint(0, 2), True)] + w          # a contig the graph does not have
            elif u < 0.07:
                w = []                                                   # empty path
            records.append(Rec(w, rnd.randint(500, 9000)))
        runs = []
        for filter_by_graph, mac in ((True, 1), (True, 2), (True, 4), (False, 1), (False, 3)):
            recs = records if filter_by_graph else [r for r in records if r.path and all(nm in graph.vertex_info for nm, _ in r.path)]
            trv = object.__new__(tvm.Traversome)
            trv.graph = graph
            trv.read_paths = OrderedDict()
            trv.read_paths_masked = set()
            trv.generate_read_paths(Ali(recs), filter_by_graph=filter_by_graph, min_alignment_counts=mac)
            runs.append({
                "filter_by_graph": filter_by_graph, "min_alignment_counts": mac,
                "record_subset": None if filter_by_graph else [i for i, r in enumerate(records)
                                                               if r.path and all(nm in graph.vertex_info for nm, _ in r.path)],
                "read_paths": [[[[nm, bool(e)] for nm, e in p], [int(x) for x in ids]] for p, ids in trv.read_paths.items()],
                "align_len_at_path_map": sorted([int(k), int(v)] for k, v in trv.align_len_at_path_map.items()),
                "num_valid_records": int(trv.num_valid_records),
                "min_alignment_length": int(trv.min_alignment_length), "max_alignment_length": int(trv.max_alignment_length),
                "max_read_path_size": int(trv.max_read_path_size)})
            print("%-32s filter %-5s min_counts %d: %4d records -> %3d read paths, %4d valid records" % (
                label, filter_by_graph, mac, len(recs), len(trv.read_paths), trv.num_valid_records))
        dumped_links = []
        for name, info in graph.vertex_info.items():
            for end in (True, False):
                for (n2, e2), ov in info.connections[end].items():
                    dumped_links.append([name, bool(end), n2, bool(e2), int(ov)])
        cases.append({"label": label,
                      "graph": {"vertices": [[n, int(graph.vertex_info[n].len)] for n in graph.vertex_info],
                                "links": dumped_links, "palindromic": sorted(graph.palindromic_repeats or ())},
                      "records": [[[[nm, bool(e)] for nm, e in r.path], r.p_align_len] for r in records],
                      "runs": runs})
    out = {"what": "read-path construction from alignment records (SURVEY.md 8f item 4, traversome.py:869-966)",
           "reference_version": "0.1.7.1", "generator": "oracle/make_golden_readpaths.py", "cases": cases}
    path = os.path.join(GOLDEN, "readpaths.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as h:
        h.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
