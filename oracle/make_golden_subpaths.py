#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/subpaths.json.gz.

Runs the UNMODIFIED reference (`/root/reference/traversome`, v0.1.7.1) in this container on seeded random assembly
graphs and records what its read-path x variant counting produces:

  VariantSubPathsGenerator.gen_subpaths            traversome/utils.py:433-501      (per-variant counters, dict order)
  Traversome.gen_all_informative_sub_paths         traversome/traversome.py:1210-1296
      -> be_unidentifiable_to, repr_to_merged_variants, all_sub_paths (order, inner_len, full_len, from_variants)

Nothing here passes through the symengine stand-in (it is installed only so that `traversome.traversome` imports).
/root/reference does not exist on the GPU box: run HERE, once; the output is committed.

Usage: python oracle/make_golden_subpaths.py
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
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def write_gfa(path, seqs, links):
    with open(path, "w") as h:
        h.write("H\tVN:Z:1.0\n")
        for name, seq in seqs.items():
            h.write("S\t%s\t%s\n" % (name, seq))
        for (n1, e1, n2, e2), ov in links.items():
            h.write("L\t%s\t%s\t%s\t%s\t%dM\n" % (n1, "+-"[not e1], n2, "+-"[not e2], ov))


def random_case(seed, n_contigs, n_variants, var_len, circular, max_alignment_len, n_read_paths, max_window,
                palindromic=False, repeat_units=False, duplicates=False, overlaps=(0, 3, 7)):
    rnd = random.Random(seed)
    names = ["c%02d" % i for i in range(n_contigs)]
    seqs = OrderedDict((n, "".join(rnd.choice("ACGT") for _ in range(rnd.randint(30, 260)))) for n in names)
    pal_name = None
    if palindromic:
        half = "".join(rnd.choice("ACGT") for _ in range(40))
        pal_name = names[1]
        seqs[pal_name] = half + revcomp(half)
    variants = []
    for _ in range(n_variants):
        n = rnd.randint(*var_len)
        if repeat_units and rnd.random() < 0.7:
            unit = [(rnd.choice(names), rnd.random() < 0.7) for _ in range(rnd.randint(2, 3))]
            v = [(rnd.choice(names), rnd.random() < 0.7) for _ in range(rnd.randint(1, 3))]
            v += unit * rnd.randint(2, 4)
            v += [(rnd.choice(names), rnd.random() < 0.7) for _ in range(rnd.randint(1, 3))]
        else:
            v = [(rnd.choice(names), rnd.random() < 0.7) for _ in range(n)]
        variants.append(v)
    if duplicates:
        variants.append(list(variants[0]))
        variants.append(variants[1][1:] + variants[1][:1])              # a rotation of a circular variant
    links = OrderedDict()

    def add_link(a, b):
        key = (a[0], a[1], b[0], b[1])
        mirror = (b[0], not b[1], a[0], not a[1])
        if key not in links and mirror not in links:
            links[key] = rnd.choice(overlaps)

    for v in variants:
        for a, b in zip(v[:-1], v[1:]):
            add_link(a, b)
        if circular:
            add_link(v[-1], v[0])
    if palindromic:
        # every link of the palindromic contig exists on both of its ends with the same overlap, so that its two
        # connection tables are equal: with more than two neighbours it is "complicated, recorded" as a palindromic
        # repeat (Assembly.py:627-639) and strand-corrected paths stay valid paths of the graph
        for other in names[2:5]:
            add_link((pal_name, True), (other, True))
        add_link((pal_name, True), (pal_name, False))       # hairpin: the order-insensitive trigger of Assembly.py:629
        def put(key, ov):
            mirror = (key[2], not key[3], key[0], not key[1])
            links[mirror if mirror in links else key] = ov

        for (n1, e1, n2, e2), ov in list(links.items()):
            if n1 == pal_name and n2 == pal_name:
                for a in (True, False):
                    for b in (True, False):
                        put((n1, a, n2, b), 0)
            elif n1 == pal_name:
                put((n1, not e1, n2, e2), links[(n1, e1, n2, e2)])
            elif n2 == pal_name:
                put((n1, e1, n2, not e2), links[(n1, e1, n2, e2)])
    return rnd, names, seqs, variants, links, max_alignment_len, n_read_paths, max_window


def run_case(label, rnd, names, seqs, variants, links, max_alignment_len, n_read_paths, max_window, circular):
    from traversome.Assembly import Assembly
    from traversome.utils import VariantSubPathsGenerator
    import traversome.traversome as tvm

    tmp = tempfile.mkdtemp(prefix="tvs_sp_")
    gfa = os.path.join(tmp, "g.gfa")
    write_gfa(gfa, seqs, links)
    graph = Assembly(gfa)
    graph.detect_palindromic_repeats()
    # variants whose steps survived the palindromic pruning only
    variant_paths = [tuple(v) for v in variants if graph.contain_path(v)]
    read_paths = OrderedDict()
    for v in variant_paths:
        n = len(v)
        for _ in range(n_read_paths):
            s, k = rnd.randrange(n), rnd.randint(1, max_window)
            if graph.is_circular_path(v):
                w = [v[(s + i) % n] for i in range(k)]
            else:
                w = list(v[s:s + k])
            read_paths[graph.get_standardized_path(w)] = None
    for _ in range(n_read_paths):                                     # read paths no variant has
        w = [(rnd.choice(names), rnd.random() < 0.5) for _ in range(rnd.randint(1, 3))]
        read_paths[graph.get_standardized_path(w)] = None
    read_paths = list(read_paths)
    rnd.shuffle(read_paths)

    trv = object.__new__(tvm.Traversome)
    trv.graph = graph
    trv.variant_paths = variant_paths
    trv.num_put_variants = len(variant_paths)
    trv.variant_sizes = [graph.get_path_length(v, check_valid=False, adjust_for_cyclic=True) for v in variant_paths]
    trv.variant_subpath_counters = OrderedDict()
    trv.read_paths = OrderedDict((rp, [0]) for rp in read_paths)
    trv.all_sub_paths = OrderedDict()
    trv.subpath_generator = VariantSubPathsGenerator(graph=graph, min_alignment_len=10,
                                                     max_alignment_len=max_alignment_len,
                                                     read_paths_hashed=set(read_paths))
    trv.gen_all_informative_sub_paths()

    def jp(path):
        return [[n, bool(e)] for n, e in path]

    dumped_links = []
    for name, info in graph.vertex_info.items():
        for end in (True, False):
            for (n2, e2), ov in info.connections[end].items():
                dumped_links.append([name, bool(end), n2, bool(e2), int(ov)])
    return {
        "label": label,
        "graph": {"vertices": [[n, int(graph.vertex_info[n].len)] for n in graph.vertex_info],
                  "links": dumped_links, "palindromic": sorted(graph.palindromic_repeats or ())},
        "max_alignment_len": max_alignment_len,
        "variant_paths": [jp(v) for v in variant_paths],
        "variant_circular": [bool(graph.is_circular_path(v)) for v in variant_paths],
        "variant_sizes": [int(x) for x in trv.variant_sizes],
        "read_paths": [jp(r) for r in read_paths],
        "expected": {
            "counters": [[[jp(p), int(c)] for p, c in trv.variant_subpath_counters[v].items()] for v in variant_paths],
            "be_unidentifiable_to": [[int(k), int(v)] for k, v in trv.be_unidentifiable_to.items()],
            "repr_to_merged_variants": [[int(k), [int(x) for x in v]] for k, v in trv.repr_to_merged_variants.items()],
            "all_sub_paths": [{"path": jp(p), "inner_len": int(i.inner_len), "full_len": int(i.full_len),
                               "from_variants": [[int(k), int(c)] for k, c in i.from_variants.items()]}
                              for p, i in trv.all_sub_paths.items()],
        },
    }


def main():
    sys.path.insert(0, HERE)
    sys.path.insert(0, REFERENCE)
    import symengine_standin
    symengine_standin.install()
    import traversome  # noqa: F401

    specs = [
        ("circular, mixed overlaps", dict(seed=11, n_contigs=14, n_variants=6, var_len=(10, 24), circular=True,
                                          max_alignment_len=500, n_read_paths=40, max_window=5), True),
        ("circular, windows longer than the variant", dict(seed=12, n_contigs=6, n_variants=4, var_len=(3, 6),
                                                           circular=True, max_alignment_len=1500, n_read_paths=30,
                                                           max_window=9), True),
        ("linear", dict(seed=13, n_contigs=12, n_variants=6, var_len=(8, 20), circular=False, max_alignment_len=450,
                        n_read_paths=40, max_window=5), False),
        ("linear with repeat units (circular windows)", dict(seed=14, n_contigs=8, n_variants=6, var_len=(8, 14),
                                                             circular=False, max_alignment_len=900, n_read_paths=60,
                                                             max_window=6, repeat_units=True), False),
        ("circular with a palindromic contig", dict(seed=15, n_contigs=10, n_variants=6, var_len=(8, 16), circular=True,
                                                    max_alignment_len=500, n_read_paths=40, max_window=5,
                                                    palindromic=True), True),
        ("linear with a palindromic contig and repeat units", dict(seed=16, n_contigs=8, n_variants=6, var_len=(8, 14),
                                                                   circular=False, max_alignment_len=700,
                                                                   n_read_paths=50, max_window=6, palindromic=True,
                                                                   repeat_units=True), False),
        ("circular with duplicated variants", dict(seed=17, n_contigs=10, n_variants=4, var_len=(8, 14), circular=True,
                                                   max_alignment_len=400, n_read_paths=30, max_window=4,
                                                   duplicates=True), True),
        ("mixed: linear variants on a graph that also closes them", dict(seed=18, n_contigs=9, n_variants=8,
                                                                          var_len=(5, 12), circular=True,
                                                                          max_alignment_len=600, n_read_paths=40,
                                                                          max_window=6, repeat_units=True), True),
    ]
    cases = []
    for label, kw, circular in specs:
        rnd, names, seqs, variants, links, mal, nrp, mw = random_case(**kw)
        if label.startswith("mixed"):
            # drop the closing link of half of the variants' ends by appending a private tail contig
            for i, v in enumerate(variants):
                if i % 2:
                    tail = "t%02d" % i
                    seqs[tail] = "".join(rnd.choice("ACGT") for _ in range(50))
                    names.append(tail)
                    links[(v[-1][0], v[-1][1], tail, True)] = 0
                    v.append((tail, True))
        case = run_case(label, rnd, names, seqs, variants, links, mal, nrp, mw, circular)
        print("%-60s variants %2d (circular %2d)  read paths %3d  informative %3d  pal %s" % (
            label, len(case["variant_paths"]), sum(case["variant_circular"]), len(case["read_paths"]),
            len(case["expected"]["all_sub_paths"]), case["graph"]["palindromic"]))
        cases.append(case)
    out = {"what": "read-path x variant occurrence counting (SURVEY.md 8f item 2)", "reference_version": "0.1.7.1",
           "generator": "oracle/make_golden_subpaths.py", "cases": cases}
    path = os.path.join(GOLDEN, "subpaths.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as h:
        h.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
