#!/usr/bin/env python
"""Developer tool: read-path x variant counting at pangenome scale on the GPU, with the CPU
oracle timed on a sample of the variants beside it.

Synthetic pangenome: a circular backbone of P positions with `alleles` alternative contigs per position, all links
between consecutive positions; a variant picks one allele per position; read paths are windows of 1..max_window contigs
cut from the variants.  Usage: python tools/subpaths_bench.py [--variants 2048 --positions 400 --read-paths 200000]
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import subpaths as osp                      # noqa: E402  (checker / CPU baseline only)
from traversome_b200 import subpaths as sp              # noqa: E402


def build(args):
    rnd = random.Random(args.seed)
    names = [["p%04da%d" % (p, a) for a in range(args.alleles)] for p in range(args.positions)]
    vertices = [[n, rnd.randint(200, 3000)] for pos in names for n in pos]
    links = []
    for p in range(args.positions):
        q = (p + 1) % args.positions
        for a in names[p]:
            for b in names[q]:
                ov = rnd.choice((0, 0, 21, 55))
                links.append([a, True, b, False, ov])
                links.append([b, False, a, True, ov])
    graph = osp.PlainGraph(vertices, links)
    weights = [0.5 + rnd.random() for _ in range(args.alleles)]
    variants = [tuple((names[p][rnd.choices(range(args.alleles), weights)[0]], True) for p in range(args.positions))
                for _ in range(args.variants)]
    read_paths = {}
    for _ in range(20 * args.read_paths):                 # bounded: the variants may hold fewer distinct windows
        if len(read_paths) >= args.read_paths:
            break
        v = variants[rnd.randrange(args.variants)]
        s, k = rnd.randrange(args.positions), rnd.randint(1, args.max_window)
        w = tuple(v[(s + i) % args.positions] for i in range(k))
        read_paths[osp.canonical(graph, w, rotations=False)] = None
    return graph, variants, list(read_paths)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=2048)
    ap.add_argument("--positions", type=int, default=400)
    ap.add_argument("--alleles", type=int, default=3)
    ap.add_argument("--read-paths", type=int, default=200000)
    ap.add_argument("--max-window", type=int, default=7)
    ap.add_argument("--max-alignment-len", type=int, default=12000)
    ap.add_argument("--cpu-variants", type=int, default=8, help="variants the CPU oracle is timed (and compared) on")
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    t0 = time.time()
    graph, variants, read_paths = build(args)
    t_build = time.time() - t0

    gen = sp.VariantSubPathsGenerator(graph, 10, args.max_alignment_len, read_paths)
    t0 = time.time()
    gen.tables
    t_tables = time.time() - t0
    t0 = time.time()
    gen._prepare()
    t_keys = time.time() - t0
    gen.count(variants[:2])                               # context creation, module load
    t0 = time.time()
    enc = sp.encode_variants(gen.tables, variants)
    t_enc = time.time() - t0
    best = None
    for _ in range(3):
        t0 = time.time()
        counts = gen.count(variants)
        t = time.time() - t0
        if best is None or t < best[0]:
            best = (t, counts)
    t_count, counts = best
    t0 = time.time()
    rows = counts.row_order()
    csr = counts.csr_in_reference_order()
    t_tables_out = time.time() - t0

    sample = list(range(0, args.variants, max(1, args.variants // args.cpu_variants)))[:args.cpu_variants]
    rp_set = set(read_paths)
    t0 = time.time()
    want = [osp.count_read_paths_in_variant(graph, variants[v], rp_set, args.max_alignment_len) for v in sample]
    t_cpu = time.time() - t0
    got = counts.counters() if args.variants <= 64 else None
    dense_rows = {p: i for i, p in enumerate(read_paths)}
    ok = True
    for v, w in zip(sample, want):
        lo_hi = [(r, int(counts.val[k])) for r in [dense_rows[p] for p in w]
                 for k in range(counts.row_ptr[r], counts.row_ptr[r + 1]) if counts.col[k] == v]
        ok &= sorted(lo_hi) == sorted((dense_rows[p], c) for p, c in w.items())
        # nothing else in that column
        ok &= int((counts.col == v).sum()) == len(w)
    out = {
        "workload": "%d variants x %d positions (%d alleles), %d read paths" % (args.variants, args.positions,
                                                                               args.alleles, len(read_paths)),
        "windows": counts.stats["n_windows"], "nnz": counts.stats["nnz"], "informative_read_paths": int(rows.size),
        "kernel_ms": round(counts.stats["kernel_ms"], 3),
        "windows_per_s_kernel": round(counts.stats["n_windows"] / (counts.stats["kernel_ms"] * 1e-3)),
        "count_call_s": round(t_count, 4), "encode_variants_s": round(t_enc, 4), "encode_read_paths_s": round(t_keys, 4),
        "graph_tables_s": round(t_tables, 4), "result_tables_s": round(t_tables_out, 4),
        "cpu_oracle_s_per_variant": round(t_cpu / len(sample), 4), "cpu_variants_timed": len(sample),
        "cpu_oracle_s_all_variants_extrapolated": round(t_cpu / len(sample) * args.variants, 1),
        "speedup_count_call_vs_cpu": round(t_cpu / len(sample) * args.variants / t_count, 1),
        "sample_equal_to_oracle": bool(ok), "build_workload_s": round(t_build, 2),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    if "--build-only" in sys.argv:                        # CPU check of the generator at full size
        sys.argv.remove("--build-only")
        _t = time.time()
        _ap = argparse.Namespace(variants=2048, positions=400, alleles=3, read_paths=200000, max_window=7, seed=7)
        _g, _v, _r = build(_ap)
        print(len(_v), len(_r), round(time.time() - _t, 1), "s")
        sys.exit(0)
    main()
