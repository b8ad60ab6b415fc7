"""Helpers of the read-path x variant counting tests."""
import random

import numpy as np

from oracle.subpaths import PlainGraph, as_path


def load_case(case):
    graph = PlainGraph.from_json(case["graph"])
    variants = [as_path(v) for v in case["variant_paths"]]
    read_paths = [as_path(r) for r in case["read_paths"]]
    return graph, variants, read_paths


def emulate_count(var_ptr, var_tok, var_step, var_circ, n_rows, key_ptr, key_tok, key_row_circ, key_row_lin,
                  max_internal_len):
    """What tvs_subpaths_count + tvs_subpaths_fetch must return for these arrays (include/tvs_subpaths.h), in plain
    Python: used on the CPU to test the host encoders, and on the GPU as the checker of the kernels."""
    table = {}
    for k in range(len(key_row_circ)):
        table[tuple(key_tok[key_ptr[k]:key_ptr[k + 1]].tolist())] = (int(key_row_lin[k]), int(key_row_circ[k]))
    V = len(var_circ)
    windows = []
    for v in range(V):
        toks = var_tok[var_ptr[v]:var_ptr[v + 1]].tolist()
        steps = var_step[var_ptr[v]:var_ptr[v + 1]].tolist()
        n = len(toks)
        for start in range(n):
            seq = [toks[start]]
            j, internal = start, 0
            while internal <= max_internal_len:
                j += 1
                if j == n:
                    if not var_circ[v]:
                        break
                    j = 0
                internal += steps[j]
                seq.append(toks[j])
            windows.append((v, start, seq))
    stride = max(len(w[2]) for w in windows) + 1 if windows else 1
    cnt, first = {}, {}
    n_windows = 0
    for v, start, seq in windows:
        mode = 1 if var_circ[v] else 0
        for ln in range(len(seq), 0, -1):
            n_windows += 1
            hit = table.get(tuple(seq[:ln]))
            if hit is not None and hit[mode] >= 0:
                cell = (hit[mode], v)
                cnt[cell] = cnt.get(cell, 0) + 1
                rank = start * stride + (stride - 1 - ln)
                first[cell] = min(first.get(cell, rank), rank)
    cells = sorted(cnt)
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    for r, _ in cells:
        row_ptr[r + 1] += 1
    np.cumsum(row_ptr, out=row_ptr)
    col = np.array([v for _, v in cells], dtype=np.int32)
    val = np.array([cnt[c] for c in cells], dtype=np.int32)
    rank = np.array([first[c] for c in cells], dtype=np.int64)
    return row_ptr, col, val, rank, {"n_windows": n_windows, "nnz": len(cells), "kernel_ms": 0.0}


def random_graph_case(seed, n_contigs, n_variants, var_len, circular, n_read_paths, max_window, max_alignment_len):
    """A seeded random graph (no palindromic contigs) with variants walking it and read paths cut from them."""
    rnd = random.Random(seed)
    names = ["k%04d" % i for i in range(n_contigs)]
    vertices = [[n, rnd.randint(30, 400)] for n in names]
    variants = []
    for _ in range(n_variants):
        variants.append(tuple((rnd.choice(names), rnd.random() < 0.6) for _ in range(rnd.randint(*var_len))))
    links, seen = [], {}

    def add(a, b):
        key = (a, b)
        mirror = ((b[0], not b[1]), (a[0], not a[1]))
        if key in seen or mirror in seen:
            return
        ov = rnd.choice((0, 0, 5, 11))
        seen[key] = ov
        # connections[...] entries in both directions, as the reference stores them
        links.append([a[0], a[1], b[0], not b[1], ov])
        links.append([b[0], not b[1], a[0], a[1], ov])

    for v in variants:
        for a, b in zip(v[:-1], v[1:]):
            add(a, b)
        if circular:
            add(v[-1], v[0])
    graph = PlainGraph(vertices, links)
    from oracle.subpaths import canonical
    read_paths = {}
    for v in variants:
        n = len(v)
        for _ in range(n_read_paths):
            s, k = rnd.randrange(n), rnd.randint(1, max_window)
            w = [v[(s + i) % n] for i in range(k)] if circular else list(v[s:s + k])
            read_paths[canonical(graph, w, rotations=False)] = None
    read_paths = list(read_paths)
    rnd.shuffle(read_paths)
    return graph, variants, read_paths, max_alignment_len
