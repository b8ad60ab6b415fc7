#!/usr/bin/env python
"""Developer tool (GPU box): read-path construction from synthetic alignment records on the
pangenome graph of tools/subpaths_bench.py, array form against the CPU oracle (host-side work, no GPU needed)."""
import argparse, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import subpaths_bench as sb                                  # noqa: E402
from oracle import readpaths as orp                          # noqa: E402  (checker / CPU baseline only)
from traversome_b200 import readpaths as rp                  # noqa: E402


class Rec(object):
    __slots__ = ("path", "p_align_len")

    def __init__(self, path, p_align_len):
        self.path, self.p_align_len = path, p_align_len


ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=300000)
ap.add_argument("--min-counts", type=int, default=3)
args = ap.parse_args()
g, variants, _ = sb.build(argparse.Namespace(variants=256, positions=400, alleles=3, read_paths=1000, max_window=7, seed=7))
rnd = random.Random(5)
records = []
for _ in range(args.records):
    v = variants[rnd.randrange(len(variants))]
    s, k = int(400 * rnd.random() ** 1.5) % 400, rnd.randint(1, 7)
    w = tuple(v[(s + i) % 400] for i in range(k))
    if rnd.random() < 0.5:
        w = tuple((n, not e) for n, e in reversed(w))
    records.append(Rec(w, rnd.randint(500, 20000)))
rp.generate_read_paths(g, records[:2000], min_alignment_counts=args.min_counts)       # context, module load
t0 = time.time()
out = rp.generate_read_paths(g, records, min_alignment_counts=args.min_counts)
t_gpu = time.time() - t0
t0 = time.time()
table, lengths = orp.read_paths_from_records(g, [r.path for r in records], [r.p_align_len for r in records], True, args.min_counts)
t_cpu = time.time() - t0
same = list(out.read_paths.items()) == list(table.items()) and out.align_len_at_path_map == lengths
print(json.dumps({"records": args.records, "min_alignment_counts": args.min_counts, "read_paths": len(out.read_paths),
                  "valid_records": out.num_valid_records, "array_form_s": round(t_gpu, 3), "cpu_oracle_s": round(t_cpu, 2),
                  "ratio": round(t_cpu / t_gpu, 1), "equal_to_oracle": bool(same)}))
