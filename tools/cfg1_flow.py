#!/usr/bin/env python
"""Developer tool (GPU box): wall time of the cfg1 flow (whole-data selection + 100 bootstrap replicates) from the
record pool in array form, repeated to separate first-call costs."""
import gzip, json, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loguru import logger  # noqa: E402
logger.remove()
from traversome_b200 import bootstrap  # noqa: E402
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike  # noqa: E402
from traversome_b200.lanes import LaneContext  # noqa: E402
from traversome_b200.resample import RecordPool  # noqa: E402
with gzip.open(os.path.join(ROOT, "tests", "golden", "plastome_cfg1.json.gz"), "rt") as h:
    plastome = json.load(h)
L = plastome["variant_sizes"]
groups = {int(k): [int(x) for x in v] for k, v in plastome["repr_to_merged_variants"]}
unid = {int(k): int(v) for k, v in plastome["be_unidentifiable_to"]}
for trial in range(3):
    t0 = time.time()
    pool = RecordPool.from_tables(plastome)
    pattern = np.zeros((pool.n_rows, len(L)), dtype=bool)
    ctx = LaneContext(len(L), L)
    for r, rp in enumerate(plastome["read_paths"]):
        fv = {int(v): int(c) for v, c in rp["from_variants"]}
        ctx.matrix.row_index_distinct(fv)
        pattern[r, list(fv)] = True
    w0, C0, N0, obs0 = pool.lane_of(np.arange(pool.n_records))
    full = ModelFitMaxLike.from_lane(ctx, (ctx.add_weights(w0), C0, N0), pattern, np.nonzero(obs0)[0], groups, unid)
    raw = full.reverse_model_selection(n_proc=1)
    t1 = time.time()
    s = bootstrap.do_subsampling_from_records(pool, pattern, ctx, full, raw, random.Random(12345), n_replicate=100)
    t2 = time.time()
    print("trial %d: whole-data selection %.3f s, 100 replicates %.3f s (%d solver batches, %d lanes), support %s" % (
        trial, t1 - t0, t2 - t1, ctx.n_batches, ctx.n_lanes_fitted, [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]))
    ctx.close()
