#!/usr/bin/env python
"""Developer tool (GPU box): per-batch solver statistics of the cfg3 lock-step selection (bench.selection_rate): lanes,
mean subset width, passes, lane-passes, ms -- where the solver time of a selection goes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from traversome_b200.lanes import LaneContext  # noqa: E402

rows = []
inner = LaneContext._fit_on


DUMP = [int(x) for x in os.environ.get("CFG3_DUMP", "").split(",") if x]      # batch indices of the first run whose passes are dumped
calls = [0]


def fit_on(self, eng, mine, out=None):
    dump = calls[0] in DUMP
    calls[0] += 1
    if dump:
        sys.stderr.write("batch %d lanes %d\n" % (calls[0] - 1, len(mine)))
        sys.stderr.flush()
        os.environ["TVS_TIME_PASSES"] = "1"
        os.environ["TVS_DUMP_PASSES"] = "1"
    try:
        res = inner(self, eng, mine, out=out)
    finally:
        if dump:
            del os.environ["TVS_TIME_PASSES"], os.environ["TVS_DUMP_PASSES"]
    st = eng.last_fit_stats()
    width = float(mine.mask.sum(axis=0).mean()) if mine.mask is not None else float(self.V)
    it = np.asarray(res["iters"])
    rows.append((len(mine), width, st["passes"], st["lane_passes"], st["total_ms"], st["pass_ms"], st["update_ms"], st["compactions"],
                 float(it.mean()), int(it.max()), int(np.unique(mine.cols).shape[0])))
    return res


LaneContext._fit_on = fit_on
out = bench.selection_rate(0, n_rep=int(sys.argv[1]) if len(sys.argv) > 1 else 256, n_groups=1)
n = len(rows) // 3
print("lanes width passes lane_passes total_ms pass_ms update_ms compactions it_mean it_max replicates")
for r in rows[2 * n:]:
    print("%5d %6.1f %5d %8d %8.2f %8.2f %8.2f %3d %7.1f %5d %4d" % r)
tot = np.array([r[4] for r in rows[2 * n:]]).sum()
print("sum total_ms %.1f" % tot, {k: out[k] for k in ("wall_s", "solver_s", "lanes", "batches")})
