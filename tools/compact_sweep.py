#!/usr/bin/env python
"""Developer tool (GPU box): fit time against the compaction threshold (TVS_COMPACT_PCT: a lane set is compacted when at
most that share of its lanes is still iterating).  Usage: tools/compact_sweep.py config lanes [pct ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
which, B = sys.argv[1], int(sys.argv[2])
pcts = [int(x) for x in sys.argv[3:]] or [50, 65, 75]
wl = synthetic.make_config(which)
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=777, keep_lane0=False)
eng.fit()
for rep in range(2):
    for p in pcts:
        os.environ["TVS_COMPACT_PCT"] = str(p)
        eng.fit()
        st = eng.last_fit_stats()
        print("%s %d lanes  compact at %d %%: %.2f ms, %d passes, %d lane-passes, %d compactions" % (
            which, B, p, st["total_ms"], st["passes"], st["lane_passes"], st["compactions"]))
