#!/usr/bin/env python
"""Developer tool (GPU box): fit time and pass counts of the bench workload for the resampling seeds of ranks 0..7
(bench.py uses seed 2024 + 7919 * rank), to see how much of the multi-GPU max-over-ranks is lane imbalance."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
wl = synthetic.make_config("cfg2")
B = synthetic.CONFIGS["cfg2"]["lanes"]
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
for rank in range(8):
    eng.resample(B, wl.n_obs, seed=2024 + 7919 * rank, keep_lane0=(rank == 0))
    eng.fit()
    ms = []
    for _ in range(5):
        out = eng.fit()
        ms.append(eng.last_fit_stats()["total_ms"])
    st = eng.last_fit_stats()
    it = out["iters"]
    print("rank %d  ms %.2f (min %.2f)  passes %d  lane_passes %d  iters mean %.1f p99 %d max %d  compactions %d" % (
        rank, float(np.median(ms)), min(ms), st["passes"], st["lane_passes"], it.mean(), np.percentile(it, 99), it.max(),
        st["compactions"]))
