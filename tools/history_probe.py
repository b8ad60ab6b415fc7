#!/usr/bin/env python
"""Developer tool (GPU box): iterations and time versus the L-BFGS memory for the large-V configs."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
which, B = sys.argv[1], int(sys.argv[2])
wl = synthetic.make_config(which)
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
for m in (8, 12, 16):
    eng.fit(history=m, max_iter=3000)
    out = eng.fit(history=m, max_iter=3000, time_passes=True)
    st = eng.last_fit_stats()
    print(json.dumps(dict(cfg=which, B=B, history=m, passes=st["passes"], lane_passes=st["lane_passes"], total_ms=round(st["total_ms"], 1),
                          pass_ms=round(st["pass_ms"], 1), upd_ms=round(st["update_ms"], 1), it_pct=[int(x) for x in np.percentile(out["iters"], [50, 90, 100])],
                          status=np.bincount(out["status"], minlength=4).tolist())), flush=True)
