#!/usr/bin/env python
"""Developer tool (GPU box): fits/s versus lanes per call (cfg2 matrix)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
wl = synthetic.make_config("cfg2")
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
for B in [int(x) for x in sys.argv[1:]]:
    eng.resample(B, wl.n_obs, seed=7, keep_lane0=True)
    eng.fit()
    out = eng.fit()
    st = eng.last_fit_stats()
    print(json.dumps(dict(B=B, total_ms=round(st["total_ms"], 1), fits_per_s=round(B / st["total_ms"] * 1e3), passes=st["passes"],
                          converged=int((out["status"] == 0).sum()))), flush=True)
