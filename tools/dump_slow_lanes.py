#!/usr/bin/env python
"""Developer tool: save the weights of the slowest-converging lanes of cfg2 for offline analysis."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic
wl = synthetic.make_config("cfg2")
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(1000, wl.n_obs, seed=2024, keep_lane0=True)
out = eng.fit()
w = eng.get_weights()
order = np.argsort(-out["iters"])
sel = list(order[:6]) + list(order[500:502])
print("iters of selected", out["iters"][sel])
np.savez_compressed("gpurun_out/slow_lanes.npz", lanes=np.array(sel), iters=out["iters"], w=w[:, sel], theta=out["theta"][:, sel])
