#!/usr/bin/env python
"""Developer tool (GPU box): find the lanes of a 20,000-replicate cfg2 fit that do not converge and refit them alone."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic
wl = synthetic.make_config("cfg2")
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
B = 20000
eng.resample(B, wl.n_obs, seed=7, keep_lane0=True)
out = eng.fit()
bad = np.nonzero(out["status"] != 0)[0]
print("bad lanes", bad, out["status"][bad], out["kkt"][bad], out["iters"][bad], eng.last_fit_stats())
w = eng.get_weights()
np.savez_compressed("gpurun_out/bad_lane.npz", w=w[:, bad], lanes=bad, theta=out["theta"][:, bad])
for b in bad:
    eng.set_lanes(w[:, [b]])
    for env in ({}, {"TVS_NEWTON_LANES": "0"}):
        os.environ.pop("TVS_NEWTON_LANES", None); os.environ.update(env)
        o = eng.fit(max_iter=3000)
        print("alone", env, o["status"], o["kkt"], o["iters"], eng.last_fit_stats()["passes"])
