#!/usr/bin/env python
"""Developer tool (GPU box): cfg4 bootstrap replicates started from the full-data optimum versus the uniform start."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 624
wl = synthetic.make_config(sys.argv[2] if len(sys.argv) > 2 else "cfg4")
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.set_lanes(wl.n_obs.reshape(-1, 1).astype(np.int32))
full = eng.fit()
st0 = eng.last_fit_stats()
print("full data: %d passes %.1f ms status %s" % (st0["passes"], st0["total_ms"], full["status"]))
eng.resample(B, wl.n_obs, seed=777, keep_lane0=False)
cold = eng.fit()
st1 = eng.last_fit_stats()
warm = eng.fit(theta0=np.repeat(full["theta"], B, axis=1))
st2 = eng.last_fit_stats()
print("cold: %d passes %.1f ms mean iters %.1f max %d conv %d" % (st1["passes"], st1["total_ms"], cold["iters"].mean(), cold["iters"].max(), (cold["status"] == 0).sum()))
print("warm: %d passes %.1f ms mean iters %.1f max %d conv %d" % (st2["passes"], st2["total_ms"], warm["iters"].mean(), warm["iters"].max(), (warm["status"] == 0).sum()))
print("max |theta_warm - theta_cold| = %.3e, max rel dLL = %.3e" % (np.abs(warm["theta"] - cold["theta"]).max(), np.abs((warm["loglike"] - cold["loglike"]) / cold["loglike"]).max()))
