#!/usr/bin/env python
"""Developer tool: iteration tail and accuracy versus the dual-feasibility tolerance multiplier."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic
wl = synthetic.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
os.environ["TVS_DUAL_MULT"] = "100"
ref = eng.fit(tol=1e-13, max_iter=5000)
print("ref status", np.bincount(ref["status"], minlength=4).tolist(), "iters max", int(ref["iters"].max()), "kkt", ref["kkt"].max(0))
for tol, mult in [(1e-10, 100), (1e-10, 1000), (1e-10, 10000), (1e-10, 100000), (1e-9, 100), (1e-9, 1000), (1e-9, 10000), (1e-8, 100), (1e-8, 1000)]:
    os.environ["TVS_DUAL_MULT"] = str(mult)
    eng.fit(tol=tol)
    out = eng.fit(tol=tol)
    st = eng.last_fit_stats()
    it = out["iters"]
    d = np.abs(out["theta"] - ref["theta"]).max(0)
    print(json.dumps(dict(tol=tol, mult=mult, passes=st["passes"], total_ms=round(st["total_ms"], 2), lane_passes=st["lane_passes"],
                          iters_pct=[int(x) for x in np.percentile(it, [50, 90, 99, 100])],
                          status=np.bincount(out["status"], minlength=4).tolist(),
                          max_dtheta=float(d.max()), p99_dtheta=float(np.percentile(d, 99)),
                          max_rel_dll=float(np.max(np.abs(out["loglike"] - ref["loglike"]) / np.abs(ref["loglike"]))))), flush=True)
