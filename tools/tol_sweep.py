#!/usr/bin/env python
"""Developer tool: iteration counts and accuracy versus the KKT tolerance (cfg2, 1000 lanes)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic
os.environ["TVS_LPT"] = "1"; os.environ["TVS_RC"] = "64"
wl = synthetic.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
ref = eng.fit(tol=1e-12, max_iter=3000)
print("ref status", np.bincount(ref["status"], minlength=4).tolist(), "iters max", int(ref["iters"].max()), "kkt", ref["kkt"].max(0))
for tol in (1e-10, 3e-10, 1e-9, 3e-9, 1e-8, 1e-7):
    eng.fit(tol=tol)
    out = eng.fit(tol=tol, check_every=8)
    st = eng.last_fit_stats()
    it = out["iters"]
    print(json.dumps(dict(tol=tol, passes=st["passes"], total_ms=round(st["total_ms"], 2), lane_passes=st["lane_passes"],
                          iters_pct=[int(x) for x in np.percentile(it, [50, 90, 99, 100])],
                          status=np.bincount(out["status"], minlength=4).tolist(),
                          max_dtheta=float(np.abs(out["theta"] - ref["theta"]).max()),
                          max_rel_dll=float(np.max(np.abs(out["loglike"] - ref["loglike"]) / np.abs(ref["loglike"]))))), flush=True)
