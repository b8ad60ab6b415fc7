#!/usr/bin/env python
"""Developer tool (GPU box): bootstrap replicates started from the full-data optimum, L-BFGS + Newton tail (default)
versus Newton for all lanes, against the cold start."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
wl = synthetic.make_config("cfg2")
ref = None
for label, lanes, warm in [("cold default", 128, False), ("warm default", 128, True), ("warm newton-all", 100000, True),
                           ("warm newton<=512", 512, True)]:
    os.environ["TVS_NEWTON_LANES"] = str(lanes)
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
    theta0 = None
    t_full = 0.0
    if warm:
        w = eng.get_weights()
        e1 = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
        e1.set_lanes(w[:, :1].copy())
        e1.fit()
        o1 = e1.fit()
        t_full = e1.last_fit_stats()["total_ms"]
        th = np.maximum(o1["theta"][:, 0], 1e-4 / wl.V)
        theta0 = np.repeat((th / th.sum())[:, None], B, axis=1)
        e1.close()
    eng.fit(theta0=theta0)
    out = eng.fit(theta0=theta0)
    st = eng.last_fit_stats()
    if ref is None:
        ref = out
    print(json.dumps(dict(label=label, total_ms=round(st["total_ms"], 2), full_data_fit_ms=round(t_full, 2), passes=st["passes"],
                          lane_passes=st["lane_passes"], it_pct=[int(x) for x in np.percentile(out["iters"], [50, 90, 99, 100])],
                          status=np.bincount(out["status"], minlength=4).tolist(),
                          max_dtheta=float(np.abs(out["theta"] - ref["theta"]).max()),
                          max_rel_dll=float(np.max(np.abs(out["loglike"] - ref["loglike"]) / np.abs(ref["loglike"]))))), flush=True)
    eng.close()
