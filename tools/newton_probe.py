#!/usr/bin/env python
"""Developer tool (GPU box): Newton-from-the-start on all lanes versus the default, accuracy and pass counts."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
wl = synthetic.make_config(which)
res = {}
for label, lanes in [("ref_lbfgs", 0), ("default", 128), ("newton_all", 100000)]:
    os.environ["TVS_NEWTON_LANES"] = str(lanes)
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
    out = eng.fit(tol=1e-12 if label == "ref_lbfgs" else 0.0, max_iter=4000)
    st = eng.last_fit_stats()
    res[label] = out
    rec = dict(label=label, passes=st["passes"], lane_passes=st["lane_passes"], total_ms=round(st["total_ms"], 1),
               it_pct=[int(x) for x in np.percentile(out["iters"], [50, 90, 99, 100])], status=np.bincount(out["status"], minlength=4).tolist(),
               kkt=[float(out["kkt"][:, 0].max()), float(out["kkt"][:, 1].max())])
    if label != "ref_lbfgs":
        rec["max_dtheta_vs_ref"] = float(np.abs(out["theta"] - res["ref_lbfgs"]["theta"]).max())
        rec["max_rel_dll"] = float(np.max(np.abs(out["loglike"] - res["ref_lbfgs"]["loglike"]) / np.abs(out["loglike"])))
    print(json.dumps(rec), flush=True)
    eng.close()
