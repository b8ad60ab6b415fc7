#!/usr/bin/env python
"""Developer tool (GPU box): one fit per BASELINE config at a given lane count; prints timing + convergence."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
for spec in sys.argv[1:]:
    which, B = spec.split(":")
    B = int(B)
    t0 = time.time()
    wl = synthetic.make_config(which)
    t1 = time.time()
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    t2 = time.time()
    eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
    t3 = time.time()
    out = eng.fit(time_passes=True, max_iter=int(os.environ.get("MAXIT", "2000")))
    t4 = time.time()
    st = eng.last_fit_stats()
    ll = eng.loglik(out["theta"])
    print(json.dumps(dict(cfg=which, B=B, R=wl.R, V=wl.V, nnz=wl.nnz, gen_s=round(t1 - t0, 1), create_s=round(t2 - t1, 2), resample_s=round(t3 - t2, 2),
                          fit_s=round(t4 - t3, 3), passes=st["passes"], pass_ms=round(st["pass_ms"], 1), upd_ms=round(st["update_ms"], 1),
                          total_ms=round(st["total_ms"], 1), lane_passes=st["lane_passes"], it_pct=[int(x) for x in np.percentile(out["iters"], [50, 90, 100])],
                          status=np.bincount(out["status"], minlength=4).tolist(), kkt=[float(out["kkt"][:, 0].max()), float(out["kkt"][:, 1].max())],
                          ll_selfcheck=float(np.max(np.abs(ll - out["loglike"]) / np.abs(ll))))), flush=True)
    eng.close()
