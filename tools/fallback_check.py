#!/usr/bin/env python
"""Developer tool (GPU box): the non-staged / non-shared-memory pass kernels (large V, dense rows) against the oracle."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
from oracle import loglik  # noqa: E402  (developer tool)
for (V, R, dens, B) in [(1024, 1500, 0.3, 24), (4096, 600, 0.3, 8), (600, 3000, 0.05, 40)]:
    wl = synthetic.make_workload(V, R, dens, seed=V)
    w = synthetic.bootstrap_weights(wl.n_obs, B, seed=1)
    w[:, B - 1] = 0                                   # a lane without any record
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.set_lanes(w)
        out = eng.fit(max_iter=3000)
        st = eng.last_fit_stats()
        ll = eng.loglik(out["theta"][:, :B - 1] if False else np.nan_to_num(out["theta"]))
    A = wl.scipy_csr()
    ref = [loglik.loglik_rows(A, wl.L, w[:, b], 0.0, out["theta"][:, b]) for b in (0, 1)]
    print(json.dumps(dict(V=V, R=wl.R, nnz=wl.nnz, B=B, status=np.bincount(out["status"], minlength=4).tolist(), passes=st["passes"],
                          ms=round(st["total_ms"], 1), kkt=[float(np.nanmax(out["kkt"][:B - 1, 0])), float(np.nanmax(out["kkt"][:B - 1, 1]))],
                          ll_rel=[abs(out["loglike"][b] - ref[b]) / abs(ref[b]) for b in (0, 1)], empty_lane_status=int(out["status"][B - 1]))), flush=True)
