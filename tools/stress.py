#!/usr/bin/env python
"""Developer tool (GPU box): convergence stress over shapes, seeds, masks; compares a sample of lanes with the oracle."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
from oracle import exact  # noqa: E402  (developer tool, not product)

rng = np.random.default_rng(123)
bad = 0
cases = [(2, 60, 0.8, 33, None), (3, 200, 0.6, 200, None), (5, 500, 0.4, 1000, 2), (16, 2000, 0.2, 500, None), (33, 3000, 0.15, 300, 5),
         (64, 6000, 0.1, 700, None), (64, 6000, 0.1, 130, 8), (64, 20000, 0.1, 1000, None), (65, 3000, 0.1, 100, None), (48, 4000, 0.3, 260, 12), (96, 3000, 0.08, 40, None), (128, 4000, 0.06, 150, None),
         (129, 3000, 0.06, 60, None)]
for ci, (V, R, dens, B, nt) in enumerate(cases):
    for seed in (1, 2):
        wl = synthetic.make_workload(V, R, dens, seed=1000 * ci + seed, n_true=nt)
        w = synthetic.bootstrap_weights(wl.n_obs, B, seed=seed)
        mask = None
        if seed == 2 and V > 3:
            mask = (rng.random((V, B)) < 0.85).astype(np.uint8)
            mask[:, 0] = 1
        t0 = time.time()
        with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
            eng.set_lanes(w, mask=mask)
            out = eng.fit()
            st = eng.last_fit_stats()
        dt = time.time() - t0
        status = np.bincount(out["status"], minlength=4)
        A = wl.scipy_csr()
        worst = 0.0
        nchk = 0
        for b in rng.choice(B, size=min(B, 6), replace=False):
            m = None if mask is None else mask[:, b].astype(bool)
            try:
                r = exact.exact_fit(A, wl.L, w[:, b], mask=m)
            except ValueError:
                if out["status"][b] != _cabi.INFEASIBLE:
                    bad += 1
                    print("  lane", b, "should be INFEASIBLE, got", out["status"][b])
                continue
            if out["status"][b] == _cabi.INFEASIBLE:
                bad += 1
                print("  lane", b, "wrongly INFEASIBLE")
                continue
            worst = max(worst, float(np.abs(out["theta"][:, b] - r["theta"]).max()))
            nchk += 1
        nbad = int(status[1] + status[2])
        bad += nbad + (worst > 1e-6)
        print(json.dumps(dict(V=V, R=wl.R, B=B, n_true=nt, masked=mask is not None, status=status.tolist(), passes=st["passes"],
                              it_max=int(out["iters"].max()), ms=round(st["total_ms"], 2), checked=nchk, max_dtheta=worst)), flush=True)
print("FAILURES", bad)
