#!/usr/bin/env python
"""Developer tool (GPU box): sparse pass kernels versus the dense DGEMM pass (TVS_DENSE) on one shape:
agreement of the fits and time per pass.  Usage: tools/dense_probe.py V R density lanes [max_iter]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402

V, R, dens, B = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
max_iter = int(sys.argv[5]) if len(sys.argv) > 5 else 0
t0 = time.time()
wl = synthetic.make_workload(V, R, dens, seed=99)
print("workload V %d R %d nnz %d density %.3f (%.1f s)" % (wl.V, wl.R, wl.nnz, wl.nnz / (wl.V * wl.R), time.time() - t0), flush=True)
res = {}
modes = os.environ.get("DENSE_MODES", "0,1").split(",")
for mode in modes:
    os.environ["TVS_DENSE"] = mode
    t0 = time.time()
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    t_create = time.time() - t0
    eng.resample(B, wl.n_obs, seed=5, keep_lane0=True)
    th = np.full((wl.V, B), 1.0 / wl.V)
    ll0 = eng.loglik(th)
    eng.fit(max_iter=max_iter)
    out = eng.fit(max_iter=max_iter)
    st = eng.last_fit_stats()
    res[mode] = (out, ll0)
    print(json.dumps({"dense": mode, "create_s": round(t_create, 2), "total_ms": round(st["total_ms"], 1), "passes": st["passes"],
                      "lane_passes": st["lane_passes"], "ms_per_pass": round(st["total_ms"] / max(st["passes"], 1), 3),
                      "status": np.bincount(out["status"], minlength=4).tolist(), "iters_max": int(out["iters"].max())}), flush=True)
    eng.close()
if len(modes) < 2:
    sys.exit(0)
a, b = res["0"], res["1"]
ok = (a[0]["status"] == 0) & (b[0]["status"] == 0)
print(json.dumps({"loglik_rel_diff_at_uniform": float(np.max(np.abs(a[1] - b[1]) / np.abs(a[1]))),
                  "max_dtheta_converged": float(np.abs(a[0]["theta"][:, ok] - b[0]["theta"][:, ok]).max()) if ok.any() else None,
                  "max_rel_dll_converged": float(np.max(np.abs(a[0]["loglike"][ok] - b[0]["loglike"][ok]) / np.abs(a[0]["loglike"][ok]))) if ok.any() else None,
                  "both_converged": int(ok.sum())}))
