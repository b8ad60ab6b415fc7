#!/usr/bin/env python
"""Developer tool (GPU box): cfg2 fit time versus the polling interval (check_every) and the first poll (TVS_FIRST_GROUP)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("_INNER") == "1":
    from traversome_b200 import _cabi, synthetic
    import numpy as np
    wl = synthetic.make_config("cfg2")
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(1000, wl.n_obs, seed=2024, keep_lane0=True)
    ce = int(os.environ["CE"])
    eng.fit(check_every=ce)
    ts = []
    for _ in range(7):
        out = eng.fit(check_every=ce)
        ts.append(eng.last_fit_stats()["total_ms"])
    st = eng.last_fit_stats()
    print(json.dumps({"ms": round(float(np.median(ts)), 3), "passes": st["passes"], "conv": int((out["status"] == 0).sum())}))
    sys.exit(0)
for ce in (2, 4, 6, 8):
    for fg in (12, 20, 24):
        env = dict(os.environ, _INNER="1", CE=str(ce), TVS_FIRST_GROUP=str(fg))
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        print(ce, fg, r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-300:])
