import os, sys, subprocess, json
sys.path.insert(0, "/root/repo")
if os.environ.get("_INNER") == "1":
    from traversome_b200 import _cabi, synthetic
    import numpy as np
    wl = synthetic.make_config("cfg2")
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(1000, wl.n_obs, seed=2024, keep_lane0=True)
    eng.fit()
    ts = []
    for _ in range(5):
        out = eng.fit()
        ts.append(eng.last_fit_stats()["total_ms"])
    st = eng.last_fit_stats()
    print(json.dumps({"ms": round(float(np.median(ts)), 3), "passes": st["passes"], "conv": int((out["status"] == 0).sum()), "iters": float(out["iters"].mean())}))
    sys.exit(0)
for nl in (128, 256, 400, 600, 1000):
    for fg in (12, 20, 28):
        env = dict(os.environ, _INNER="1", TVS_NEWTON_LANES=str(nl), TVS_FIRST_GROUP=str(fg))
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        print(nl, fg, r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-300:])
