#!/usr/bin/env python
"""Developer tool (GPU box): compare builds of libtvs.so made by tools/build_variant.sh.  For every name given, the variant
is copied over traversome_b200/libtvs.so and a child process fits cfg2 (1,000 lanes; median of 7 fits, per-pass times of
one more fit) and, with `--cfg3`, a 2,048-lane fit of the cfg3 matrix; prints fit time, full-width pass time and a hash of
theta (bitwise guard).  NOTE: it overwrites traversome_b200/libtvs.so with the last variant -- meant for the scratch copy
on the GPU box; rebuild (traversome_b200/csrc/build.sh) after running it anywhere else.  usage: python tools/variant_probe.py [--cfg3] [--cfg4] [--only] base W2 ..."""
import hashlib, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("_INNER") == "1":
    import numpy as np
    from traversome_b200 import _cabi, synthetic
    which, B = sys.argv[1], int(sys.argv[2])
    wl = synthetic.make_config(which)
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
    eng.fit()
    ts = []
    for _ in range(int(os.environ.get("_REPS", "7"))):
        out = eng.fit()
        ts.append(eng.last_fit_stats()["total_ms"])
    st = eng.last_fit_stats()
    res = {"cfg": which, "lanes": B, "fit_ms": round(float(np.median(ts)), 3), "fit_ms_min": round(float(min(ts)), 3), "passes": st["passes"],
           "conv": int((out["status"] == 0).sum()), "iters": round(float(out["iters"].mean()), 3),
           "theta": hashlib.sha256(out["theta"].tobytes()).hexdigest()[:12], "ll": hashlib.sha256(out["loglike"].tobytes()).hexdigest()[:12]}
    print("RES " + json.dumps(res))
    eng.fit(time_passes=True)
    sys.exit(0)
args = sys.argv[1:]
cfg3 = "--cfg3" in args
cfg4 = "--cfg4" in args
only = "--only" in args          # skip cfg2
names = [a for a in args if not a.startswith("--")]
for name in names:
    shutil.copy(os.path.join(ROOT, "build", "variants", "libtvs_%s.so" % name), os.path.join(ROOT, "traversome_b200", "libtvs.so"))
    for which, B, reps in ([] if only else [("cfg2", 1000, 7)]) + ([("cfg3", 2048, 3)] if cfg3 else []) + ([("cfg4", 1248, 1)] if cfg4 else []):
        env = dict(os.environ, _INNER="1", TVS_DUMP_PASSES="1", _REPS=str(reps))
        r = subprocess.run([sys.executable, __file__, which, str(B)], env=env, capture_output=True, text=True)
        txt = r.stdout + r.stderr
        rows = re.findall(r"pass (\d+) width (-?\d+) pass_us ([\d.]+) update_us ([\d.]+)", txt)
        full = [float(p) for _, w, p, _ in rows if abs(int(w)) >= B]
        upd = [float(u) for _, w, _, u in rows if abs(int(w)) >= B]
        res = [l[4:] for l in txt.split("\n") if l.startswith("RES ")]
        print(name, res[-1] if res else txt[-400:], "full-width passes %d, pass_us median %.1f, update_us median %.1f" % (
            len(full), sorted(full)[len(full) // 2] if full else -1, sorted(upd)[len(upd) // 2] if upd else -1), flush=True)
