#!/usr/bin/env python
"""Developer tool (GPU box): per-iteration widths and kernel times of one cfg2 fit with an early switch to the Newton phase
(TVS_NEWTON_LANES / TVS_FIRST_GROUP from the environment)."""
import os, sys, re, subprocess
if os.environ.get("_INNER") != "1":
    env = dict(os.environ, _INNER="1", TVS_DUMP_PASSES="1")
    out = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    txt = out.stdout + out.stderr
    rows = re.findall(r"pass (\d+) width (-?\d+) pass_us ([\d.]+) update_us ([\d.]+)", txt)
    newt = re.findall(r"newton (\d+) hess_us ([\d.]+) solve_us ([\d.]+)", txt)
    ni = 0
    for i, w, p, u in rows:
        extra = ""
        if int(w) < 0 and ni < len(newt):
            extra = " hess %s solve %s" % (newt[ni][1], newt[ni][2]); ni += 1
        print("pass %s width %s pass %s upd %s%s" % (i, w, p, u, extra))
    print([l for l in txt.split("\n") if l.startswith("{")][-1:])
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
wl = synthetic.make_config("cfg2")
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(1000, wl.n_obs, seed=2024, keep_lane0=True)
eng.fit()
eng.fit(time_passes=True)
print(eng.last_fit_stats())
