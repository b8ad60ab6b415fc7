#!/usr/bin/env python
"""Developer tool (GPU box): width / time profile of one fit of a named config (TVS_DUMP_PASSES), summarised by width."""
import os, sys, re, subprocess
if os.environ.get("_INNER") != "1":
    env = dict(os.environ, _INNER="1", TVS_DUMP_PASSES="1")
    out = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    rows = re.findall(r"pass (\d+) width (-?\d+) pass_us ([\d.]+) update_us ([\d.]+)", out.stdout + out.stderr)
    by = {}
    for _, w, p, u in rows:
        k = int(w)
        a = by.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1; a[1] += float(p); a[2] += float(u)
    print("width  passes  pass_ms_total  update_ms_total  pass_us_each")
    for k in sorted(by, key=lambda x: -abs(x)):
        n, p, u = by[k]
        print("%6d %6d %12.1f %12.1f %12.1f" % (k, n, p / 1e3, u / 1e3, p / n))
    print([l for l in (out.stdout + out.stderr).split("\n") if l.startswith("{")][-1:])
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 625
wl = synthetic.make_config(which)
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=777, keep_lane0=False)
eng.fit(time_passes=True)
print(eng.last_fit_stats())
