#!/usr/bin/env python
"""Developer tool (GPU box): a short fit on BASELINE configs[3] (2,048 variants x 200k read paths) for ncu captures of
blk_kernel -- a handful of passes at one lane count.  Usage: tools/profile_blk.py [lanes] [max_iter] [scale] [config]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1248
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
t0 = time.time()
wl = synthetic.make_config(sys.argv[4] if len(sys.argv) > 4 else "cfg4", scale=scale)
t1 = time.time()
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
t2 = time.time()
eng.resample(B, wl.n_obs, seed=777, keep_lane0=False)
eng.fit(max_iter=max_iter, time_passes=True)
st = eng.last_fit_stats()
print("make %.1fs create %.1fs; %d passes, pass %.2f ms each, update %.2f ms each" % (
    t1 - t0, t2 - t1, st["passes"], st["pass_ms"] / max(st["passes"], 1), st["update_ms"] / max(st["passes"], 1)))
