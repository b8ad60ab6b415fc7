#!/usr/bin/env python
"""Developer tool: a short fit on cfg2 for ncu captures (few passes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else synthetic.CONFIGS[which]["lanes"]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
wl = synthetic.make_config(which)
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
out = eng.fit(max_iter=iters)
print("passes", eng.last_fit_stats())
