#!/usr/bin/env python
"""Developer tool (GPU box): per-pass width / pass time / update time of one tvs_fit (TVS_DUMP_PASSES)."""
import os, sys
os.environ["TVS_DUMP_PASSES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else synthetic.CONFIGS[which]["lanes"]
wl = synthetic.make_config(which)
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
eng.fit()
eng.fit(time_passes=True)
print(eng.last_fit_stats())
