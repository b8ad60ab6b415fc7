#!/usr/bin/env python
"""Developer tool (GPU box): TVS_DUMP_PASSES trace of one fit of the bench workload for a given resampling seed."""
import os, sys
os.environ["TVS_DUMP_PASSES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from traversome_b200 import _cabi, synthetic  # noqa: E402
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2024 + 7919 * 5
wl = synthetic.make_config("cfg2")
B = synthetic.CONFIGS["cfg2"]["lanes"]
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
eng.resample(B, wl.n_obs, seed=seed, keep_lane0=False)
eng.fit()
out = eng.fit(time_passes=True)
print(eng.last_fit_stats())
slow = np.argsort(out["iters"])[-5:]
print("slowest lanes", slow, out["iters"][slow], out["kkt"][slow], out["status"][slow])
th = out["theta"][:, slow[-1]]
print("support of slowest", int((th > 1e-12).sum()), "tiny", int(((th > 0) & (th < 1e-6)).sum()), np.sort(th)[:12])
