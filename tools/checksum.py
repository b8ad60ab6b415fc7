#!/usr/bin/env python
"""Developer tool (GPU box): bitwise checksum of the fit results of a few fixed problems (refactoring guard)."""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
h = hashlib.sha256()
for (V, R, d, B, m) in [(64, 20000, 0.1, 1000, 16), (64, 20000, 0.1, 300, 8), (16, 2000, 0.2, 200, 4), (200, 2500, 0.05, 24, 16), (96, 3000, 0.08, 40, 12)]:
    wl = synthetic.make_workload(V, R, d, seed=V + B)
    with _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L) as eng:
        eng.resample(B, wl.n_obs, seed=5, keep_lane0=True)
        out = eng.fit(history=m)
        h.update(out["theta"].tobytes()); h.update(out["loglike"].tobytes()); h.update(out["iters"].tobytes())
        print(V, B, m, eng.last_fit_stats()["passes"], int(out["iters"].max()), hashlib.sha256(out["theta"].tobytes()).hexdigest()[:12])
print("checksum", h.hexdigest()[:16])
