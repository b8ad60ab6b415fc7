#!/usr/bin/env python
"""Developer tool (GPU box): one 1,000-lane cfg2 batch fitted as a whole versus as two concurrent halves on two handles
(LaneContext.concurrent_min_lanes), wall clock per step."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from traversome_b200 import synthetic  # noqa: E402
from traversome_b200.lanes import LaneBatch, LaneContext  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
wl = synthetic.make_config("cfg2")
for limit in (0, B // 2):
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    ctx.concurrent_min_lanes = limit
    cols = ctx.add_resampled(wl.n_obs, B, seed=2024)
    batch = LaneBatch(cols)
    if limit:
        ctx._slots(ctx.second_engine(), np.asarray(cols))
    for _ in range(3):
        out = ctx.fit_batch(batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        out = ctx.fit_batch(batch)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("concurrent_min_lanes %d: %.3f ms per %d-lane batch = %.1f k fits/s, converged %d" % (limit, 1e3 * dt, B, B / dt / 1e3, int(ctx.lane_ok(out).sum())))
    ctx.close()
