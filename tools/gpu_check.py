#!/usr/bin/env python
"""Developer tool (run on the GPU box): quick correctness + timing sweep of the pass kernel variants.
Not part of the product path or the test-suite; prints one line per configuration."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402


def run(wl, B, env, label, theta_ref=None, check_every=0, history=0):
    for k in ("TVS_LPT", "TVS_RC", "TVS_NO_XS", "TVS_SPLITS", "TVS_NO_COMPACT", "TVS_UPD_V1", "TVS_MAX_SPLITS"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
    eng.fit()                                    # warm-up (allocations, clocks)
    t0 = time.time()
    out = eng.fit(time_passes=True, check_every=check_every, history=history)
    wall = time.time() - t0
    st = eng.last_fit_stats()
    it = out["iters"]
    rec = dict(label=label, B=B, passes=st["passes"], pass_ms=round(st["pass_ms"], 3), total_ms=round(st["total_ms"], 3),
               wall_ms=round(wall * 1e3, 2), ms_per_pass=round(st["pass_ms"] / max(st["passes"], 1), 4),
               iters_min=int(it.min()), iters_med=float(np.median(it)), iters_max=int(it.max()),
               status=np.bincount(out["status"], minlength=4).tolist(), kkt=[float(out["kkt"][:, 0].max()), float(out["kkt"][:, 1].max())])
    rec["iters_pct"] = [int(x) for x in np.percentile(it, [50, 75, 90, 95, 99, 99.9])]
    rec["lane_passes"] = st["lane_passes"]; rec["update_ms"] = round(st["update_ms"], 3); rec["compactions"] = st["compactions"]
    if theta_ref is not None:
        rec["max_dtheta_vs_ref"] = float(np.abs(out["theta"] - theta_ref).max())
    bytes_pass = synthetic.algorithmic_bytes_per_pass(wl.R, wl.V, wl.nnz, B)
    rec["alg_GBps"] = round(bytes_pass / (rec["ms_per_pass"] * 1e-3) / 1e9, 1)
    rec["fits_per_s"] = round(B / (st["total_ms"] * 1e-3), 1)
    print(json.dumps(rec), flush=True)
    eng.close()
    return out


if __name__ == "__main__":
    print("devices", _cabi.device_count(), "fp64 peak TFLOP/s", round(_cabi.measure_fp64_peak(0), 2), flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    wl = synthetic.make_config(which) if which in synthetic.CONFIGS else None
    B = int(sys.argv[2]) if len(sys.argv) > 2 else synthetic.CONFIGS[which]["lanes"]
    print(which, "R", wl.R, "V", wl.V, "nnz", wl.nnz, "N", wl.N, "B", B, flush=True)
    ref = run(wl, B, {}, "default")
    for env, label, kw in [({"TVS_UPD_V1": 1}, "upd_v1", {}), ({}, "m4", {"history": 4}), ({"TVS_LPT": 1, "TVS_RC": 64}, "lpt1_rc64", {}),
                           ({"TVS_LPT": 1, "TVS_RC": 64}, "lpt1_rc64_ce8", {"check_every": 8}), ({"TVS_LPT": 2, "TVS_RC": 64}, "lpt2_rc64", {})]:
        try:
            run(wl, B, env, label, theta_ref=ref["theta"], **kw)
        except Exception as e:
            print(label, "FAILED", e, flush=True)
