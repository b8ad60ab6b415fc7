#!/usr/bin/env python
"""Developer tool (GPU box): timing of tvs_fit variants selected by environment variables."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402

ENVS = ("TVS_NW", "TVS_NO_STAGE", "TVS_LPT", "TVS_RC", "TVS_NO_XS", "TVS_SPLITS", "TVS_NO_COMPACT", "TVS_UPD_V1", "TVS_MAX_SPLITS", "TVS_PASS_V1", "TVS_K")


def run(wl, B, env, label, theta_ref=None, **kw):
    for k in ENVS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    eng.resample(B, wl.n_obs, seed=2024, keep_lane0=True)
    eng.fit(**kw)
    out = eng.fit(time_passes=True, **kw)
    st = eng.last_fit_stats()
    it = out["iters"]
    rec = dict(label=label, passes=st["passes"], pass_ms=round(st["pass_ms"], 2), upd_ms=round(st["update_ms"], 2),
               total_ms=round(st["total_ms"], 2), lane_passes=st["lane_passes"], launches=st["kernel_launches"],
               it_pct=[int(x) for x in np.percentile(it, [50, 90, 99, 100])], status=np.bincount(out["status"], minlength=4).tolist(),
               kkt=[float(out["kkt"][:, 0].max()), float(out["kkt"][:, 1].max())])
    out2 = eng.fit(**kw)
    rec["total_ms_untimed"] = round(eng.last_fit_stats()["total_ms"], 2)
    if theta_ref is not None:
        rec["max_dtheta"] = float(np.abs(out["theta"] - theta_ref).max())
    print(json.dumps(rec), flush=True)
    eng.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    wl = synthetic.make_config(which)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else synthetic.CONFIGS[which]["lanes"]
    print(which, "R", wl.R, "V", wl.V, "nnz", wl.nnz, "B", B, flush=True)
    import itertools
    ref = None
    combos = [({}, "auto")]
    for spec in sys.argv[3:]:
        env = dict(kv.split("=") for kv in spec.split(","))
        combos.append((env, spec))
    for env, label in combos:
        try:
            out = run(wl, B, env, label, theta_ref=None if ref is None else ref["theta"])
            if ref is None:
                ref = out
        except Exception as e:
            print(label, "FAILED", e, flush=True)
