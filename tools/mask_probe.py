#!/usr/bin/env python
"""Developer tool (GPU box): pass time of one lane count with and without subset masks (cfg3 shape by default) -- the
selection batches carry masks that keep ~75 of 256 variants.  Usage: tools/mask_probe.py [lanes] [config] [keep]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2980
wl = synthetic.make_config(sys.argv[2] if len(sys.argv) > 2 else "cfg3")
keep = int(sys.argv[3]) if len(sys.argv) > 3 else 75
rng = np.random.default_rng(0)
w = synthetic.bootstrap_weights(wl.n_obs, B, seed=3)
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
carrying = np.unique(wl.col[np.repeat(wl.n_obs > 0, np.diff(wl.row_ptr))])


def run(label, mask, theta0=None):
    eng.set_lanes(w, mask=mask)
    eng.fit(max_iter=6, time_passes=True, theta0=theta0)
    st = eng.last_fit_stats()
    print("%-34s %d passes, pass %.3f ms, update %.3f ms" % (label, st["passes"], st["pass_ms"] / st["passes"], st["update_ms"] / st["passes"]))


run("no mask", None)
run("mask of ones", np.ones((wl.V, B), np.uint8))
m = np.zeros((wl.V, B), np.uint8)
for b in range(B):
    m[rng.choice(wl.V, keep, replace=False), b] = 1
run("random %d of %d" % (keep, wl.V), m)
m2 = np.zeros((wl.V, B), np.uint8)
m2[:keep, :] = 1
run("first %d variants, all lanes" % keep, m2)
t0 = m2 / float(keep)
run("first %d + theta0" % keep, m2, theta0=t0)
