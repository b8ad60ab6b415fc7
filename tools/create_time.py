#!/usr/bin/env python
"""Developer tool (GPU box): time of tvs_create (host-side build of chunks, Hessian slabs, uploads) per shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
_cabi.device_count()
import torch; torch.zeros(1, device="cuda")       # context creation out of the way
for which in ("cfg2", "cfg3"):
    wl = synthetic.make_config(which)
    for rep in range(2):
        t = time.time()
        eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
        dt = time.time() - t
        eng.close()
        print(which, "V", wl.V, "R", wl.R, "nnz", wl.nnz, "create_s %.3f" % dt, flush=True)
