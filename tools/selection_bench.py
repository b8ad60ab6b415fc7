#!/usr/bin/env python
"""Developer tool (GPU box): complete backward model selections (SURVEY 8d "replicates/s") of bootstrap replicates in
lock step.  cfg3-shaped by default: V candidates of which n_true carry reads; every replicate = full fit, zero drops,
leave-one-out rounds until no criterion improvement.  Reports replicates/s, lane-fits/s and the share of wall time
spent outside the solver (Python decision logic + lane assembly)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loguru import logger  # noqa: E402
logger.remove()
from tests.helpers import fit_context, model_from_workload  # noqa: E402
from traversome_b200 import bootstrap, synthetic  # noqa: E402
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike  # noqa: E402
from traversome_b200.lanes import LaneContext  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
n_true = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n_rep = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dens = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
wl = synthetic.make_workload(V, R, dens, seed=4242, n_true=n_true)
W = synthetic.bootstrap_weights(wl.n_obs, n_rep, seed=3, keep_lane0=True)
t0 = time.time()
models = []
shared_fv = []
for b in range(n_rep):
    wb = synthetic.Workload()
    for k in synthetic.Workload.__slots__:
        setattr(wb, k, getattr(wl, k, None))
    wb.n_obs = W[:, b]
    models.append(model_from_workload(wb, shared_from_variants=shared_fv))
t_models = time.time() - t0
ctx = LaneContext(wl.V, wl.L)
t_fit = [0.0]
orig_fit = ctx.fit_batch


def timed_fit(batch, **kw):
    t = time.time()
    out = orig_fit(batch, **kw)
    t_fit[0] += time.time() - t
    return out


ctx.fit_batch = timed_fit
fitters = []
kw0 = fit_context(models[0])
for m in models:
    kw = dict(kw0, model=m, sbp_to_sbp_id={p: i for i, p in enumerate(m.all_sub_paths)})     # shared variant tables, own ids
    f = ModelFitMaxLike(context=ctx, **kw)
    f._shuffle = lambda x: None
    fitters.append(f)
fitters[0]._ensure_lane()
ctx.engine()
t0 = time.time()
res = bootstrap.run_selections(fitters, criterion="BIC")
wall = time.time() - t0
sizes = [len(r[0]) for r in res if not isinstance(r, Exception)]
print(json.dumps(dict(V=wl.V, R=wl.R, nnz=wl.nnz, n_true=n_true, replicates=n_rep, wall_s=round(wall, 3), solver_s=round(t_fit[0], 3),
                      host_share=round(1 - t_fit[0] / wall, 3), lanes=ctx.n_lanes_fitted, batches=ctx.n_batches,
                      replicates_per_s=round(n_rep / wall, 2), lane_fits_per_s=round(ctx.n_lanes_fitted / wall, 1),
                      selected_sizes=[int(min(sizes)), int(np.median(sizes)), int(max(sizes))], failures=sum(isinstance(r, Exception) for r in res),
                      model_build_s=round(t_models, 1))))
