"""Launched by tests/test_gpu_multi.py with torchrun (one process per GPU, NCCL): the product's multi-GPU path.

1. cfg1: whole-data backward selection + 100 bootstrap replicates of the plastome fixture with the lanes of every batch
   sharded over the ranks (LaneContext(sharder=LaneSharder()), NCCL all_gather of the per-lane results from the device
   buffers): every rank must reproduce the reference's support table (tests/golden/plastome_cfg1.json.gz).
2. A synthetic bootstrap batch (device-drawn replicate columns, ragged over the ranks): the gathered results on every
   rank equal a single-GPU fit of the same columns bit for bit (the fit of a lane does not depend on its neighbours'
   rank), and LaneContext.fit_weights (host weights, the end-to-end path) agrees with it.
3. Several batches in flight (LaneContext.fit_batches / fit_weight_batches, one handle and host thread per batch in
   flight, gathers in batch order): every batch equals the same batch fitted alone, on every rank.
Exit code 0 = all checks passed on this rank."""
import gzip
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loguru import logger  # noqa: E402

logger.remove()
from tests.helpers import fit_context, model_from_tables  # noqa: E402
from traversome_b200 import bootstrap, synthetic  # noqa: E402
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike  # noqa: E402
from traversome_b200.lanes import LaneBatch, LaneContext, LaneSharder  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
try:
    # ---- 1. cfg1 flow, sharded
    with gzip.open(os.path.join(ROOT, "tests", "golden", "plastome_cfg1.json.gz"), "rt") as h:
        plastome = json.load(h)
    L = plastome["variant_sizes"]
    models = [model_from_tables(L, m["bins_list"]) for m in plastome["models"]]
    sharder = LaneSharder()
    ctx = LaneContext(len(L), L, device=local, sharder=sharder)
    kw0 = fit_context(models[0], plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
    full = ModelFitMaxLike(context=ctx, **kw0)
    raw = full.reverse_model_selection(n_proc=1)
    shared = {k: kw0[k] for k in ("variant_paths", "variant_subpath_counters", "repr_to_merged_variants", "be_unidentifiable_to")}
    reps = [(m, fit_context(m, plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])["sbp_to_sbp_id"]) for m in models[1:]]
    s = bootstrap.do_subsampling(reps, shared, full, raw, n_replicate=100, threshold=0.95)
    got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
    flow_ok = got == plastome["final"]["support"] and abs(raw[0][0] - 23 / 35.) < 1e-7 and s.n_replicates_run == 100
    worst = max(abs(v[c] - p) for v, f in zip(s.variant_proportions_reps, plastome["fits"][1:]) for c, p in f["prop"])
    print("rank %d/%d: cfg1 support table %s, max |prop - reference| %.2e, %d batches, %d gathers, %d bytes gathered" % (
        rank, world, "identical" if flow_ok else "DIFFERS", worst, ctx.n_batches, sharder.n_gathers, sharder.bytes_gathered), flush=True)
    ok = ok and flow_ok and worst < 5e-4 and sharder.n_gathers == ctx.n_batches
    ctx.close()

    # ---- 2. sharded bootstrap batch == single-GPU fit of the same columns
    wl = synthetic.make_workload(48, 2500, 0.12, seed=515)
    n_lanes = 64 * world + 3                                   # ragged over the ranks
    ctx2 = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val, device=local, sharder=sharder)
    cols = ctx2.add_resampled(wl.n_obs, n_lanes, seed=99)
    out = ctx2.fit_batch(LaneBatch(cols))
    solo = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val, device=local)
    cols_s = solo.add_resampled(wl.n_obs, n_lanes, seed=99)
    lo, hi = sharder.my_slice(n_lanes)
    ref = solo.fit_batch(LaneBatch(cols_s[lo:hi]))             # this rank's lanes alone: the same columns, the same width
    same = all(np.array_equal(out[k][..., lo:hi] if k == "theta" else out[k][lo:hi], ref[k]) for k in ("theta", "loglike", "iters", "status"))
    conv = bool(ctx2.lane_ok(out).all())
    # every rank received the same gathered block
    digest = torch.tensor([float(np.sum(out["theta"] * np.arange(1, n_lanes + 1))), float(out["loglike"].sum())], dtype=torch.float64, device="cuda")
    every = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(every, digest)
    agree = all(bool((e == every[0]).all()) for e in every)
    # the end-to-end entry point (host weights of this rank's lanes)
    eng = solo.engine()
    eng.set_lanes_columns(solo._slots(eng, np.asarray(cols_s[lo:hi])))
    w_mine = eng.get_weights().astype(np.uint16)
    out_w = ctx2.fit_weights(w_mine, B=n_lanes)
    e2e_same = np.abs(out_w["theta"] - out["theta"]).max() < 1e-9 and np.abs(out_w["loglike"] - out["loglike"]).max() < 1e-6
    print("rank %d/%d: sharded batch of %d lanes: own slice bitwise %s, all converged %s, ranks agree %s, fit_weights agrees %s" % (
        rank, world, n_lanes, same, conv, agree, e2e_same), flush=True)
    ok = ok and same and conv and agree and e2e_same
    # ---- 3. several batches in flight on several handles (LaneContext.fit_batches / fit_weight_batches): gathers in batch
    #         order on every rank, every batch equal to the same batch fitted alone
    sizes = [32 * world + 1, 5, 48 * world, 16 * world + 2, 7 * world, 9]
    col_sets = [ctx2.add_resampled(wl.n_obs, n, seed=300 + j) for j, n in enumerate(sizes)]
    batches = [LaneBatch(c) for c in col_sets]
    alone = [ctx2.fit_batch(b) for b in batches]
    g0 = sharder.n_gathers
    piped = ctx2.fit_batches(batches)
    pipe_same = sharder.n_gathers - g0 == len(batches) and all(
        np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for a, b in zip(alone, piped) for k in ("theta", "loglike", "iters", "status"))
    w_blocks = [sharder.my_slice(len(c)) for c in col_sets]
    # this rank's weights of every batch, taken from the sharded context's own handle
    eng2 = ctx2.engine()
    w_locals = []
    for c, (a, b) in zip(col_sets, w_blocks):
        if b > a:
            eng2.set_lanes_columns(ctx2._slots(eng2, np.asarray(c[a:b])))
            w_locals.append(eng2.get_weights().astype(np.uint16))
        else:
            w_locals.append(np.zeros((wl.R, 0), dtype=np.uint16))
    piped_w = ctx2.fit_weight_batches(w_locals, B=sizes)
    pipe_w_same = all(np.abs(np.asarray(a["theta"]) - np.asarray(b["theta"])).max() < 1e-9 and
                      np.abs(a["loglike"] - b["loglike"]).max() < 1e-6 and np.array_equal(a["status"], b["status"]) for a, b in zip(alone, piped_w))
    print("rank %d/%d: %d batches pipelined over %d handles: equal to one at a time %s, from host weights %s" % (
        rank, world, len(batches), 4, pipe_same, pipe_w_same), flush=True)
    ok = ok and pipe_same and pipe_w_same
    ctx2.close()
    solo.close()
finally:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
