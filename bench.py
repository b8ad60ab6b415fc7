#!/usr/bin/env python
"""bench.py -- bootstrap ML fits/sec of the B200 path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = the maximum-likelihood fit of one batch of B bootstrap replicates (lanes) of the
workload (BASELINE.json configs[1]: 64 candidate variants x 20k read paths, B = 1,000 replicates per
GPU), every lane iterated to the KKT tolerance 1e-9 (the library default).  `value` = lanes fitted per second by the
whole job with the replicate weights already resident in HBM; `e2e` = the same fit driven through
the public API with the weights in pinned HOST memory (H2D of the weights and D2H of the results
inside the timed region).  Weak scaling: every rank fits its own B replicates (different seeds) of
the same read-path x variant matrix; there is no data-path collective, results are gathered once.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bootstrap ML fits/sec"
UNIT = "replicate-fits/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as h:
            d = json.load(h)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json: torch copy_ read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes (read + write) of ONE full-width launch of the dominant kernel from the committed `ncu --set full`
    capture (profiles/): the latest r*_pass3_full.csv.  None when no capture is committed."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_pass3_full.csv")))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reversed(files):
        try:
            found = {}
            for r in csv.reader(open(path)):                 # rows: metric, unit, value
                if len(r) >= 3 and r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    found[r[0]] = float(r[2]) * scale[r[1]]
            if len(found) == 2:
                return sum(found.values()), os.path.relpath(path, ROOT) + " (ncu --set full, one full-width launch, dram read + write)"
        except Exception:
            continue
    return None, "no committed capture"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        """gpu_indices: the GPUs of this job.  ONE nvidia-smi process (on rank 0) watches all of them: a poller per
        rank takes the driver's locks N times as often and shows up in the step time at N = 8."""
        self.gpu_index = ",".join(str(i) for i in gpu_indices)
        self.proc = None
        self.lines = []

    def start(self):
        if not self.gpu_index:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu_index, "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.gpu_index:
            return None
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ model selection
def selection_rate(device, n_rep=256, nit_cfg2=None, n_groups=None):
    """SURVEY.md 8d asks for replicates/s too (one replicate = one complete backward model selection): BASELINE
    configs[2] -- 256 candidate variants (32 carry reads) x 50k read paths -- for `n_rep` bootstrap replicates advanced
    in lock step through the public classes (array-form fitters, traversome_b200.bootstrap.run_selections)."""
    from loguru import logger
    from traversome_b200 import bootstrap, synthetic
    from traversome_b200.ModelFitMaxLike import ModelFitMaxLike
    from traversome_b200.lanes import LaneContext
    logger.disable("traversome_b200")
    wl = synthetic.make_config("cfg3")
    W = synthetic.bootstrap_weights(wl.n_obs, n_rep, seed=3, keep_lane0=True)
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val, device=device)
    pattern = np.zeros((wl.R, wl.V), dtype=bool)
    pattern[np.repeat(np.arange(wl.R), np.diff(wl.row_ptr)), wl.col] = True
    groups = {v: [v] for v in range(wl.V)}
    ident = {v: v for v in range(wl.V)}
    fitters = []
    for b in range(n_rep):
        col = ctx.add_weights(W[:, b])
        f = ModelFitMaxLike.from_lane(ctx, (col, 0.0, int(W[:, b].sum())), pattern, np.nonzero(W[:, b])[0], groups, ident,
                                      bootstrap_mode="BS%d" % (b + 1))
        f._shuffle = lambda x: None
        fitters.append(f)
    ctx.engine()
    solver = [0.0]
    inner = ctx.fit_batch

    def timed(batch, **kw):
        t = time.perf_counter()
        out = inner(batch, **kw)
        solver[0] += time.perf_counter() - t
        return out
    ctx.fit_batch = timed
    walls = []
    for _ in range(3):                                   # one warm-up run, then the better of two (a 0.5 s figure on a shared host)
        solver[0] = 0.0
        lanes0, batches0 = ctx.n_lanes_fitted, ctx.n_batches
        t0 = time.perf_counter()
        res = bootstrap.run_selections(fitters, criterion="BIC", groups=n_groups)
        walls.append((time.perf_counter() - t0, solver[0], ctx.n_lanes_fitted - lanes0, ctx.n_batches - batches0))
    wall, solver[0], n_lanes, n_batches = min(walls[1:])
    sizes = [len(r[0]) for r in res if not isinstance(r, Exception)]
    out = {"workload": "cfg3: backward selection (BIC) from %d candidate variants x %d read paths, %d bootstrap replicates in "
                       "lock step" % (wl.V, wl.R, n_rep),
           "replicates_per_s": n_rep / wall, "lane_fits_per_s": n_lanes / wall, "lanes": n_lanes,
           "batches": n_batches, "wall_s": wall, "solver_s": solver[0], "runs": "1 warm-up + 2 timed, better of the two",
           "selected_variants_median": float(np.median(sizes)) if sizes else None,
           "failures": sum(isinstance(r, Exception) for r in res)}
    ctx.close()
    if nit_cfg2 is not None:
        out["cpu_baseline"] = cpu_eval_extrapolation(wl, nit_cfg2)
        out["cpu_baseline"]["unit"] = "lane-fits/s (full candidate set; leave-one-out lanes are at most as wide)"
    return out


# ------------------------------------------------------------------------------------------ reference arm
def _cpu_fit_one(args):
    """One lane through the restated reference driver (oracle/slsqp_port.py): returns (seconds, loglike)."""
    A_data, A_idx, A_ptr, shape, L, w, seed = args
    import scipy.sparse as sp
    from oracle import slsqp_port
    A = sp.csr_matrix((A_data, A_idx, A_ptr), shape=shape)
    fun, n = slsqp_port.neg_loglike_closure(A, L, w, 0.0)
    try:                                   # one BLAS thread per worker process: no oversubscription
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    t0 = time.perf_counter()
    res = slsqp_port.minimize_neg_likelihood_port(fun, n, rng=np.random.default_rng(seed))
    dt = time.perf_counter() - t0
    if limiter is not None:
        limiter.restore_original_limits()
    return dt, (-float(res.fun) if res is not False else float("nan")), (int(res.nit) if res is not False else 0)


def cpu_reference_rate(wl, lanes, n_proc, seed=0):
    """lane-fits/s of the reference's CPU path on `lanes` bootstrap lanes using n_proc processes."""
    from traversome_b200 import synthetic
    A = wl.scipy_csr()
    w = synthetic.bootstrap_weights(wl.n_obs, lanes, seed=seed + 17, keep_lane0=False)
    jobs = [(A.data, A.indices, A.indptr, A.shape, wl.L, w[:, b].copy(), seed * 1000 + b) for b in range(lanes)]
    t0 = time.perf_counter()
    if n_proc <= 1:
        res = [_cpu_fit_one(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(n_proc) as pool:
            res = pool.map(_cpu_fit_one, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return lanes / wall, wall, [r[1] for r in res], [r[2] for r in res]


def run_reference(args, wl):
    """`--impl reference`: the reference's own CPU implementation of the path (restated driver, see
    oracle/slsqp_port.py) on all host cores; each step = a bounded sample of the workload's lanes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    lanes_per_step = max(cores * 8, 16)
    for _ in range(args.warmup):
        cpu_reference_rate(wl, max(cores, 2), cores, seed=99)
    times, lls = [], []
    for k in range(args.steps):
        rate, wall, ll, _nit = cpu_reference_rate(wl, lanes_per_step, cores, seed=k)
        times.append(wall)
        lls += ll
    total = float(np.sum(times))
    value = lanes_per_step * args.steps / total
    sample = "%d bootstrap lanes per step of %s (full workload: %d lanes/GPU), %d processes" % (
        lanes_per_step, args.workload, args.lanes, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, wl),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "restated reference driver (scipy SLSQP multistart, ftol 1e-5, 6 successes) + numpy closure "
                                     "of the reference objective; omits the reference's symengine build + LLVM JIT per fit"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, wl):
    return {"workload": "%s: %d variants x %d read paths (nnz %d, %d reads), %d bootstrap replicates per GPU, "
                        "full-candidate-set ML fit to KKT 1e-9" % (args.workload, wl.V, wl.R, wl.nnz, wl.N, args.lanes),
            "lanes_per_gpu": args.lanes, "V": wl.V, "R": wl.R, "nnz": wl.nnz,
            "l2": "inputs larger than L2: 8 distinct batches fitted round-robin (consecutive steps read different weights, 8 x lanes x R x 2 bytes "
                  "> 126 MB); no flush between steps.  `one_batch_at_a_time` repeats the measurement with a 256 MiB flush before every step",
            "steps_in_flight": int(os.environ.get("TVS_BENCH_DEPTH", "4")),
            "parallelism": "lanes sharded over ranks, no data-path collective"}


# ---------------------------------------------------------------------------------- other BASELINE configs
def cpu_eval_extrapolation(wl, nit_mean, lanes=8, evals=3, seed=31):
    """CPU baseline for the wide configs (a reference lane-fit takes minutes to hours there):
    seconds per objective evaluation MEASURED on `lanes` bootstrap lanes (numpy closure of the reference objective,
    oracle/slsqp_port.py) x the evaluations of one reference fit.  The reference's SLSQP has no analytic gradient
    (ModelFitMaxLike.py:36-40, jac=False), so one iteration costs V+1 evaluations; 6 successful restarts per fit
    (:31-53).  `nit_mean` = SLSQP iterations per restart measured on cfg2 in this run: a LOWER bound for wider candidate
    sets, i.e. the extrapolated baseline is faster than the reference would be."""
    from oracle import slsqp_port
    from traversome_b200 import synthetic
    A = wl.scipy_csr()
    w = synthetic.bootstrap_weights(wl.n_obs, lanes, seed=seed, keep_lane0=False)
    rng = np.random.default_rng(seed)
    per_eval = []
    for b in range(lanes):
        fun, n = slsqp_port.neg_loglike_closure(A, wl.L, w[:, b], 0.0)
        x0 = rng.random(n)
        x0 /= x0.sum()
        fun(x0)
        t0 = time.perf_counter()
        for _ in range(evals):
            fun(x0)
        per_eval.append((time.perf_counter() - t0) / evals)
    sec_eval = float(np.median(per_eval))
    evals_per_fit = 6.0 * max(nit_mean, 1.0) * (wl.V + 1)
    return {"value": 1.0 / (sec_eval * evals_per_fit), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "EXTRAPOLATED: %.4f s per objective evaluation measured on %d bootstrap lanes x (V+1 = %d evaluations per "
                      "SLSQP iteration, finite-difference gradient) x %.1f iterations per restart (measured on cfg2 in this run; a "
                      "lower bound for wider candidate sets) x 6 restarts" % (sec_eval, lanes, wl.V + 1, nit_mean),
            "sec_per_eval": sec_eval, "evals_per_fit": evals_per_fit}


def pass_roofline(wl, st, fp64_peak, hbm_peak):
    """Roofline figures of the likelihood-pass launches of one profiled fit (SURVEY.md 8d algorithmic bytes / flops)."""
    alg_bytes = (8.0 * wl.nnz + 4.0 * (wl.R + 1)) * st["passes"] + (4.0 * wl.R + 16.0 * wl.V) * st["lane_passes"]
    flops = (4.0 * wl.nnz + 6.0 * wl.R) * st["lane_passes"]
    sec = st["pass_ms"] * 1e-3
    if sec <= 0:
        return None
    return {"bound": "fp64", "achieved": flops / sec / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": flops / sec / 1e12 / fp64_peak if fp64_peak > 0 else None,
            "hbm": {"achieved": alg_bytes / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / sec / 1e9 / hbm_peak},
            "pass_ms_total": st["pass_ms"], "passes": st["passes"], "lane_passes": st["lane_passes"],
            "pass_share_of_fit": st["pass_ms"] / st["total_ms"] if st["total_ms"] > 0 else None}


def run_cfg4(torch, dist, local_rank, rank, world, sharder, lanes_total, fp64_peak, hbm_peak, nit_cfg2):
    """BASELINE configs[3]: 2,048 variants x 200k read paths, `lanes_total` bootstrap replicates STRONG-scaled over the
    ranks through the product's multi-GPU path: LaneContext(sharder=LaneSharder()).fit_batch -- every rank draws and fits
    its slice of the replicates, one NCCL all_gather of {theta, loglike, kkt, iters, status} inside the timed region."""
    from traversome_b200 import synthetic
    from traversome_b200.lanes import LaneBatch, LaneContext
    t0 = time.perf_counter()
    wl = synthetic.make_config("cfg4")
    t_make = time.perf_counter() - t0
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val, device=local_rank, sharder=sharder, max_iter=4000)
    ctx.concurrent_min_lanes = 0
    ctx.time_passes = True                      # events around 18 ms passes cost nothing measurable
    cols = ctx.add_resampled(wl.n_obs, lanes_total, seed=4004)
    batch = LaneBatch(cols)
    lo, hi = (0, lanes_total) if sharder is None else sharder.my_slice(lanes_total)
    t0 = time.perf_counter()
    ctx.make_resident(cols[lo:hi])
    t_setup = time.perf_counter() - t0
    ctx.max_iter = 1                             # warm-up: allocates the solver state, loads the kernels (2 passes)
    ctx.fit_batch(batch)
    ctx.max_iter = 4000
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    g0 = sharder.bytes_gathered if sharder is not None else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ctx.fit_batch(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = ctx.engine().last_fit_stats()
    per_rank = [ms]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank = [float(x[0]) for x in every]
    ms_max = max(per_rank)
    res = None
    if rank == 0:
        ok = ctx.lane_ok(out)
        res = {"workload": "cfg4: %d variants x %d read paths (nnz %d), %d bootstrap replicates strong-scaled over %d GPU(s), "
                           "full-candidate-set ML fit to KKT 1e-9" % (wl.V, wl.R, wl.nnz, lanes_total, world),
               "value": lanes_total / (ms_max * 1e-3), "unit": UNIT, "scaling": "strong", "n_gpus": world, "lanes": lanes_total,
               "ms_per_step": ms_max, "ms_per_step_by_rank": [round(x, 1) for x in per_rank],
               "converged_lanes": int(ok.sum()), "mean_iters": float(out["iters"].mean()), "max_iters": int(out["iters"].max()),
               "collective": {"op": "all_gather_into_tensor (NCCL)" if world > 1 else "none (single GPU)",
                              "bytes_received_per_rank": int((sharder.bytes_gathered - g0) if sharder is not None else 0),
                              "inside_timed_region": True},
               "kernel": "blk_kernel (csrc/tvs_blk.cuh): register accumulators, operand tiles by cp.async.bulk (TMA), forward + backward launch per pass",
               "roofline_rank0": pass_roofline(wl, st, fp64_peak, hbm_peak),
               "setup_s": {"make_workload": round(t_make, 1), "draw_replicates_on_device": round(t_setup, 2)}}
        if world == 1 and nit_cfg2 is not None:
            res["cpu_baseline"] = cpu_eval_extrapolation(wl, nit_cfg2)
    ctx.close()
    return res


def run_cfg5(torch, local_rank, fp64_peak, hbm_peak, nit_cfg2, lanes=256):
    """BASELINE configs[4]: 4,096 variants at 30 % density x 100k read paths, 256 bootstrap lanes on one GPU (dense path)."""
    from traversome_b200 import synthetic
    from traversome_b200.lanes import LaneBatch, LaneContext
    wl = synthetic.make_config("cfg5")
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val, device=local_rank, max_iter=4000)
    ctx.time_passes = True
    cols = ctx.add_resampled(wl.n_obs, lanes, seed=5005)
    batch = LaneBatch(cols)
    ctx.make_resident(cols)
    ctx.max_iter = 1
    ctx.fit_batch(batch)
    ctx.max_iter = 4000
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ctx.fit_batch(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = ctx.engine().last_fit_stats()
    rf = pass_roofline(wl, st, fp64_peak, hbm_peak)
    if rf is not None:      # the dense contraction executes 2*2*R*V flops per lane and pass, whatever the sparsity
        dense_flops = 4.0 * wl.R * wl.V * st["lane_passes"]
        rf["dense_contraction"] = {"executed_tflops": dense_flops / (st["pass_ms"] * 1e-3) / 1e12,
                                   "frac_of_fp64_peak": dense_flops / (st["pass_ms"] * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None}
    res = {"workload": "cfg5: %d variants x %d read paths (nnz %d, %.0f %% dense), %d bootstrap lanes on 1 GPU" % (
               wl.V, wl.R, wl.nnz, 100.0 * wl.nnz / (wl.R * wl.V), lanes),
           "value": lanes / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "converged_lanes": int(ctx.lane_ok(out).sum()),
           "mean_iters": float(out["iters"].mean()), "roofline": rf}
    if nit_cfg2 is not None:
        res["cpu_baseline"] = cpu_eval_extrapolation(wl, nit_cfg2, lanes=8, evals=2)
    ctx.close()
    return res


# ------------------------------------------------------------------------------------------------- ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selection", action="store_true", help="skip the secondary figures (model selection, two streams)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (cfg4 strong scaling, cfg5)")
    ap.add_argument("--cfg4-lanes", type=int, default=10000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from traversome_b200 import synthetic
    if args.lanes <= 0:
        args.lanes = synthetic.CONFIGS[args.workload]["lanes"]
    wl = synthetic.make_config(args.workload)

    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    from traversome_b200 import _cabi
    from traversome_b200.lanes import LaneBatch, LaneContext, LaneSharder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    sharder = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        sharder = LaneSharder()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the product path: one LaneContext per rank over the shared read-path x variant matrix; the lanes of every
    #      batch (B replicates per rank: weak scaling) are split over the ranks by the LaneSharder, each rank fits its slice
    #      and ONE NCCL all_gather returns {theta, loglike, kkt, iters, status} of all lanes to all ranks
    B = args.lanes
    total = B * world
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val, device=local_rank, sharder=sharder)
    ctx.concurrent_min_lanes = 0                  # a batch is never cut in halves here: the batches themselves overlap
    ctx.collect_stats = True
    ctx.pipeline_depth = int(os.environ.get("TVS_BENCH_DEPTH", "4"))
    # N_SETS distinct batches of B replicates per rank, fitted round-robin: a long bootstrap cut into batches.  A step is
    # one batch; consecutive steps read different weights (N_SETS x 36 MB of 16-bit counts on cfg2 > the 126 MB L2), so
    # no step finds its inputs in L2 and no flush is needed between steps.  LaneContext.fit_batches keeps two batches in
    # flight (two handles = two CUDA streams, two host threads; batch k runs on handle k % 2).
    N_SETS = 12 if ctx.pipeline_depth == 3 else 8
    lo, hi = (0, total) if sharder is None else sharder.my_slice(total)
    sets = [ctx.add_resampled(wl.n_obs, total, seed=2024 + j) for j in range(N_SETS)]   # replicate = Philox stream of its column id: the same on any rank
    set_batches = [LaneBatch(c) for c in sets]
    ctx.fit_batches(set_batches)                  # draws every set on the handle that will fit it: resident in HBM from here on
    eng = ctx.engine()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    # end-to-end input: this rank's replicate weights of 4 of the sets in pinned HOST memory, 16 bits per count
    N_HOST = 4
    w_hosts = []
    for j in range(N_HOST):
        e = ctx.engines(ctx.pipeline_depth)[j % ctx.pipeline_depth]
        e.set_lanes_columns(ctx._slots(e, np.asarray(sets[j][lo:hi])))
        w32 = e.get_weights()
        assert int(w32.max()) <= 65535
        wh = torch.empty((wl.R, B), dtype=torch.uint16, pin_memory=True)
        wh.numpy()[:] = w32.astype(np.uint16)
        w_hosts.append(wh)
        del w32
    cols = sets[0]
    batch = set_batches[0]

    def step_batches(n):
        return [set_batches[k % N_SETS] for k in range(n)]

    def step_weights(n):
        return [w_hosts[k % N_HOST] for k in range(n)]

    # ---- warm-up: both timed paths
    ctx.fit_batches(step_batches(args.warmup))
    ctx.fit_weight_batches(step_weights(args.warmup), B=total)
    import gc
    gc.collect()
    gc.disable()                                  # no collector pauses inside the timed regions
    # ---- timed: device-resident replicate weights.  CUDA events around the region of K steps (column gather, fit, NCCL
    #      gather of the results from the device buffers, device->host copy of the gathered block, for every step)
    sampler = ClockSampler(range(world) if rank == 0 else ())     # single node: local ranks = GPU indices
    sampler.start()
    barrier()
    g0 = (sharder.bytes_gathered, sharder.n_gathers) if sharder is not None else (0, 0)
    gs0 = dict(sharder.seconds) if sharder is not None else {}
    s0 = dict(ctx.stats)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    outs = ctx.fit_batches(step_batches(args.steps))
    e1.record()
    torch.cuda.synchronize()
    region_ms = e0.elapsed_time(e1)
    barrier()
    wall = time.perf_counter() - t0
    g1 = (sharder.bytes_gathered, sharder.n_gathers) if sharder is not None else (0, 0)
    gather_ms = {k: 1e3 * (sharder.seconds[k] - gs0[k]) / args.steps for k in gs0} if sharder is not None else None
    s1 = dict(ctx.stats)
    launches = (s1["kernel_launches"] - s0["kernel_launches"]) + 4 * (s1["fits"] - s0["fits"])   # + column gather, lane sums (tvs_set_lanes_columns)
    passes, lane_passes = s1["passes"] - s0["passes"], s1["lane_passes"] - s0["lane_passes"]
    fit_ms = [(s1["total_ms"] - s0["total_ms"]) / max(1, s1["fits"] - s0["fits"])]
    dev_ms = [region_ms / args.steps] * args.steps
    converged = int(sum(int(ctx.lane_ok(o).sum()) for o in outs[-N_SETS:]) / len(outs[-N_SETS:]))    # per batch, all lanes of the job (gathered)
    iters_mean = float(np.mean([o["iters"].mean() for o in outs[-N_SETS:]]))
    out = outs[-1]
    # ---- timed: end to end through the public API, weights in pinned host memory (per step: H2D of this rank's weights,
    #      fit, gather, D2H of all results), two steps in flight -- wall clock over the region
    torch.cuda.synchronize()
    barrier()
    t1 = time.perf_counter()
    ctx.fit_weight_batches(step_weights(args.steps), B=total)
    torch.cuda.synchronize()
    e2e_region_ms = 1e3 * (time.perf_counter() - t1)
    e2e_ms = [e2e_region_ms / args.steps] * args.steps
    barrier()
    clocks = sampler.stop()
    # ---- the same two measurements with ONE batch at a time (round 1's definition of a step: L2 flushed, one handle)
    serial = {}
    ctx.pipeline_batches = False
    n_serial = max(6, args.steps // 3)
    ms_a, ms_b = [], []
    for k in range(n_serial):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record()
        ctx.fit_batch(set_batches[ctx.pipeline_depth * (k % (N_SETS // ctx.pipeline_depth))])        # sets of handle 0
        e1.record()
        torch.cuda.synchronize()
        ms_a.append(e0.elapsed_time(e1))
    for k in range(n_serial):
        flush.zero_()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ctx.fit_weights(w_hosts[0], B=total)
        ms_b.append(1e3 * (time.perf_counter() - t1))
    ctx.pipeline_batches = True
    serial = {"value": total / (float(np.mean(ms_a)) * 1e-3), "e2e": total / (float(np.mean(ms_b)) * 1e-3), "unit": UNIT, "steps": n_serial,
              "ms_per_step": float(np.mean(ms_a)), "e2e_ms_per_step": float(np.mean(ms_b)),
              "note": "this rank only; one batch at a time on one handle, L2 flushed (256 MiB write) between steps"}
    # ---- the same fits once more with CUDA events around every launch of the dominant kernel (the events
    #      cost ~1 ms per step, so they are kept out of the steps that produce `value`)
    prof = {"pass_ms": 0.0, "update_ms": 0.0, "total_ms": 0.0, "passes": 0, "lane_passes": 0}
    eng.set_lanes_columns(ctx._slots(eng, np.asarray(cols[lo:hi])))
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        eng.fit(time_passes=True)
        stp = eng.last_fit_stats()
        for k in prof:
            prof[k] += stp[k]
    fp64_peak = _cabi.measure_fp64_peak(local_rank)
    smem_peak = _cabi.measure_smem_peak(local_rank)
    step_ms = float(np.sum(dev_ms)) / args.steps
    e2e_step_ms = float(np.sum(e2e_ms)) / args.steps
    per_rank_ms = [step_ms]
    per_rank_fit_ms = [float(np.mean(fit_ms))]
    if world > 1:
        t = torch.tensor([step_ms, e2e_step_ms, float(np.mean(fit_ms))], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x[0]) for x in every]
        per_rank_fit_ms = [float(x[2]) for x in every]
        t = t[:2].clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_step_ms = float(t[0]), float(t[1])
        c = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches = int(c[0])
    line = None
    nit_cfg2 = None
    if rank == 0:
        value = total / (step_ms * 1e-3)
        peak, peak_src = measured_peaks()
        # algorithmic bytes of the pass-kernel launches of the profiled steps (SURVEY.md 8d figure at each launch's width)
        alg_bytes = (8.0 * wl.nnz + 4.0 * (wl.R + 1)) * prof["passes"] + (4.0 * wl.R + 16.0 * wl.V) * prof["lane_passes"]
        pass_s = prof["pass_ms"] * 1e-3
        hbm_achieved = alg_bytes / pass_s / 1e9 if pass_s > 0 else 0.0
        flops = 4.0 * wl.nnz * prof["lane_passes"] + 6.0 * wl.R * prof["lane_passes"]
        tflops = flops / pass_s / 1e12 if pass_s > 0 else 0.0
        # shared-memory operand bytes: every FMA of the two sparse products reads one 8-byte operand from shared memory
        smem_bytes = 2.0 * 8.0 * wl.nnz * prof["lane_passes"]
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "ms_per_step_by_rank": [round(x, 3) for x in per_rank_ms],
            "tvs_fit_ms_by_rank": [round(x, 3) for x in per_rank_fit_ms],      # the solver alone on each rank's lanes: the slowest rank sets the step
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, wl),
            "e2e": {"value": total / (e2e_step_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(wl.R) * B * 2 * world,
                    "region_ms": e2e_region_ms,
                    "d2h_bytes_per_step": (int(wl.V) * total * 8 + total * (8 + 16 + 4 + 4)) * world,
                    "api": "LaneContext.fit_weight_batches(pinned uint16 weight blocks) -> result arrays of all lanes of every batch on every rank"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "collective": {"op": "all_gather_into_tensor (NCCL) of the packed per-lane results, from the device buffers tvs_fit wrote"
                                 if world > 1 else "none (single GPU)",
                           "per_step": (g1[1] - g0[1]) / args.steps, "bytes_received_per_rank_per_step": (g1[0] - g0[0]) / args.steps,
                           "inside_timed_region": True, "host_ms_per_step_rank0": gather_ms,
                           "plus": "one 16-byte all_gather per fit_batches call: every rank holds the same batches (sizes, digests)" if world > 1 else None},
            "roofline": {"bound": "fp64", "kernel": "pass3_kernel (fused p = A x, w/p, w log p, t = A^T(w/p))",
                         "achieved": tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tflops / fp64_peak if fp64_peak > 0 else None,
                         "peak_source": "tvs_measure_fp64_peak (DFMA micro-benchmark on this GPU; MEASURED_PEAKS.json has no fp64 entry)",
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_full_width_launch": 8.0 * wl.nnz + 4.0 * (wl.R + 1) + (4.0 * wl.R + 16.0 * wl.V) * B,
                         "avg_launch_us": 1e3 * prof["pass_ms"] / max(prof["passes"], 1), "launches_profiled": prof["passes"],
                         "pass_share_of_step": prof["pass_ms"] / prof["total_ms"] if prof["total_ms"] > 0 else None,
                         "binding_roof_note": "the weights are 16-bit and L2-resident, the pass is not HBM-bound (SURVEY F8); the design reads one 8-byte "
                                              "shared-memory operand per DFMA, which caps it at 16 DFMA/clk/SM = 25 % of the fp64 pipe",
                         "hbm": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src},
                         "smem_operand": {"achieved_gbs": smem_bytes / pass_s / 1e9 if pass_s > 0 else 0.0, "peak_gbs": smem_peak,
                                          "frac": (smem_bytes / pass_s / 1e9) / smem_peak if pass_s > 0 and smem_peak > 0 else None,
                                          "peak_source": "tvs_measure_smem_peak (LDS.128 micro-benchmark on this GPU)"}},
            "solver": {"converged_lanes": converged, "lanes": total, "mean_iters": iters_mean,
                       "kkt_tol": 1e-9, "passes_per_step": passes / args.steps, "lane_passes_per_step": lane_passes / args.steps,
                       "tvs_fit_ms_per_step": float(np.mean(fit_ms)), "wall_ms_per_step": 1e3 * wall / args.steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            n_sample = 120                     # ~14 s of single-core CPU work
            rate, wall_cpu, lls, nits = cpu_reference_rate(wl, n_sample, cores, seed=5)
            nit_cfg2 = float(np.mean(nits))
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d of the %d bootstrap lanes, one process (the reference's default -p 1; its bootstrap loop is "
                          "serial), %.1f s of CPU work" % (n_sample, B, wall_cpu),
                "mean_slsqp_iterations_per_restart": nit_cfg2,
                "note": "restated reference driver (scipy SLSQP multistart) + numpy closure of the reference objective; "
                        "omits the reference's symengine build + LLVM JIT, i.e. faster than the true reference"}
        line["one_batch_at_a_time"] = serial
    ctx.close()
    del flush
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs (each guarded: a secondary figure must never cost the headline line)
    configs = {}
    if not args.no_extra:
        try:
            r4 = run_cfg4(torch, dist, local_rank, rank, world, sharder, args.cfg4_lanes, fp64_peak, measured_peaks()[0], nit_cfg2)
            if rank == 0:
                configs["cfg4"] = r4
        except Exception as e:
            if rank == 0:
                configs["cfg4"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        if world == 1 and not args.no_selection:
            try:
                configs["cfg3"] = selection_rate(local_rank, nit_cfg2=nit_cfg2)
            except Exception as e:
                configs["cfg3"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_extra:
            try:
                configs["cfg5"] = run_cfg5(torch, local_rank, fp64_peak, measured_peaks()[0], nit_cfg2)
            except Exception as e:
                configs["cfg5"] = {"error": "%s: %s" % (type(e).__name__, e)}
        line["configs"] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
