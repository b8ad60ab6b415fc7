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
def selection_rate(device, n_rep=8):
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
    ctx = LaneContext(wl.V, wl.L, device=device)
    pattern = np.zeros((wl.R, wl.V), dtype=bool)
    for r in range(wl.R):
        lo, hi = int(wl.row_ptr[r]), int(wl.row_ptr[r + 1])
        ctx.matrix.row_index_distinct(dict(zip(wl.col[lo:hi].tolist(), wl.val[lo:hi].tolist())))
        pattern[r, wl.col[lo:hi]] = True
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
    inner = ctx.fit

    def timed(requests):
        t = time.perf_counter()
        out = inner(requests)
        solver[0] += time.perf_counter() - t
        return out
    ctx.fit = timed
    walls = []
    for _ in range(3):                                   # one warm-up run, then the better of two (a 0.5 s figure on a shared host)
        solver[0] = 0.0
        lanes0, batches0 = ctx.n_lanes_fitted, ctx.n_batches
        t0 = time.perf_counter()
        res = bootstrap.run_selections(fitters, criterion="BIC")
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
    return dt, (-float(res.fun) if res is not False else float("nan"))


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
    return lanes / wall, wall, [r[1] for r in res]


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
        rate, wall, ll = cpu_reference_rate(wl, lanes_per_step, cores, seed=k)
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
            "l2": "L2 flushed between timed steps (256 MiB write); within a step the solver re-reads its own weights",
            "parallelism": "lanes sharded over ranks, no data-path collective"}


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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.lanes
    eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L, device=local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    # every rank: its own B bootstrap replicates, drawn on the device (lane 0 of rank 0 = the observed data)
    eng.resample(B, wl.n_obs, seed=2024 + 7919 * rank, keep_lane0=(rank == 0))
    # end-to-end input: the replicate weights in pinned host memory, 16 bits per count (bootstrap counts are small;
    # tvs_set_lanes_u16 widens them on the device)
    w32 = eng.get_weights()
    assert int(w32.max()) <= 65535
    w_host = torch.empty((wl.R, B), dtype=torch.uint16, pin_memory=True)
    w_host.numpy()[:] = w32.astype(np.uint16)
    outs = {"theta": torch.empty((wl.V, B), dtype=torch.float64, pin_memory=True).numpy(),
            "loglike": torch.empty(B, dtype=torch.float64, pin_memory=True).numpy(),
            "kkt": torch.empty((B, 2), dtype=torch.float64, pin_memory=True).numpy(),
            "iters": torch.empty(B, dtype=torch.int32, pin_memory=True).numpy(),
            "status": torch.empty(B, dtype=torch.int32, pin_memory=True).numpy()}

    # ---- warm-up: both timed paths (resident weights; weights uploaded from pinned host memory)
    for _ in range(args.warmup):
        flush.zero_()
        eng.fit(out=outs)
    for _ in range(args.warmup):
        flush.zero_()
        eng.set_lanes(w_host, B=B)
        eng.fit(out=outs)
    eng.resample(B, wl.n_obs, seed=2024 + 7919 * rank, keep_lane0=(rank == 0))      # back to the device-resident lanes
    import gc
    gc.collect()
    gc.disable()                                  # no collector pauses inside the timed regions
    # ---- timed: device-resident weights
    sampler = ClockSampler(range(world) if rank == 0 else ())     # single node: local ranks = GPU indices
    sampler.start()
    barrier()
    dev_ms, launches, passes, lane_passes = [], 0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        out = eng.fit(out=outs)
        st = eng.last_fit_stats()
        dev_ms.append(st["total_ms"])
        launches += st["kernel_launches"]
        passes += st["passes"]
        lane_passes += st["lane_passes"]
    barrier()
    wall = time.perf_counter() - t0
    converged = int((out["status"] == _cabi.CONVERGED).sum())
    iters_mean = float(out["iters"].mean())
    # ---- timed: end to end through the public API, weights in pinned host memory
    e2e_ms = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.set_lanes(w_host, B=B)            # H2D of this step's weights
        eng.fit(out=outs)                     # D2H of theta / loglike / kkt / iters / status into pinned memory
        e2e_ms.append(1e3 * (time.perf_counter() - t1))
    barrier()
    clocks = sampler.stop()
    # ---- the same steps once more with CUDA events around every launch of the dominant kernel (the events
    #      cost ~1 ms per step, so they are kept out of the steps that produce `value`)
    prof = {"pass_ms": 0.0, "update_ms": 0.0, "total_ms": 0.0, "passes": 0, "lane_passes": 0}
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        eng.fit(out=outs, time_passes=True)
        stp = eng.last_fit_stats()
        for k in prof:
            prof[k] += stp[k]
    fp64_peak = _cabi.measure_fp64_peak(local_rank)
    # ---- informational: the same batches through two handles (two CUDA streams, two host threads), i.e. consecutive
    #      batches of a long bootstrap overlapped -- the narrow tail of one hides under the wide phase of the next.
    #      Not the headline: `value` times one batch at a time.
    two_streams = None
    if world == 1 and not args.no_selection:
        try:
            import threading
            eng2 = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L, device=local_rank)
            eng2.resample(B, wl.n_obs, seed=4049, keep_lane0=False)
            eng2.fit()
            n_each = max(4, args.steps // 2)

            def loop(e):
                for _ in range(n_each):
                    e.fit()
            torch.cuda.synchronize()
            th = [threading.Thread(target=loop, args=(e,)) for e in (eng, eng2)]
            t2 = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            torch.cuda.synchronize()
            wall2 = time.perf_counter() - t2
            two_streams = {"value": 2 * n_each * B / wall2, "unit": UNIT, "batches": 2 * n_each, "wall_s": wall2,
                           "note": "two handles fitting %d-lane batches concurrently (wall clock, no L2 flush); "
                                   "traversome_b200.lanes.LaneContext does this for batches of 3,000+ lanes" % B}
            eng2.close()
        except Exception as e:
            two_streams = {"error": "%s: %s" % (type(e).__name__, e)}

    step_ms = float(np.sum(dev_ms)) / args.steps
    e2e_step_ms = float(np.sum(e2e_ms)) / args.steps
    per_rank_ms = [step_ms]
    if world > 1:
        t = torch.tensor([step_ms, e2e_step_ms], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x[0]) for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_step_ms = float(t[0]), float(t[1])
        c = torch.tensor([converged, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        converged, launches = int(c[0]), int(c[1])
    if rank == 0:
        total_lanes = B * world
        value = total_lanes / (step_ms * 1e-3)
        peak, peak_src = measured_peaks()
        # algorithmic bytes of the pass-kernel launches of the profiled steps (SURVEY.md 8d figure at each launch's width)
        alg_bytes = (8.0 * wl.nnz + 4.0 * (wl.R + 1)) * prof["passes"] + (4.0 * wl.R + 16.0 * wl.V) * prof["lane_passes"]
        pass_s = prof["pass_ms"] * 1e-3
        achieved = alg_bytes / pass_s / 1e9 if pass_s > 0 else 0.0
        flops = 4.0 * wl.nnz * prof["lane_passes"] + 6.0 * wl.R * prof["lane_passes"]
        # shared-memory operand bytes: every FMA of the two sparse products reads one 8-byte operand from shared memory
        smem_bytes = 2.0 * 8.0 * wl.nnz * prof["lane_passes"]
        sm_count, smem_bw = 148, 128.0 * 1.965e9          # 128 B/clk/SM at the 1965 MHz boost clock
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "ms_per_step_by_rank": [round(x, 3) for x in per_rank_ms], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, wl),
            "e2e": {"value": total_lanes / (e2e_step_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(wl.R) * B * 2,
                    "step_ms_p50": float(np.median(e2e_ms)), "step_ms_max": float(np.max(e2e_ms)),
                    "d2h_bytes_per_step": int(wl.V) * B * 8 + B * (8 + 16 + 4 + 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "pass3_kernel (fused p = A x, w/p, w log p, t = A^T(w/p))",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": 76.2e6,
                         "traffic_note": "dram read+write of ONE full-width (1000-lane) launch, ncu --set full (profiles/r1_end_pass3_full.csv); "
                                         "algorithmic bytes of that launch: 74.3e6",
                         "avg_launch_us": 1e3 * prof["pass_ms"] / max(prof["passes"], 1), "launches_profiled": prof["passes"],
                         "pass_share_of_step": prof["pass_ms"] / prof["total_ms"] if prof["total_ms"] > 0 else None,
                         "binding_roof_note": "at this lane count the pass is NOT HBM-bound (SURVEY F8): the weights are L2-resident (72 MB) "
                                              "and every FMA of the sparse products needs an 8-byte shared-memory operand; see fp64 / smem below",
                         "fp64": {"achieved_tflops": flops / pass_s / 1e12 if pass_s > 0 else 0.0, "peak_tflops": fp64_peak,
                                  "frac": (flops / pass_s / 1e12 / fp64_peak) if pass_s > 0 and fp64_peak > 0 else None,
                                  "peak_source": "tvs_measure_fp64_peak (DFMA micro-benchmark on this GPU)"},
                         "smem_operand": {"achieved_gbs": smem_bytes / pass_s / 1e9 if pass_s > 0 else 0.0,
                                          "peak_gbs": sm_count * smem_bw / 1e9,
                                          "frac": (smem_bytes / pass_s) / (sm_count * smem_bw) if pass_s > 0 else None,
                                          "peak_source": "128 B/clk/SM x 148 SMs x 1.965 GHz (B300_MICROARCH.md shared-memory figure)"}},
            "solver": {"converged_lanes": converged, "lanes": total_lanes, "mean_iters": iters_mean,
                       "kkt_tol": 1e-9, "passes_per_step": passes / args.steps, "lane_passes_per_step": lane_passes / args.steps,
                       "wall_ms_per_step": 1e3 * wall / args.steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            n_sample = 120                     # ~14 s of single-core CPU work
            rate, wall_cpu, lls = cpu_reference_rate(wl, n_sample, cores, seed=5)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d of the %d bootstrap lanes, one process (the reference's default -p 1; its bootstrap loop is "
                          "serial), %.1f s of CPU work" % (n_sample, B, wall_cpu),
                "note": "restated reference driver (scipy SLSQP multistart) + numpy closure of the reference objective; "
                        "omits the reference's symengine build + LLVM JIT, i.e. faster than the true reference"}
        if two_streams is not None:
            line["two_streams"] = two_streams
        if world == 1 and not args.no_selection:
            try:
                line["selection"] = selection_rate(local_rank)
            except Exception as e:                       # a secondary figure must never cost the headline line
                line["selection"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
