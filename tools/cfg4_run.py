#!/usr/bin/env python
"""Developer tool (GPU box): BASELINE configs[3] -- 2,048 variants x 200k read paths, 10,000 bootstrap replicates
sharded over the ranks of a torchrun job (or `--lanes N` on one GPU).  Every rank holds the whole CSR, resamples its own
lanes on the device and fits them in batches; the device time of the slowest rank is reported.
Usage: [torchrun --nproc-per-node N] tools/cfg4_run.py [--lanes 10000] [--batch 625] [--config cfg4]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from traversome_b200 import _cabi, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=10000)
ap.add_argument("--batch", type=int, default=625)
ap.add_argument("--config", default="cfg4")
ap.add_argument("--max-batches", type=int, default=0, help="stop after this many batches per rank (0 = all)")
ap.add_argument("--handles", type=int, default=1, help="handles (CUDA streams + host threads) per GPU; batches alternate between them")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
wl = synthetic.make_config(args.config)
t_make = time.time() - t0
t0 = time.time()
eng = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L, device=local)
t_create = time.time() - t0
mine = args.lanes // world + (1 if rank < args.lanes % world else 0)
sizes = []
left = mine
while left > 0 and (args.max_batches == 0 or len(sizes) < args.max_batches):
    sizes.append(min(args.batch, left)); left -= sizes[-1]
engines = [eng] + [_cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L, device=local) for _ in range(args.handles - 1)]
import threading
lock = threading.Lock()
state = {"next": 0, "done": 0, "dev_ms": 0.0, "passes": 0, "conv": 0}
iters = []


def worker(e):
    while True:
        with lock:
            b = state["next"]
            if b >= len(sizes):
                return
            state["next"] += 1
        n = sizes[b]
        e.resample(n, wl.n_obs, seed=777 + 100003 * rank + b, keep_lane0=False)
        out = e.fit()
        st = e.last_fit_stats()
        with lock:
            state["dev_ms"] += st["total_ms"]; state["passes"] += st["passes"]; state["conv"] += int((out["status"] == 0).sum())
            state["done"] += n; iters.append(out["iters"])
            if rank == 0:
                print("batch %d: %d lanes %.0f ms %d passes" % (b + 1, n, st["total_ms"], st["passes"]), flush=True)


torch.cuda.synchronize()
t_wall = time.perf_counter()
th = [threading.Thread(target=worker, args=(e,)) for e in engines]
for x in th: x.start()
for x in th: x.join()
torch.cuda.synchronize()
wall_ms = 1e3 * (time.perf_counter() - t_wall)
done, passes, conv = state["done"], state["passes"], state["conv"]
dev_ms = wall_ms if args.handles > 1 else state["dev_ms"]      # with several streams only the wall clock is meaningful
iters = np.concatenate(iters)
t = torch.tensor([dev_ms, float(done), float(conv)], dtype=torch.float64, device="cuda")
if world > 1:
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    dev_ms, done_all, conv_all = float(mx[0]), int(sm[1]), int(sm[2])
else:
    done_all, conv_all = done, conv
if rank == 0:
    print(json.dumps({"config": args.config, "V": int(wl.V), "R": int(wl.R), "nnz": int(wl.nnz), "gpus": world,
                      "lanes": done_all, "converged": conv_all, "batch": args.batch, "handles_per_gpu": args.handles, "device_s_max_rank": round(dev_ms / 1e3, 3),
                      "fits_per_s": round(done_all / (dev_ms / 1e3), 1), "passes_rank0": passes,
                      "iters_mean_rank0": round(float(iters.mean()), 1), "iters_max_rank0": int(iters.max()),
                      "make_workload_s": round(t_make, 1), "tvs_create_s": round(t_create, 2)}))
