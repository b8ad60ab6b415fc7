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
done, dev_ms, passes, conv, iters = 0, 0.0, 0, 0, []
b = 0
while done < mine and (args.max_batches == 0 or b < args.max_batches):
    n = min(args.batch, mine - done)
    eng.resample(n, wl.n_obs, seed=777 + 100003 * rank + b, keep_lane0=False)
    out = eng.fit()
    st = eng.last_fit_stats()
    dev_ms += st["total_ms"]; passes += st["passes"]; conv += int((out["status"] == 0).sum()); iters.append(out["iters"])
    done += n; b += 1
    if rank == 0:
        print("batch %d: %d lanes %.0f ms %d passes" % (b, n, st["total_ms"], st["passes"]), flush=True)
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
                      "lanes": done_all, "converged": conv_all, "batch": args.batch, "device_s_max_rank": round(dev_ms / 1e3, 3),
                      "fits_per_s": round(done_all / (dev_ms / 1e3), 1), "passes_rank0": passes,
                      "iters_mean_rank0": round(float(iters.mean()), 1), "iters_max_rank0": int(iters.max()),
                      "make_workload_s": round(t_make, 1), "tvs_create_s": round(t_create, 2)}))
