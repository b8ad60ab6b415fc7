#!/usr/bin/env python
"""GPU box, launched with torchrun: cfg1 bootstrap flow with the lanes of every batch sharded over the ranks
(NCCL all_gather of the per-lane results); every rank checks the support table against the golden reference run."""
import gzip, json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loguru import logger  # noqa: E402
logger.remove()
from tests.helpers import fit_context, model_from_tables  # noqa: E402
from traversome_b200 import bootstrap  # noqa: E402
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike  # noqa: E402
from traversome_b200.lanes import LaneContext, LaneSharder  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
with gzip.open(os.path.join(ROOT, "tests", "golden", "plastome_cfg1.json.gz"), "rt") as h:
    plastome = json.load(h)
L = plastome["variant_sizes"]
models = [model_from_tables(L, m["bins_list"]) for m in plastome["models"]]
ctx = LaneContext(len(L), L, device=local, sharder=LaneSharder())
kw0 = fit_context(models[0], plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
full = ModelFitMaxLike(context=ctx, **kw0)
t0 = time.time()
raw = full.reverse_model_selection(n_proc=1)
shared = {k: kw0[k] for k in ("variant_paths", "variant_subpath_counters", "repr_to_merged_variants", "be_unidentifiable_to")}
reps = [(m, fit_context(m, plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])["sbp_to_sbp_id"]) for m in models[1:]]
s = bootstrap.do_subsampling(reps, shared, full, raw, n_replicate=100, threshold=0.95)
dt = time.time() - t0
got = [[list(k), s.vp_unique_results[k]["support"]] for k in s.vp_unique_results_sorted]
ok = got == plastome["final"]["support"] and abs(raw[0][0] - 23 / 35.) < 1e-7 and s.n_replicates_run == 100
worst = max(abs(v[c] - p) for v, f in zip(s.variant_proportions_reps, plastome["fits"][1:]) for c, p in f["prop"])
print("rank %d/%d: support table %s, max |prop - reference| %.2e, %d batches, %.2f s" % (
    rank, dist.get_world_size(), "identical" if ok else "DIFFERS", worst, ctx.n_batches, dt), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
