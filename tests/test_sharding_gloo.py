"""CPU, world_size 2 over gloo: lanes of one batch are split over ranks, each rank fits only its
slice, one all_gather returns every lane's result to every rank, and the (redundantly replayed)
selection decisions are identical on all ranks and identical to the single-process run.
The engine is the oracle-backed stand-in of tests/helpers.py (no GPU here); on the B200 box the same
code path runs with backend nccl and the CUDA engine (tests/test_gpu_boundary.py, bench.py --gpus N)."""
import gzip
import json
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _replicates(case, n):
    from tests.helpers import model_from_tables
    rng = np.random.default_rng(3)
    models = []
    for _ in range(n):
        bins = [dict(b, rp_bins=[dict(rp, num_matched=int(rng.poisson(rp["num_matched"])) + 1) for rp in b["rp_bins"]])
                for b in case["bins_list"]]
        models.append(model_from_tables(case["variant_sizes"], bins))
    return models


def _run(ctx, case, models, groups=None):
    from tests.helpers import fit_context
    from traversome_b200 import bootstrap
    from traversome_b200.ModelFitMaxLike import ModelFitMaxLike
    fitters = []
    for m in models:
        kw = fit_context(m, case["be_unidentifiable_to"], case["repr_to_merged_variants"])
        f = ModelFitMaxLike(context=ctx, **kw)
        f._shuffle = lambda x: None
        fitters.append(f)
    return bootstrap.run_selections(fitters, criterion="BIC", groups=groups)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.helpers import OracleEngine
        from traversome_b200.lanes import LaneContext, LaneRequest, LaneSharder
        with gzip.open(os.path.join(GOLDEN, "synthetic_small.json.gz"), "rt") as h:
            case = json.load(h)["cases"][2]
        L = case["variant_sizes"]
        models = _replicates(case, 5)
        # unsharded reference run inside this process
        ctx1 = LaneContext(len(L), L)
        ctx1.engine_factory = OracleEngine
        ref = _run(ctx1, case, models)
        # sharded run
        ctx2 = LaneContext(len(L), L, sharder=LaneSharder())
        ctx2.engine_factory = OracleEngine
        got = _run(ctx2, case, models)
        same = all(list(a[0]) == list(b[0]) and all(abs(a[0][k] - b[0][k]) < 1e-12 for k in a[0]) and a[1] == b[1] and a[2] == b[2]
                   for a, b in zip(ref, got))
        # the replicates as two groups on two host threads: their batches -- and collectives -- strictly in turn on every rank
        ctx3 = LaneContext(len(L), L, sharder=LaneSharder())
        ctx3.engine_factory = OracleEngine
        got = _run(ctx3, case, models, groups=2)
        same = same and ctx3.sharder.n_gathers > ctx2.sharder.n_gathers and all(
            list(a[0]) == list(b[0]) and all(abs(a[0][k] - b[0][k]) < 1e-12 for k in a[0]) and a[1] == b[1] and a[2] == b[2]
            for a, b in zip(ref, got))
        # a ragged batch: 7 lanes over 2 ranks, one of them infeasible (uncovered subset)
        col = ctx2.add_weights(ctx2.dense_columns([0])[:, 0])
        reqs = [LaneRequest(col, 1.5 * i, set(range(len(L))) - {i}) for i in range(6)] + [LaneRequest(col, 0.0, {0})]
        res2 = ctx2.fit(reqs)
        res1 = ctx1.fit([LaneRequest(ctx1.add_weights(ctx1.dense_columns([0])[:, 0]), r.C, r.ids) for r in reqs])
        ragged = all(a.success == b.success and a.status == b.status and (not a.success or (np.array_equal(a.x, b.x) and a.fun == b.fun))
                     for a, b in zip(res1, res2))
        # several independent batches, two in flight on two handles: the gathers stay in batch order on both ranks
        from traversome_b200.lanes import LaneBatch
        cols2 = [ctx2.add_weights(ctx2.dense_columns([j])[:, 0]) for j in range(len(models))]
        bs = [LaneBatch(np.array(cols2[0:3])), LaneBatch(np.array(cols2[3:4])), LaneBatch(np.array(cols2[1:5]), C=np.arange(4.0)),
              LaneBatch(np.array(cols2[0:1])), LaneBatch(np.array(cols2[2:5]))]
        seq = [ctx2.fit_batch(b) for b in bs]
        g0 = ctx2.sharder.n_gathers
        pip = ctx2.fit_batches(bs)
        ragged = ragged and ctx2.sharder.n_gathers - g0 == len(bs) and all(
            np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for a, b in zip(seq, pip) for k in ("theta", "loglike", "status"))
        lanes_here = ctx2._engine.calls
        q.put((rank, bool(same), bool(ragged), [sorted(r[0]) for r in got], lanes_here, res2[-1].success))
    finally:
        dist.destroy_process_group()


def test_two_rank_lane_sharding_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o[0] for o in out] == [0, 1]
    assert all(o[1] for o in out), "sharded selection differs from the single-process run"
    assert all(o[2] for o in out), "ragged sharded batch differs from the single-process run"
    assert out[0][3] == out[1][3], "ranks disagree on the selected models"
    assert all(o[5] is False for o in out)          # the uncovered lane fails identically everywhere


def _worker_shuffle(rank, world, port, q):
    """The DEFAULT shuffle (no pinned permutation) with more candidates than one chunk holds: every rank must draw the
    same permutations (seed broadcast by rank 0), and a rank that holds a different batch must be caught."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        np.random.seed(1000 + 17 * rank)                     # the process-global generator differs per rank on purpose
        from tests.helpers import OracleEngine, fit_context, model_from_workload
        from traversome_b200 import synthetic
        from traversome_b200.ModelFitMaxLike import ModelFitMaxLike
        from traversome_b200.lanes import LaneBatch, LaneContext, LaneSharder
        wl = synthetic.make_workload(9, 260, 0.45, seed=77, n_true=3)
        model = model_from_workload(wl)
        kw = fit_context(model)
        ctx = LaneContext(wl.V, wl.L, sharder=LaneSharder())
        ctx.engine_factory = OracleEngine
        f = ModelFitMaxLike(context=ctx, **kw)
        res = f.reverse_model_selection(n_proc=1, criterion="BIC", random_size=3)     # 9 candidates in chunks of 3: shuffled
        seeds_equal = ctx.shuffle_seed is not None
        # a rank-dependent batch must raise on every rank instead of pairing lanes with the wrong results
        col = f._lane[0]
        caught = False
        try:
            ctx.fit_batch(LaneBatch([col, col], mask=np.eye(wl.V, 2, k=-rank, dtype=np.uint8) + np.eye(wl.V, 2, k=-3, dtype=np.uint8)))
        except RuntimeError as e:
            caught = "different lane batches" in str(e)
        # the same from a thread that fits through a gate (no collective before the fit): the digest inside the gathered block
        import contextlib
        try:
            ctx.fit_batch(LaneBatch([col, col], mask=np.eye(wl.V, 2, k=-rank, dtype=np.uint8) + np.eye(wl.V, 2, k=-3, dtype=np.uint8)),
                          handle=1, gate=contextlib.nullcontext())
            caught = False
        except RuntimeError as e:
            caught = caught and "different lane batches" in str(e)
        q.put((rank, sorted(res[0]), [float(res[0][k]) for k in sorted(res[0])], float(res[1]), float(res[2]), seeds_equal, caught,
               ctx.shuffle_seed))
    finally:
        dist.destroy_process_group()


def test_default_shuffle_is_rank_independent_and_batch_mismatch_is_caught():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_shuffle, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1:5] == out[1][1:5], "ranks selected different models / proportions with the default shuffle"
    assert out[0][7] == out[1][7] and all(o[5] for o in out), "the shuffle seed was not shared"
    assert all(o[6] for o in out), "a rank-dependent batch was not detected"
