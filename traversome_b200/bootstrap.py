"""Bootstrap / jackknife replicates as batch lanes (replaces the serial loop of
``Traversome.do_subsampling``, traversome.py:501-617, "TODO parallelize bootstrap if necessary").

The reference runs, per replicate, resample -> bins -> PathMultinomialModel -> a complete backward
model selection, one after the other.  Here every replicate's selection is a coroutine
(``ModelFitMaxLike.selection_steps``); all coroutines advance in lock step and the candidate
subsets they ask for in one step -- lanes (replicate, subset) -- are fitted by ONE call of the CUDA
solver on a shared ``LaneContext`` (sharded over the GPUs of the box when the context has a
``LaneSharder``).  The decision logic inside each replicate is the reference's, untouched, so the
selected models and the support table are those of the serial loop.

Two entry points: ``do_subsampling`` takes the replicate models the reference's own code built
(traversome.py:515-539); ``do_subsampling_from_records`` builds the replicates itself from the record pool in
array form (``traversome_b200.resample``: the reference's draws replayed, its bins reproduced bit for bit), which
removes the per-replicate Python object graph.
"""
import sys
import threading
from collections import OrderedDict

import numpy as np
from loguru import logger

from .ModelFitMaxLike import ModelFitMaxLike
from .lanes import LaneContext
from .utils import Criterion


class _Turns(object):
    """Whose batch goes to the GPU next when the replicates advance as several groups: strictly in turn (group 0, 1, ..,
    0, ..; a finished group is skipped), so the sequence of batches -- and with a LaneSharder the sequence of collectives --
    is the same on every run and on every rank, whatever the thread timing."""

    def __init__(self, n):
        self.cond = threading.Condition()
        self.done = [False] * n
        self.turn = 0

    def _advance(self, g):
        n = len(self.done)
        for k in range(1, n + 1):
            if not self.done[(g + k) % n]:
                self.turn = (g + k) % n
                break
        self.cond.notify_all()

    def gate(self, g):
        turns = self

        class _Gate(object):
            def __enter__(self):
                with turns.cond:
                    while turns.turn != g:
                        turns.cond.wait()

            def __exit__(self, *exc):
                with turns.cond:
                    turns._advance(g)
        return _Gate()

    def finish(self, g):
        with self.cond:
            self.done[g] = True
            if self.turn == g:
                self._advance(g)


def _advance_group(ctx, fitters, steps, results, members, gate=None):
    """Lock-step rounds of the replicates `members`: one GPU batch per round for all their pending chunks."""
    pending = {}                                  # replicate -> subsets it waits for
    for i in members:
        try:
            pending[i] = next(steps[i])
        except StopIteration as stop:             # cannot happen before the first fit, kept for symmetry
            results[i] = stop.value
        except Exception as e:
            results[i] = e
    n_rounds = 0
    while pending:
        order = sorted(pending)
        fitted = ctx.fit_items([fitters[i].lane_requests(pending[i]) for i in order], turn=gate)   # the gate: around the solver call only
        n_rounds += 1
        for i, mine in zip(order, fitted):
            fitters[i].n_gpu_fits += len(pending[i])
            try:
                pending[i] = steps[i].send(mine)
            except StopIteration as stop:
                results[i] = stop.value
                del pending[i]
            except Exception as e:
                results[i] = e
                del pending[i]
    return n_rounds


def run_selections(fitters, criterion=Criterion.BIC, n_proc=1, random_size=20, user_fixed_ids=None, chosen_ids=None, groups=None):
    """Backward model selection of every fitter (one per replicate), lock step.

    fitters: ModelFitMaxLike objects built on ONE shared LaneContext.  Returns a list of
    (prop, loglike, criterion) -- or the Exception a replicate raised (the reference would have
    propagated it out of its serial loop at that replicate).

    groups (default: the context's `selection_groups`, 2, when there are at least 64 replicates, else 1): the replicates
    advance as that many interleaved groups, each in lock step on its own host thread, their batches taking the GPU
    strictly in turn (group 0, 1, .., 0, ..; with a LaneSharder this is also the order of the collectives on every rank):
    while the solver works on one group's batch (the C call releases the interpreter lock) the decision logic and lane
    assembly of the other group run on the host.  Measured on BASELINE configs[2], 256 replicates: 3.33 s as one group,
    2.94 s as two; letting the groups' batches run concurrently on separate handles instead is slower (3.12 s: the pass
    kernels of a wide batch fill the GPU, two of them only delay each other and both groups then need the host at once)."""
    if not fitters:
        return []
    ctx = fitters[0]._ensure_lane()
    for f in fitters:
        if f._ensure_lane() is not ctx:
            raise ValueError("run_selections: all fitters must share one LaneContext")
    steps = [f.selection_steps(n_proc, criterion, chosen_ids, random_size, user_fixed_ids) for f in fitters]
    results = [None] * len(fitters)
    if groups is None:
        groups = int(getattr(ctx, "selection_groups", 2)) if len(fitters) >= 64 else 1
    groups = max(1, min(int(groups), len(fitters)))
    if groups == 1:
        n_rounds = _advance_group(ctx, fitters, steps, results, range(len(fitters)))
    else:
        # the candidate permutations come from the process-global generator in the reference (ModelFitMaxLike.py:203); with
        # several host threads the order of its draws would depend on timing: one draw here, one generator per replicate
        base = int(np.random.randint(0, 2 ** 31 - 1))
        pinned = []
        for i, f in enumerate(fitters):
            if f._shuffle is np.random.shuffle:
                f._shuffle = np.random.default_rng([base, i]).shuffle
                pinned.append(f)
        turns = _Turns(groups)
        rounds = [0] * groups
        errors = [None] * groups

        def work(g):
            try:
                rounds[g] = _advance_group(ctx, fitters, steps, results, range(g, len(fitters), groups), gate=turns.gate(g))
            except BaseException as e:            # re-raised in the calling thread
                errors[g] = e
            finally:
                turns.finish(g)
        interval = sys.getswitchinterval()
        sys.setswitchinterval(min(interval, 2e-4))        # the thread that returns from the solver takes the interpreter back at once
        threads = [threading.Thread(target=work, args=(g,)) for g in range(1, groups)]
        try:
            for t in threads:
                t.start()
            work(0)
            for t in threads:
                t.join()
        finally:
            sys.setswitchinterval(interval)
            for f in pinned:
                f._shuffle = np.random.shuffle
        for e in errors:
            if e is not None:
                raise e
        n_rounds = max(rounds)
    logger.debug("run_selections: {} replicates in {} group(s), {} lock-step rounds, {} lanes".format(
        len(fitters), groups, n_rounds, ctx.n_lanes_fitted))
    return results


def check_bs_threshold(count_unique, new_v_prop, n_reps, threshold):
    """Traversome._check_bs_threshold (traversome.py:629-644): False when even if every remaining
    replicate agreed with the current leader its support would stay below `threshold`."""
    tuple_v_chosen = tuple(sorted(list(new_v_prop.keys())))
    count_unique[tuple_v_chosen] = count_unique.get(tuple_v_chosen, 0) + 1
    counts = count_unique.values()
    current_max = max(counts)
    remaining = n_reps - sum(counts)
    return not (float(current_max + remaining) / n_reps < threshold)


def sorting_cid(variant_proportions_reps):
    """Traversome.__sorting_cid (traversome.py:619-627): a universal candidate-id order, first seen
    by decreasing proportion inside each replicate."""
    cid_sorter = {}
    for v_prop in variant_proportions_reps:
        sorted_res = sorted(list(v_prop.items()), key=lambda x: (-x[1], x[0]))
        for cid_, _p in sorted_res:
            if cid_ not in cid_sorter:
                cid_sorter[cid_] = len(cid_sorter)
    return cid_sorter


class BootstrapSummary(object):
    """What do_subsampling leaves on the Traversome object (traversome.py:557-617)."""

    def __init__(self):
        self.variant_proportions_reps = []
        self.bs_eligible = True
        self.cid_sorter = {}
        self.vp_unique_results = {}
        self.vp_unique_results_sorted = []
        self.variant_proportions_best = None
        self.n_replicates_run = 0


def summarize_replicates(variant_proportions_reps, raw_variant_proportions, raw_loglike, raw_criterion, bs_eligible=True,
                         refit=None):
    """The tally of traversome.py:561-617.  `refit(chosen_ids: set) -> (prop, loglike, criterion)`
    is the whole-dataset point estimate used to re-evaluate supported solutions (:592-598)."""
    out = BootstrapSummary()
    out.variant_proportions_reps = list(variant_proportions_reps)
    out.bs_eligible = bs_eligible
    if not out.variant_proportions_reps:
        return out
    num_reps = len(out.variant_proportions_reps)
    out.cid_sorter = sorting_cid(out.variant_proportions_reps)
    for go_r, v_prop in enumerate(out.variant_proportions_reps):
        sorted_res = sorted(list(v_prop.items()), key=lambda x: out.cid_sorter[x[0]])
        tuple_v_chosen, tuple_v_props = zip(*sorted_res)
        if tuple_v_chosen not in out.vp_unique_results:
            out.vp_unique_results[tuple_v_chosen] = {"rep_ids": [], "support": None, "values": []}
        out.vp_unique_results[tuple_v_chosen]["rep_ids"].append(go_r)
        out.vp_unique_results[tuple_v_chosen]["values"].append(tuple_v_props)
    for res_info in out.vp_unique_results.values():
        res_info["support"] = len(res_info["rep_ids"]) / float(num_reps)
    raw_res = tuple(sorted(raw_variant_proportions.keys(), key=lambda x: (out.cid_sorter.get(x, 0), x)))
    out.vp_unique_results_sorted = sorted(
        out.vp_unique_results, key=lambda x: (-out.vp_unique_results[x]["support"], x != raw_res, x))
    if bs_eligible and (len(out.vp_unique_results_sorted) > 1 or out.vp_unique_results_sorted[0] != raw_res):
        logger.info("Re-evaluate the supported models using whole dataset ..")
    for go_s, tuple_v_chosen in enumerate(out.vp_unique_results_sorted):
        chosen_dict = out.vp_unique_results[tuple_v_chosen]
        if tuple_v_chosen == raw_res:
            raw_prop, raw_like, raw_crit = raw_variant_proportions, raw_loglike, raw_criterion
        elif bs_eligible and (chosen_dict["support"] > 0.05 or go_s == 0) and refit is not None:
            raw_prop, raw_like, raw_crit = refit(set(tuple_v_chosen))
        else:
            values = np.average(chosen_dict["values"], axis=0)
            raw_prop = {c_id: values[go_c] for go_c, c_id in enumerate(tuple_v_chosen)}
            raw_like, raw_crit = "*", "*"
        chosen_dict["raw_prop"] = raw_prop
        chosen_dict["raw_like"] = raw_like
        chosen_dict["raw_criterion"] = raw_crit
        if go_s == 0 and tuple_v_chosen != raw_res and chosen_dict["support"] > 0.05:
            out.variant_proportions_best = raw_prop
    return out


def do_subsampling_from_records(pool, pattern, context, full_fit, raw_result, rng, n_replicate, criterion=Criterion.BIC,
                                threshold=0.95, jackknife=0, n_proc=1, user_fixed_ids=None, masked_rows=(), batch=None,
                                loglevel="INFO", device_bins=False):
    """``Traversome.do_subsampling`` from the record pool in array form (traversome_b200.resample.RecordPool): the
    replicates are drawn with the reference's own draws (`rng` = the reference's ``self.random``, e.g.
    ``random.Random(12345)``), binned bit-exactly, registered as weight columns of `context` and selected in lock
    step -- no per-replicate object graph.  `pattern` [rows, V] bool is the non-zero pattern of from_variants in
    ``all_sub_paths`` order (the rows of `pool`); `context.matrix` must hold exactly those rows in that order.
    device_bins: the bin statistics of a whole batch of replicates in ONE device call (include/tvs_pool.h) instead of the
    host array code, replicate by replicate -- same weights, N and observed read paths, C equal to rounding."""
    n_digit = len(str(n_replicate))
    count_unique = {}
    reps = []
    bs_eligible = True
    batch = n_replicate if not batch else int(batch)
    go_bs = 0
    while go_bs < n_replicate and bs_eligible:
        fitters = []
        states = []                                   # rng state after each replicate's draws (see the early stop below)
        want = min(batch, n_replicate - go_bs)

        def draw():
            if jackknife:
                return pool.sample_jackknife(rng, int(pool.n_records / float(n_replicate)))
            return pool.sample_bootstrap(rng, pool.n_records)

        def add(w, C, N, observed):
            k = len(fitters)
            states.append(rng.getstate() if hasattr(rng, "getstate") else None)
            col = context.add_weights(w)
            f = ModelFitMaxLike.from_lane(context, (col, C, N), pattern, np.nonzero(observed)[0], full_fit.repr_to_merged_variants,
                                          full_fit.be_unidentifiable_to, loglevel=loglevel,
                                          bootstrap_mode=f"BS{go_bs + k + 1: 0{n_digit}d}")
            f._shuffle = full_fit._shuffle
            fitters.append(f)

        if device_bins:
            # draw the whole batch first (the generator state after each draw is kept for the early stop), then one device
            # call; an empty replicate is drawn again like traversome.py:528-529, which costs one more (small) device call
            while len(fitters) < want:
                samples, after = [], []
                for _ in range(want - len(fitters)):
                    samples.append(draw())
                    after.append(rng.getstate() if hasattr(rng, "getstate") else None)
                w_all, C_all, N_all, obs_all = pool.lanes_on_device(samples, masked_rows=masked_rows, device=context.device)
                for j in range(len(samples)):
                    if int(N_all[j]) == 0 or not obs_all[:, j].any():
                        continue
                    add(w_all[:, j], float(C_all[j]), int(N_all[j]), obs_all[:, j])
                    states[-1] = after[j]
        else:
            while len(fitters) < want:
                w, C, N, observed = pool.lane_of(draw(), masked_rows)
                if N == 0 or not np.any(observed):            # traversome.py:528-529 `if not sampled_sub_paths: continue`: draw again
                    continue
                add(w, C, N, observed)
        results = run_selections(fitters, criterion=criterion, n_proc=n_proc, user_fixed_ids=user_fixed_ids)
        batch_first = len(reps)
        for res in results:
            if isinstance(res, Exception):
                raise res
            v_prop = res[0]
            reps.append(v_prop)
            if not check_bs_threshold(count_unique, v_prop, n_replicate, threshold):
                if go_bs < n_replicate - 1:
                    logger.info("Sampling terminates due to divergence in bootstraps (see '--bs-threshold'). "
                                "No convincing support can be found given the dataset and parameters. ")
                bs_eligible = False
                # the serial loop stops drawing here: leave the shared generator where it would be (the draws of the
                # replicates of this batch that come after the stop are undone)
                st = states[len(reps) - 1 - batch_first] if len(reps) - 1 - batch_first < len(states) else None
                if st is not None:
                    rng.setstate(st)
                break
            go_bs += 1
    logger.info("Summarizing the replicates ..")
    raw_prop, raw_like, raw_crit = raw_result
    summary = summarize_replicates(
        reps, raw_prop, raw_like, raw_crit, bs_eligible=bs_eligible,
        refit=lambda ids: full_fit.point_estimate(chosen_ids=ids, criterion=criterion))
    summary.n_replicates_run = len(reps)
    return summary


def do_subsampling(replicate_models, fit_kwargs, full_fit, raw_result, criterion=Criterion.BIC, n_replicate=None,
                   threshold=0.95, n_proc=1, user_fixed_ids=None, context=None, batch=None):
    """Batched ``Traversome.do_subsampling``.

    replicate_models  iterable of (model, sbp_to_sbp_id) per replicate, in the reference's order
                      (what traversome.py:520-539 builds for replicate go_bs).
    fit_kwargs        the constructor arguments shared by all replicates: variant_paths,
                      variant_subpath_counters, repr_to_merged_variants, be_unidentifiable_to, loglevel.
    full_fit          the ModelFitMaxLike of the whole dataset (self.max_like_fit), used by the
                      re-evaluation of supported solutions exactly like traversome.py:592-598.
    raw_result        (variant_proportions, res_loglike, res_criterion) of the whole-dataset selection.
    batch             replicates advanced together (default: all).  The reference's early stop
                      (--bs-threshold) is applied after each batch at the index where the serial loop
                      would have stopped, so the surviving replicates are the same.
    """
    replicate_models = list(replicate_models)
    n_replicate = len(replicate_models) if n_replicate is None else int(n_replicate)
    n_digit = len(str(n_replicate))
    ctx = context if context is not None else full_fit._ensure_lane()
    count_unique = {}
    reps = []
    bs_eligible = True
    batch = n_replicate if not batch else int(batch)
    go_bs = 0
    while go_bs < n_replicate and bs_eligible:
        chunk = replicate_models[go_bs: go_bs + batch]
        fitters = [ModelFitMaxLike(model=model, sbp_to_sbp_id=sbp_to_sbp_id, bootstrap_mode=f"BS{go_bs + k + 1: 0{n_digit}d}",
                                   context=ctx, **fit_kwargs)
                   for k, (model, sbp_to_sbp_id) in enumerate(chunk)]
        for f in fitters:
            f._shuffle = full_fit._shuffle
        results = run_selections(fitters, criterion=criterion, n_proc=n_proc, user_fixed_ids=user_fixed_ids)
        for res in results:
            if isinstance(res, Exception):
                raise res                          # the serial loop would have raised at this replicate
            v_prop = res[0]
            reps.append(v_prop)
            if not check_bs_threshold(count_unique, v_prop, n_replicate, threshold):
                if go_bs < n_replicate - 1:
                    logger.info("Sampling terminates due to divergence in bootstraps (see '--bs-threshold'). "
                                "No convincing support can be found given the dataset and parameters. ")
                bs_eligible = False
                break
            go_bs += 1
    logger.info("Summarizing the replicates ..")
    raw_prop, raw_like, raw_crit = raw_result
    summary = summarize_replicates(
        reps, raw_prop, raw_like, raw_crit, bs_eligible=bs_eligible,
        refit=lambda ids: full_fit.point_estimate(chosen_ids=ids, criterion=criterion))
    summary.n_replicates_run = len(reps)
    return summary
