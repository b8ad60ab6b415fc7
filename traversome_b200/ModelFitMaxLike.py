"""``ModelFitMaxLike`` -- drop-in replacement of ``traversome/ModelFitMaxLike.py`` on the GPU.

Same constructor, methods, attributes, log lines and exceptions as the reference class
(ModelFitMaxLike.py:61-580), but

  * the symengine expression build + LLVM lambdify (``get_neg_likelihood_of_var_freq``, :515-558)
    becomes an integer CSR build (``traversome_b200.csr``) uploaded once per model;
  * the scipy SLSQP multistart (``minimize_neg_likelihood``, :19-58) becomes ONE batched call of
    the CUDA solver (``tvs_fit``), one "lane" per candidate-variant subset, so that all
    leave-one-out candidates of a backward-selection step are fitted together;
  * ``n_proc`` is accepted and ignored (lanes replace the Pool/dill workers of :219-262).

There is no CPU fallback: constructing the engine raises when ``libtvs.so`` or a CUDA device is
missing.
"""
import os
from collections import OrderedDict
from typing import OrderedDict as typingODict
from typing import Set, Union

import numpy as np
from loguru import logger

from .lanes import BlockResult, FitResult, IdSet, LaneBlock, LaneContext, LaneRequest, LeaveOneOut  # noqa: F401  (FitResult re-exported)
from .utils import Criterion, LogLikeFuncInfo, aic, bic


def minimize_neg_likelihood(neg_loglike_func, num_variables, verbose, err_queue=None):
    """Module-level entry point kept for API compatibility (ModelFitMaxLike.py:19-58).

    The reference receives an opaque compiled callable; here the callable returned by
    ``ModelFitMaxLike.get_neg_likelihood_of_var_freq`` carries its lane description, and the fit is
    done by the CUDA solver.  Returns a FitResult, or False when the fit failed."""
    try:
        fitter = getattr(neg_loglike_func, "tvs_fit", None)
        if fitter is None:
            raise TypeError("minimize_neg_likelihood: expected the callable produced by "
                            "traversome_b200.ModelFitMaxLike.get_neg_likelihood_of_var_freq")
        res = fitter()
        return res if res.success else False
    except Exception as e:
        if err_queue:
            err_queue.put(e)
        else:
            raise e


def prune_model_bins(model):
    """The in-place pruning get_like_formula performs on the model (ModelGenerator.py:138-151)."""
    kept = []
    for bins in model.bins_list:
        bins.rp_bins[:] = [rp for rp in bins.rp_bins if not (rp.num_possible_X < 1)]
        if bins.rp_bins:
            kept.append(bins)
    model.bins_list[:] = kept


def _debug_logging():
    """True when a DEBUG record written from this module would reach some loguru handler.  The reference writes several
    DEBUG lines per candidate subset (each O(candidates) to format); when nobody listens the selection loop skips
    building them -- and builds the proportion dictionaries of the ACCEPTED candidate only -- without changing any
    decision."""
    try:
        core = logger._core
        if core.min_level > 10:                              # no handler at DEBUG or below
            return False
        dotted = __name__ + "."                              # logger.disable("traversome_b200") and friends
        for dotted_module_name, status in core.activation_list:
            if dotted[: len(dotted_module_name)] == dotted_module_name:
                return bool(status)
        return True
    except Exception:
        return True


class ModelFitMaxLike(object):
    """Find the variant proportions that maximise the likelihood (batched, on one B200)."""

    # solver settings (class level so that callers of the reference signature need not know them)
    tol = 1e-9      # KKT tolerance; observed |theta error| <= ~20 x tol
    max_iter = 2000
    history = 16

    def __init__(self,
                 model,
                 variant_paths,
                 variant_subpath_counters,
                 sbp_to_sbp_id,
                 repr_to_merged_variants,
                 be_unidentifiable_to,
                 loglevel="INFO",
                 bootstrap_mode=False,
                 device=0,
                 context=None):
        self.model = model
        self.variant_paths = variant_paths
        self.num_put_variants = len(variant_paths)
        self.all_sub_paths = model.all_sub_paths
        self.variant_subpath_counters = variant_subpath_counters
        self.sbp_to_sbp_id = sbp_to_sbp_id
        self.repr_to_merged_variants = repr_to_merged_variants
        self.be_unidentifiable_to = be_unidentifiable_to
        self.loglevel = loglevel
        self.bootstrap_mode = bootstrap_mode
        self.device = device

        # to be generated
        self.variant_percents = None
        self.observed_sbp_id_set = set()

        # res without model selection
        self.pe_neg_loglike_obj = None
        self.pe_best_proportions = None

        self.__warning_sent = False
        # `context` = the CSR + GPU engine; shared by the models of every bootstrap replicate when the
        # caller batches them (traversome_b200.bootstrap), otherwise private to this object
        self._context = context
        self._own_context = context is None
        self._lane = None            # (weight column id, C, N) of this model inside the context
        self._shuffle = np.random.shuffle   # ModelFitMaxLike.py:203; tests may pin the permutation
        self.n_gpu_fits = 0          # lanes fitted so far (diagnostics)
        self._cover_cache = None
        self._cover_base = None
        # backward selection starts each candidate subset from the accepted model's optimum (see _start_point)
        self.warm_start = os.environ.get("TVS_WARM_START", "1") != "0"
        self._warm = None            # {representative id: proportion} of the accepted model, inside a selection only
        self._warm_vec = None
        self._batch_x = None

    @classmethod
    def from_lane(cls, context, lane, pattern, observed_rows, repr_to_merged_variants, be_unidentifiable_to,
                  loglevel="INFO", bootstrap_mode=False):
        """A fitter for one replicate given in ARRAY form (traversome_b200.resample.RecordPool.lane_of) instead of
        model objects: `lane` = (weight column id in `context`, C, N); `pattern` [rows, V] bool = read path r occurs
        in variant v (the non-zero pattern of from_variants); `observed_rows` = the replicate's read paths.  Same
        decision logic, logging and results as a fitter built from a PathMultinomialModel."""
        self = cls.__new__(cls)
        self.model = None
        self.variant_paths = [None] * int(pattern.shape[1])
        self.num_put_variants = int(pattern.shape[1])
        self.all_sub_paths = None
        self.variant_subpath_counters = None
        self.sbp_to_sbp_id = None
        self.repr_to_merged_variants = repr_to_merged_variants
        self.be_unidentifiable_to = be_unidentifiable_to
        self.loglevel = loglevel
        self.bootstrap_mode = bootstrap_mode
        self.device = context.device
        self.variant_percents = None
        self.observed_sbp_id_set = set(int(r) for r in observed_rows)
        self.pe_neg_loglike_obj = None
        self.pe_best_proportions = None
        self._ModelFitMaxLike__warning_sent = False
        self._context = context
        self._own_context = False
        self._lane = (int(lane[0]), float(lane[1]), int(lane[2]))
        self._shuffle = np.random.shuffle
        self._pin_shuffle()
        self.n_gpu_fits = 0
        observed = np.zeros(pattern.shape[0], dtype=bool)
        observed[list(self.observed_sbp_id_set)] = True
        pattern = np.asarray(pattern, dtype=bool)
        # the variant-major copy of the pattern is shared by all replicates of the context
        shared = getattr(context, "_pattern_by_variant", None)
        if shared is None or shared[0] is not pattern:
            shared = context._pattern_by_variant = (pattern, np.ascontiguousarray(pattern.T))
        self._cover_cache = (self._cover_key(), pattern, observed, shared[1])
        self._cover_base = None
        self.warm_start = os.environ.get("TVS_WARM_START", "1") != "0"
        self._warm = None
        self._warm_vec = None
        self._batch_x = None
        self._array_form = True
        return self

    # ------------------------------------------------------------------ context / lanes
    def _ensure_lane(self):
        if self._lane is None:
            prune_model_bins(self.model)
            if self._context is None:
                self._context = LaneContext(self.num_put_variants, self.model.variant_sizes, device=self.device,
                                            tol=self.tol, max_iter=self.max_iter, history=self.history)
            self._lane = self._context.add_model(self.model)
            self._pin_shuffle()
        return self._context

    def _pin_shuffle(self):
        """Inside a sharded job every rank replays the selection of every replicate and the lanes of a batch are matched
        by POSITION, so the candidate permutation (ModelFitMaxLike.py:203, process-global np.random in the reference) must
        be the same on all ranks: a generator seeded with the context's broadcast seed and this model's column id."""
        seed = getattr(self._context, "shuffle_seed", None)
        if seed is not None and self._shuffle is np.random.shuffle:
            self._shuffle = np.random.default_rng([int(seed), int(self._lane[0])]).shuffle

    def close(self):
        if self._context is not None and self._own_context:
            self._context.close()

    def lane_requests(self, subsets):
        """Lane descriptions of ML fits restricted to subsets[b] (sets of representative ids, or IdSet)."""
        self._ensure_lane()
        col, C, _N = self._lane
        if isinstance(subsets, LeaveOneOut):                    # a whole chunk of leave-one-out candidates: array form
            return LaneBlock(col, C, subsets, self._warm_dense())
        return [LaneRequest(col, C, ids, self._start_point(ids)) for ids in subsets]

    def _warm_dense(self):
        """The accepted model's optimum as a dense [V] vector (None outside a selection / with warm starts off)."""
        warm = self._warm
        if not warm or not self.warm_start:
            return None
        if self._warm_vec is None or self._warm_vec[0] is not warm:
            vec = np.zeros(self.num_put_variants, dtype=np.float64)
            vec[np.fromiter(warm.keys(), dtype=np.int64, count=len(warm))] = np.fromiter(warm.values(), dtype=np.float64, count=len(warm))
            self._warm_vec = (warm, vec)
        return self._warm_vec[1]

    def _start_point(self, ids):
        """Starting proportions of a candidate subset during backward selection: the accepted model's optimum restricted
        to the subset (a leave-one-out optimum is close to it), floored so that no component starts at the stationary
        point z = 0 of the solver's parametrisation.  None (uniform start) outside a selection."""
        vec = self._warm_dense()
        if vec is None:
            return None
        arr = ids.arr if isinstance(ids, IdSet) else np.array(sorted(ids), dtype=np.int64)
        x = vec[arr]
        total = x.sum()
        if not np.isfinite(total) or total <= 0.0:
            return None
        x = np.maximum(x / total, 1e-3 / len(x))
        return x / x.sum()

    def _fit_subsets(self, subsets):
        """One GPU call: lane b = ML fit restricted to subsets[b].  Returns a list of FitResult
        (x ordered by increasing variant id)."""
        ctx = self._ensure_lane()
        if ctx.matrix.R == 0:
            raise Exception("Likelihood maximization failed.")   # nothing to fit
        results = ctx.fit_items([self.lane_requests(subsets)])[0]
        self.n_gpu_fits += len(subsets)
        return results

    # ------------------------------------------------------------------ public API
    def point_estimate(self,
                       chosen_ids: set = None,
                       criterion=Criterion.BIC):
        if chosen_ids:
            chosen_ids = {self.be_unidentifiable_to[variant_id] for variant_id in chosen_ids}
        else:
            chosen_ids = {variant_id for variant_id in self.repr_to_merged_variants}
        self.variant_percents = ["P" + str(variant_id) if variant_id in chosen_ids else False
                                 for variant_id in range(self.num_put_variants)]
        if self.bootstrap_mode:
            logger.debug("Generating the likelihood function .. ")
        else:
            logger.info("Generating the likelihood function .. ")
        self.pe_neg_loglike_obj = self.get_neg_likelihood_of_var_freq(within_variant_ids=chosen_ids)
        if self.bootstrap_mode:
            logger.debug("Maximizing the likelihood function .. ")
        else:
            logger.info("Maximizing the likelihood function .. ")
        success_run = minimize_neg_likelihood(
            neg_loglike_func=self.pe_neg_loglike_obj.loglike_func,
            num_variables=len(chosen_ids),
            verbose=self.loglevel in ("TRACE", "ALL"))
        if success_run:
            use_prop, echo_prop, this_like, this_criteria = \
                self.__summarize_like_and_criteria(success_run, chosen_ids, criterion, self.pe_neg_loglike_obj)
            return use_prop, this_like, this_criteria
        else:
            raise Exception("Likelihood maximization failed.")

    def reverse_model_selection(self,
                                n_proc,
                                criterion=Criterion.BIC,
                                chosen_ids: Union[typingODict[int, bool], Set] = None,
                                random_size: int = 20,
                                user_fixed_ids: Union[list, tuple, set, None] = None):
        """Backward stepwise selection by AIC/BIC (ModelFitMaxLike.py:140-322).

        Decision logic identical to the reference; every chunk of leave-one-out candidates
        (``random_size`` of them) is fitted as one batch of GPU lanes instead of sequentially /
        through Pool workers.  ``n_proc`` is ignored."""
        steps = self.selection_steps(n_proc, criterion, chosen_ids, random_size, user_fixed_ids)
        try:
            subsets = next(steps)
            while True:
                subsets = steps.send(self._fit_subsets(subsets))
        except StopIteration as stop:
            return stop.value

    def selection_steps(self, n_proc, criterion=Criterion.BIC, chosen_ids=None, random_size=20, user_fixed_ids=None):
        """reverse_model_selection as a coroutine: yields the list of candidate subsets whose fits it
        needs next, is sent the list of FitResult, and returns (prop, loglike, criterion) through
        StopIteration.  `traversome_b200.bootstrap` advances the coroutines of all replicates in lock
        step so that their candidate lanes share one GPU batch."""
        if random_size != 0 and n_proc > random_size and not self.__warning_sent:
            self.__warning_sent = True
        diff_tolerance = 1e-6
        if chosen_ids:
            chosen_ids = {self.be_unidentifiable_to[variant_id] for variant_id in chosen_ids}
        else:
            chosen_ids = {variant_id for variant_id in self.repr_to_merged_variants}
        chosen_ids = set(chosen_ids)
        self.variant_percents = ["P" + str(variant_id) if variant_id in chosen_ids else False
                                 for variant_id in range(self.num_put_variants)]
        logger.debug("Test variants {}".format(list(chosen_ids)))

        if user_fixed_ids:
            indispensable_ids = {u_id: True for u_id in user_fixed_ids}
        else:
            indispensable_ids = {}

        first_sets = [set(chosen_ids)]
        self._warm = None
        previous_prop, previous_echo, previous_like, previous_criteria = \
            self.__summarize_batch(first_sets, (yield first_sets), criterion)[0]
        self._warm = self._batch_x[0]
        self.__drop_zero_variants(chosen_ids, previous_prop, previous_echo, diff_tolerance, indispensable_ids)

        while len(chosen_ids) > 1:
            logger.debug("Trying dropping {} variant(s) ..".format(self.num_put_variants - len(chosen_ids) + 1))
            chosen_ids_sorted = sorted(chosen_ids)
            self._set_cover_base(chosen_ids)
            chosen_rd_list = list(range(len(chosen_ids_sorted)))
            self._shuffle(chosen_rd_list)
            changed = False
            rs = random_size if random_size > 0 else len(chosen_rd_list)
            verbose = _debug_logging()
            chosen_arr = np.array(chosen_ids_sorted, dtype=np.int64)
            for go_rd_sp in range(0, len(chosen_rd_list), rs):
                this_rd_ids = chosen_rd_list[go_rd_sp: go_rd_sp + rs]
                # -- which candidates of this chunk get a fit (same screening as __test_one_drop, :349-370)
                to_fit = []
                label = ", ".join([self.__str_rep_id(_c_i) for _c_i in chosen_ids_sorted]) if verbose else None
                for rd_id in this_rd_ids:
                    var_id = chosen_ids_sorted[rd_id]
                    if var_id not in indispensable_ids:
                        if self._cover_without(var_id):
                            if verbose:
                                logger.debug("Test variants [{}] - {}".format(label, self.__str_rep_id(var_id)))
                            to_fit.append((var_id, rd_id))           # leave-one-out subset = the sorted accepted set minus position rd_id
                        else:
                            indispensable_ids[var_id] = True
                            if verbose:
                                logger.debug("Test variants [{}] - {}: skipped for necessary subpath(s) (case *)"
                                             .format(label, self.__str_rep_id(var_id)))
                    elif verbose:
                        logger.debug("Test variants [{}] - {}: skipped for necessary subpath(s) (case **)"
                                     .format(label, self.__str_rep_id(var_id)))
                # -- one batch of lanes for the whole chunk; the criterion of every candidate, the proportion tables of
                #    the best one only (the reference builds them per candidate, :448-510, and reads the best's)
                if to_fit:
                    id_sets = LeaveOneOut(chosen_arr, [rd for _, rd in to_fit])
                    results = (yield id_sets)
                    if isinstance(results, BlockResult) and not verbose:
                        # the criterion of every candidate in one array expression (ModelFitMaxLike.py:472-497 per lane)
                        if criterion not in ("AIC", "BIC"):
                            raise Exception("Invalid criterion {}".format(criterion))
                        if not self.bootstrap_mode:
                            for _ in range(len(to_fit)):
                                logger.info("Maximizing the likelihood function for {} variants".format(len(chosen_arr) - 1))
                        if not bool(np.all(results.success)):
                            raise Exception("Likelihood maximization failed.")
                        k_par = self.__variable_size(id_sets[0])
                        like = -results.fun
                        crit_arr = aic(loglike=like, len_param=k_par) if criterion == "AIC" else \
                            bic(loglike=like, len_param=k_par, len_data=self._lane[2])       # same expression, elementwise
                        best = int(np.argmin(crit_arr))                              # first minimum = sorted(...)[0] of the reference
                        best_val = float(crit_arr[best])
                        best_res = results[best]
                        best_ids = id_sets[best]
                    else:
                        per_lane = [results[j] for j in range(len(to_fit))]
                        sets = [id_sets[j] for j in range(len(to_fit))]
                        crit_vals = self.__criteria_of_batch(sets, per_lane, criterion, verbose)
                        best = min(range(len(to_fit)), key=crit_vals.__getitem__)  # first minimum = sorted(...)[0] of the reference
                        best_val, best_res, best_ids = crit_vals[best], per_lane[best], sets[best]
                    best_drop_id = to_fit[best][0]
                    if best_val < previous_criteria:
                        previous_criteria = best_val
                        previous_like = -best_res.fun
                        previous_prop, previous_echo = self.__summarize_run_prop(best_res, best_ids)
                        chosen_ids.remove(best_drop_id)
                        self._warm = dict(zip(best_ids.arr.tolist(), (float(v) for v in best_res.x))) or self._warm
                        log = logger.debug if self.bootstrap_mode else logger.info
                        log("Drop {}".format(self.__str_rep_id(best_drop_id)))
                        log("Proportions: " + ", ".join(["%s:%.4f" % (_id, _p) for _id, _p, in previous_echo.items()]))
                        log("Log-likelihood: %s" % previous_like)
                        self.__drop_zero_variants(
                            chosen_ids, previous_prop, previous_echo, diff_tolerance, indispensable_ids)
                        changed = True
                        break
            if not changed:
                self._warm = None
                self.__echo_res(previous_echo, previous_like)
                return previous_prop, previous_like, previous_criteria
        self._warm = None
        self.__echo_res(previous_echo, previous_like)
        return previous_prop, previous_like, previous_criteria

    # ------------------------------------------------------------------ helpers (reference-private names kept)
    def __echo_res(self, previous_echo, previous_like):
        if self.bootstrap_mode:
            logger.info(f"{self.bootstrap_mode} Proportions: " +
                        ", ".join(["%s:%.4f" % (_id, _p) for _id, _p, in previous_echo.items()]))
            logger.debug("Log-likelihood: %s" % previous_like)
        else:
            logger.info("Proportions: " + ", ".join(["%s:%.4f" % (_id, _p) for _id, _p, in previous_echo.items()]))
            logger.info("Log-likelihood: %s" % previous_like)

    def __drop_zero_variants(self, chosen_ids, representative_props, echo_props, diff_tolerance, indispensable_ids):
        """|prop| < 1e-6 -> dropped without refit (ModelFitMaxLike.py:333-347)."""
        for var_id in list(chosen_ids):
            if var_id in indispensable_ids:
                continue
            if abs(representative_props[var_id] - 0.) < diff_tolerance:
                chosen_ids.remove(var_id)
                for cid_var_id in self.repr_to_merged_variants[var_id]:
                    del representative_props[cid_var_id]
                del echo_props[self.__str_rep_id(var_id)]
                if self.bootstrap_mode:
                    logger.debug("Drop {}".format(self.__str_rep_id(var_id)))
                else:
                    logger.info("Drop {}".format(self.__str_rep_id(var_id)))

    def __summarize_batch(self, id_sets, results, criteria):
        """Batched __compute_like_and_criteria (ModelFitMaxLike.py:448-459): one lane per id set."""
        logger.debug("Generating the likelihood function .. ")
        for ids in id_sets:
            if self.bootstrap_mode:
                logger.debug("Maximizing the likelihood function for {} variants".format(len(ids)))
            else:
                logger.info("Maximizing the likelihood function for {} variants".format(len(ids)))
        _col, _C, sample_size = self._lane
        out = []
        self._batch_x = [dict(zip(sorted(ids), (float(v) for v in res.x))) if res.success else None
                         for ids, res in zip(id_sets, results)]
        for ids, res in zip(id_sets, results):
            info = LogLikeFuncInfo(loglike_func=None, variable_size=self.__variable_size(ids), sample_size=sample_size)
            out.append(self.__summarize_like_and_criteria(res if res.success else False, ids, criteria, info))
        return out

    def __criteria_of_batch(self, id_sets, results, criteria, verbose):
        """AIC / BIC of every lane of a batch (ModelFitMaxLike.py:472-497 per lane).  Raises like the reference when a fit
        failed.  With DEBUG logging on, the reference's per-candidate lines are written as well."""
        _col, _C, sample_size = self._lane
        if criteria not in ("AIC", "BIC"):
            raise Exception("Invalid criterion {}".format(criteria))
        if verbose:
            logger.debug("Generating the likelihood function .. ")
        vals = []
        for ids, res in zip(id_sets, results):
            if verbose or not self.bootstrap_mode:
                msg = "Maximizing the likelihood function for {} variants".format(len(ids))
                if self.bootstrap_mode:
                    logger.debug(msg)
                else:
                    logger.info(msg)
            if not res.success:
                raise Exception("Likelihood maximization failed.")
            this_like = -res.fun
            k = self.__variable_size(ids)
            this_criteria = aic(loglike=this_like, len_param=k) if criteria == "AIC" else bic(loglike=this_like, len_param=k, len_data=sample_size)
            if verbose:
                _use, echo_prop = self.__summarize_run_prop(res, ids)
                logger.debug("Proportions: " + ", ".join(["%s:%.4f" % (_id, _p) for _id, _p, in echo_prop.items()]))
                logger.debug("Log-likelihood: %s" % this_like)
                logger.debug("len_param: %s" % k)
                if criteria == "BIC":
                    logger.debug("len_data: %s" % sample_size)
                logger.debug("%s: %s" % (criteria, this_criteria))
            vals.append(this_criteria)
        return vals

    def _cover_without(self, var_id):
        """cover_all_observed_subpaths(accepted set - {var_id}) from the round's cached cover counts (_set_cover_base):
        uncovered iff some observed read path is held by `var_id` alone."""
        base = self._cover_base
        if base[1]:
            return False
        sole = base[3]
        return not (sole.size and bool(self._cover_cache[3][int(var_id), sole].any()))

    def __variable_size(self, ids):
        # ModelGenerator.py:125-126,175: an empty subset or the full set counts every variant
        n = len(ids)
        if n == 0 or n >= self.num_put_variants and set(ids) == set(range(self.num_put_variants)):
            return self.num_put_variants
        return n

    def __summarize_like_and_criteria(self, success_run, chosen_id_set, criteria, neg_loglike_func_obj):
        if success_run:
            this_like = -success_run.fun
            use_prop, echo_prop = self.__summarize_run_prop(success_run, chosen_id_set)
            logger.debug("Proportions: " + ", ".join(["%s:%.4f" % (_id, _p) for _id, _p, in echo_prop.items()]))
            logger.debug("Log-likelihood: %s" % this_like)
            if criteria == "AIC":
                logger.debug("len_param: %s" % neg_loglike_func_obj.variable_size)
                this_criteria = aic(loglike=this_like, len_param=neg_loglike_func_obj.variable_size)
                logger.debug("%s: %s" % (criteria, this_criteria))
            elif criteria == "BIC":
                logger.debug("len_param: %s" % neg_loglike_func_obj.variable_size)
                logger.debug("len_data: %s" % neg_loglike_func_obj.sample_size)
                this_criteria = bic(loglike=this_like, len_param=neg_loglike_func_obj.variable_size,
                                    len_data=neg_loglike_func_obj.sample_size)
                logger.debug("%s: %s" % (criteria, this_criteria))
            else:
                raise Exception("Invalid criterion {}".format(criteria))
            return use_prop, echo_prop, this_like, this_criteria
        else:
            raise Exception("Likelihood maximization failed.")

    def __summarize_run_prop(self, success_run, within_var_ids):
        """Split a representative's proportion evenly over its unidentifiable group (:499-510)."""
        prop_dict = {}
        representatives = [rep_id for rep_id in sorted(within_var_ids) if rep_id in self.repr_to_merged_variants]
        echo_prop = OrderedDict()
        for go, this_prop in enumerate(success_run.x):
            this_prop = float(this_prop)
            echo_prop[self.__str_rep_id(representatives[go])] = this_prop
            unidentifiable_var_ids = self.repr_to_merged_variants[representatives[go]]
            this_prop /= len(unidentifiable_var_ids)
            for cid_var_id in unidentifiable_var_ids:
                prop_dict[cid_var_id] = this_prop
        use_prop = OrderedDict([(_id, prop_dict[_id]) for _id in sorted(prop_dict)])
        return use_prop, echo_prop

    def __str_rep_id(self, rep_id):
        return "+".join([f"cid_{_cid_var_id}" for _cid_var_id in self.repr_to_merged_variants[rep_id]])

    def get_neg_likelihood_of_var_freq(self, within_variant_ids: set = None, scipy_style=True):
        """LogLikeFuncInfo whose ``loglike_func`` evaluates -LL on the GPU (replaces the
        symengine.lambdify'd callable of ModelFitMaxLike.py:515-558).  scipy_style: one array
        argument, returns a length-1 array; otherwise positional floats."""
        ctx = self._ensure_lane()
        _col, _C, N = self._lane
        if within_variant_ids is None:
            within_variant_ids = set(range(self.num_put_variants))
        ids = sorted(within_variant_ids)
        variable_size = self.__variable_size(set(ids))
        owner = self
        request = self.lane_requests([ids])[0]

        def neg_loglike(*args):
            x = np.asarray(args[0] if scipy_style else args, dtype=np.float64).reshape(-1)
            return ctx.loglik(request, x)

        neg_loglike.tvs_fit = lambda: owner._fit_subsets([set(ids)])[0]
        return LogLikeFuncInfo(loglike_func=neg_loglike, variable_size=variable_size, sample_size=N)

    def update_observed_sp_ids(self):
        if getattr(self, "_array_form", False):
            return                                          # fixed at construction (from_lane)
        self.observed_sbp_id_set = set()
        for go_sp, (this_sub_path, this_sub_path_info) in enumerate(self.all_sub_paths.items()):
            if this_sub_path_info.mapped_records:
                self.observed_sbp_id_set.add(go_sp)
            else:
                logger.trace("Drop subpath without observation: {}: {}".format(go_sp, this_sub_path))
        self._cover_cache = None

    def _cover_key(self):
        """What the cover tables were built from: the reference keeps observed_sbp_id_set and sbp_to_sbp_id as public,
        replaceable attributes, so the cache is keyed on their identity and size (a same-sized replacement rebuilds)."""
        obs, sbp = self.observed_sbp_id_set, self.sbp_to_sbp_id
        return (id(obs), len(obs), id(sbp), len(sbp) if sbp is not None else -1)

    def _cover_tables(self):
        """Matrix form of the tables cover_all_observed_subpaths walks (ModelFitMaxLike.py:568-580):
        contains[s, v] = sub-path id s occurs in variant v (through variant_subpath_counters and sbp_to_sbp_id),
        observed[s] = s is in observed_sbp_id_set.  Same sets, evaluated with numpy instead of Python set unions
        (the reference spends O(sub-paths) dict operations per variant per candidate)."""
        cache = getattr(self, "_cover_cache", None)
        key = self._cover_key()
        if cache is not None and cache[0] == key:
            return cache[1], cache[2]
        n_sbp = (max(self.sbp_to_sbp_id.values()) + 1) if self.sbp_to_sbp_id else 0
        n_sbp = max(n_sbp, (max(self.observed_sbp_id_set) + 1) if self.observed_sbp_id_set else 0)
        # variant -> sub-path membership is the same for every replicate of a run: built once per context in the key
        # space of the sub paths, then re-indexed by this model's own sbp_to_sbp_id
        gid, contains_g = self._global_cover_table()
        contains = np.zeros((n_sbp, self.num_put_variants), dtype=bool)
        if self.sbp_to_sbp_id:
            keys = list(self.sbp_to_sbp_id.keys())
            local = np.fromiter(self.sbp_to_sbp_id.values(), dtype=np.int64, count=len(keys))
            glob = np.fromiter((gid.get(k, -1) for k in keys), dtype=np.int64, count=len(keys))
            ok = glob >= 0
            if np.unique(local[ok]).size == int(ok.sum()):            # ids are distinct: a plain row assignment
                contains[local[ok]] = contains_g[glob[ok]]
            else:
                np.logical_or.at(contains, local[ok], contains_g[glob[ok]])
        observed = np.zeros(n_sbp, dtype=bool)
        observed[list(self.observed_sbp_id_set)] = True
        # variant-major copy: a candidate set is a gather of contiguous rows
        self._cover_cache = (key, contains, observed, np.ascontiguousarray(contains.T))
        self._cover_base = None
        return contains, observed

    def _global_cover_table(self):
        """(sub-path key -> row, contains[row, v]) for this run's variants; cached on the LaneContext when there is one."""
        holder = self._context if self._context is not None else self
        key = (id(self.variant_subpath_counters), id(self.variant_paths))
        cached = getattr(holder, "_cover_global", None)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        gid = {}
        rows, cols = [], []
        for go_var in range(self.num_put_variants):
            for sp in self.variant_subpath_counters[self.variant_paths[go_var]]:
                r = gid.get(sp)
                if r is None:
                    r = gid[sp] = len(gid)
                rows.append(r)
                cols.append(go_var)
        contains_g = np.zeros((len(gid), self.num_put_variants), dtype=bool)
        contains_g[rows, cols] = True
        holder._cover_global = (key, gid, contains_g)
        return gid, contains_g

    def cover_all_observed_subpaths(self, variant_ids):
        """Every observed read path must occur in >=1 variant of the set (ModelFitMaxLike.py:568-580)."""
        if not self.observed_sbp_id_set:
            self.update_observed_sp_ids()
        contains, observed = self._cover_tables()
        by_variant = self._cover_cache[3]                       # [V, sub paths]
        ids = frozenset(int(v) for v in variant_ids)
        if not ids:
            return not observed.any()
        base = getattr(self, "_cover_base", None)
        if base is not None and ids <= base[0] and len(base[0]) - len(ids) <= 1:
            # leave-one-out of the set whose per-sub-path cover counts are cached (the candidates of one selection
            # round): uncovered iff some observed sub path is held by the dropped variant alone
            if base[1]:
                return False
            dropped = base[0] - ids
            if not dropped:
                return True
            return not bool(np.any(by_variant[next(iter(dropped))] & base[2]))
        return bool(by_variant[sorted(ids)].any(axis=0)[observed].all())

    def _set_cover_base(self, variant_ids):
        """Cache the cover counts of `variant_ids` so that its leave-one-out subsets are tested in O(sub paths)."""
        if not self.observed_sbp_id_set:
            self.update_observed_sp_ids()
        self._cover_tables()
        ids = frozenset(int(v) for v in variant_ids)
        observed = self._cover_cache[2]
        by_variant = self._cover_cache[3]
        prev = self._cover_base
        if prev is not None and len(prev) > 4 and prev[5] is by_variant and ids <= prev[0] and len(prev[0]) - len(ids) <= 8:
            counts = prev[4].copy()                            # the previous round's set minus the variants dropped since
            for v in prev[0] - ids:
                counts -= by_variant[v]
        else:
            counts = by_variant[sorted(ids)].sum(axis=0, dtype=np.int32)
        sole = (counts == 1) & observed
        self._cover_base = (ids, bool(np.any((counts == 0) & observed)), sole, np.nonzero(sole)[0], counts, by_variant)
