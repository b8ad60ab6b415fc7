"""``PathMultinomialModel`` -- the likelihood model object of the ML-fit boundary.

Mirror of ``traversome/ModelGenerator.py:8-15,110-178``.  ``get_like_formula`` stays backend
agnostic (symbols + a symbolic ``log``, pytensor variables + ``tt.log`` for the untouched Bayesian
path -- ModelFitBayesian.py:66-69 -- or plain floats + ``math.log``) and keeps the reference's
term order, so that evaluated with floats it returns the reference's number bit for bit.  The
GPU path does not call it: it consumes the integer tables through ``traversome_b200.csr``.
"""
from typing import Set

from .utils import LogLikeFormulaInfo


class PathMultinomialModel(object):
    def __init__(self, variant_sizes, variant_topos, bins_list, all_sub_paths):
        self.variant_sizes = variant_sizes
        self.variant_topos = variant_topos
        self.num_put_variants = len(variant_sizes)
        self.bins_list = bins_list
        self.all_sub_paths = all_sub_paths   # only used for assessing read-path coverage
        self.sample_size = None

    def prune_bins(self):
        """Drop bins whose expected number of start points is < 1, then empty ranges.

        Side effect on the model, exactly as the reference does inside get_like_formula
        (ModelGenerator.py:138-151)."""
        kept_ranges = []
        for bins in self.bins_list:
            bins.rp_bins[:] = [rp_bin for rp_bin in bins.rp_bins if not (rp_bin.num_possible_X < 1)]
            if bins.rp_bins:
                kept_ranges.append(bins)
        self.bins_list[:] = kept_ranges

    def get_like_formula(self, variant_percents, log_func, within_variant_ids: Set = None):
        """LL = sum_bins n_b * log( X_b * sum_v p_v f_bv / sum_v p_v L_v ), restricted to
        ``within_variant_ids`` (ModelGenerator.py:110-178).  Returns LogLikeFormulaInfo."""
        subset = within_variant_ids
        if not subset or subset == set(range(self.num_put_variants)):
            subset = None
        total_length = 0
        for variant_id, variant_len in enumerate(self.variant_sizes):
            if subset is None or variant_id in subset:
                total_length += variant_percents[variant_id] * float(variant_len)
        self.prune_bins()
        loglike = 0
        n_records = 0
        for bins in self.bins_list:
            for rp_bin in bins.rp_bins:
                weight = 0
                for variant_id, occurrences in rp_bin.from_variants.items():
                    if subset is None or variant_id in subset:
                        weight += variant_percents[variant_id] * occurrences
                loglike += log_func(weight * rp_bin.num_possible_X / total_length) * rp_bin.num_matched
                n_records += rp_bin.num_matched
        self.sample_size = n_records
        n_param = len(subset) if subset else self.num_put_variants
        return LogLikeFormulaInfo(loglike, n_param, n_records)
