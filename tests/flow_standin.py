"""TEST INFRASTRUCTURE: a reference-free stand-in for the ML part of the reference's orchestrator module
(``traversome/traversome.py``), so that ``traversome_b200.install()`` can be exercised on the GPU box, where
``/root/reference`` does not exist.

It has exactly the module surface ``install()`` touches -- the names ``ModelFitMaxLike``, ``VariantSubPathsGenerator``,
``PathMultinomialModel`` and a class ``Traversome`` with ``generate_read_paths`` / ``do_subsampling`` -- and its
``Traversome`` drives the same call sequence as the reference's ``run()`` for this path: whole-data backward selection
(traversome.py:383-392 -> :1864-1896), then ``do_subsampling`` (:501-617).  The per-replicate inputs that the reference
derives from its record pool (``sample_sub_paths`` :1636-1699, ``generate_multinomial_bin_stats`` :1305-1384) are replayed
from the golden run (tests/golden/plastome_cfg1.json.gz holds the bins of the 100 replicates the reference drew with
--rs 12345).  The UNINSTALLED ``do_subsampling`` below is the serial loop, one replicate after the other."""
from traversome_b200.ModelGenerator import PathMultinomialModel  # noqa: F401  (module attribute read by install())
from traversome_b200.utils import Criterion

ModelFitMaxLike = None               # the reference imports its fitter here (traversome.py:22); install() swaps it
VariantSubPathsGenerator = None


class Traversome(object):
    def __init__(self, variant_sizes, variant_paths, variant_subpath_counters, repr_to_merged_variants, be_unidentifiable_to,
                 full_model, full_sbp_to_sbp_id, replicate_models, bootstrap=100, bs_threshold=0.95, model_criterion=Criterion.BIC):
        self.kwargs = {"bootstrap": bootstrap, "jackknife": 0, "bs_threshold": bs_threshold, "model_criterion": model_criterion}
        self.variant_sizes, self.variant_topos = variant_sizes, [True] * len(variant_sizes)
        self.variant_paths, self.variant_subpath_counters = variant_paths, variant_subpath_counters
        self.repr_to_merged_variants, self.be_unidentifiable_to = repr_to_merged_variants, be_unidentifiable_to
        self.model, self._full_sbp = full_model, full_sbp_to_sbp_id
        self._replicates = list(replicate_models)              # [(all_sub_paths, bins_list, sbp_to_sbp_id)]
        self._drawn = 0
        self.loglevel, self.num_processes, self.user_variant_fixed_ids = "ERROR", 1, None
        self.num_valid_records, self.read_paths_masked = 0, set()
        self.max_like_fit = None
        self.variant_proportions = self.res_loglike = self.res_criterion = None
        self.variant_proportions_reps, self.bs_eligible = [], True
        self.vp_unique_results, self.vp_unique_results_sorted, self.variant_proportions_best = {}, [], None
        self.n_serial_fits = 0

    # ---- the pieces of the reference the fit path calls (replayed from the golden run)
    def _prepare_for_sampling(self):
        self._drawn = 0

    def sample_sub_paths(self, bootstrap_size=None, jackknife_size=None, masking=None):
        sub_paths, bins_list, sbp = self._replicates[self._drawn]
        self._drawn += 1
        return sub_paths, ("bins", bins_list), ("sbp", sbp)

    def generate_multinomial_bin_stats(self, all_sub_paths, rec_id_sorted_by_len, align_len_at_path_sorted):
        self._sbp_of_last = align_len_at_path_sorted[1]
        return rec_id_sorted_by_len[1]

    def update_sp_to_sp_id_dict(self, sub_paths):
        return self._sbp_of_last

    def generate_read_paths(self, graph_alignment, filter_by_graph=True, min_alignment_counts=1):
        raise NotImplementedError("not part of this stand-in")

    # ---- traversome.py:1864-1896
    def fit_model_using_reverse_model_selection(self, model, sbp_to_sbp_id, criterion=Criterion.BIC, chosen_ids=None,
                                                init_self_max_like=True, bootstrap_str=""):
        max_like_fit = ModelFitMaxLike(model=model, variant_paths=self.variant_paths, variant_subpath_counters=self.variant_subpath_counters,
                                       sbp_to_sbp_id=sbp_to_sbp_id, repr_to_merged_variants=self.repr_to_merged_variants,
                                       be_unidentifiable_to=self.be_unidentifiable_to, loglevel=self.loglevel, bootstrap_mode=bootstrap_str)
        if init_self_max_like:
            self.max_like_fit = max_like_fit
        return max_like_fit.reverse_model_selection(n_proc=self.num_processes, criterion=criterion, chosen_ids=chosen_ids,
                                                    user_fixed_ids=self.user_variant_fixed_ids)

    # ---- traversome.py:501-556, the serial loop (what install() replaces)
    def do_subsampling(self):
        self._prepare_for_sampling()
        n_replicate = self.kwargs["bootstrap"]
        self.variant_proportions_reps = []
        for go_bs in range(n_replicate):
            sub_paths, rec, lens = self.sample_sub_paths(bootstrap_size=self.num_valid_records, masking=self.read_paths_masked)
            bins_list = self.generate_multinomial_bin_stats(sub_paths, rec, lens)
            model = PathMultinomialModel(variant_sizes=self.variant_sizes, variant_topos=self.variant_topos, bins_list=bins_list,
                                         all_sub_paths=sub_paths)
            v_prop, *_ = self.fit_model_using_reverse_model_selection(model=model, sbp_to_sbp_id=self.update_sp_to_sp_id_dict(sub_paths),
                                                                      criterion=self.kwargs["model_criterion"], init_self_max_like=False,
                                                                      bootstrap_str="BS%d" % (go_bs + 1))
            self.variant_proportions_reps.append(v_prop)
            self.n_serial_fits += 1

    # ---- traversome.py:383-392 then :501
    def run(self):
        self.variant_proportions, self.res_loglike, self.res_criterion = self.fit_model_using_reverse_model_selection(
            model=self.model, sbp_to_sbp_id=self._full_sbp, criterion=self.kwargs["model_criterion"])
        self.do_subsampling()
