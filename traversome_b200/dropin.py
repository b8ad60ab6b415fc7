"""``traversome_b200.install()`` -- run the reference's ``traversome thorough`` flow on the B200 path.

The reference (JianjunJin/Traversome v0.1.7.1) has no plugin interface: its orchestrator imports the fitter by name
(``traversome/traversome.py:22``).  ``install(traversome.traversome)`` performs, on the imported module object, the
swaps INTEGRATION.md describes, so that the unmodified CLI (``traversome/__main__.py:238-653`` -> ``Traversome.run``,
``traversome.py:145-439``) fits on the GPU:

    1. ``ModelFitMaxLike``              <- ``traversome_b200.ModelFitMaxLike.ModelFitMaxLike``      (traversome.py:22, :1850, :1882)
    2. ``VariantSubPathsGenerator``     <- ``traversome_b200.subpaths.VariantSubPathsGenerator``    (traversome.py:306-311)
    3. ``Traversome.generate_read_paths`` <- array form (``traversome_b200.readpaths``)            (traversome.py:869-966)
    4. ``Traversome.do_subsampling``    <- all replicates in lock step on one ``LaneContext``       (traversome.py:501-617)

Nothing is copied from the reference and the reference tree itself is not modified; ``uninstall()`` restores the
module.  ``python -m traversome_b200 thorough ...`` (``__main__.py`` of this package) installs the swaps and hands over
to the reference's own CLI app.  There is no CPU fallback: without ``libtvs.so`` / a CUDA device the first fit raises.
"""
from loguru import logger

from . import bootstrap, readpaths
from .ModelFitMaxLike import ModelFitMaxLike
from .ModelGenerator import PathMultinomialModel
from .subpaths import VariantSubPathsGenerator
from .utils import Criterion

_SAVED = {}


def _generate_read_paths(self, graph_alignment, filter_by_graph=True, min_alignment_counts=1):
    """Traversome.generate_read_paths (traversome.py:869-966) in array form."""
    table = readpaths.generate_read_paths(self.graph, graph_alignment.raw_records, filter_by_graph=filter_by_graph,
                                          min_alignment_counts=min_alignment_counts)
    readpaths.install_on(self, table)


def _do_subsampling(self):
    """Traversome.do_subsampling (traversome.py:501-617): the replicate models are built by the reference's own
    resampling / bin-statistics methods (:515-539, untouched, same random draws), then ALL replicates run their backward
    model selection in lock step -- every round is one batch of GPU lanes -- and the reference's tally follows."""
    self._prepare_for_sampling()
    n_replicate = self.kwargs.get("bootstrap", 0) if self.kwargs.get("bootstrap", 0) else self.kwargs.get("jackknife", 0)
    threshold = self.kwargs.get("bs_threshold", 0.95)
    criterion = self.kwargs.get("model_criterion", Criterion.BIC)
    model_cls = _SAVED.get("PathMultinomialModel_of_module", PathMultinomialModel)
    reps = []
    while len(reps) < n_replicate:
        logger.debug(f"Sampling {len(reps) + 1} --------")
        if self.kwargs.get("bootstrap", 0):
            sampled_sub_paths, rec_ids, lens = self.sample_sub_paths(bootstrap_size=self.num_valid_records, masking=self.read_paths_masked)
        else:
            sampled_sub_paths, rec_ids, lens = self.sample_sub_paths(
                jackknife_size=int(self.num_valid_records / float(n_replicate)), masking=self.read_paths_masked)
        if not sampled_sub_paths:                       # traversome.py:528-529: draw again
            continue
        bins_list = self.generate_multinomial_bin_stats(all_sub_paths=sampled_sub_paths, rec_id_sorted_by_len=rec_ids,
                                                        align_len_at_path_sorted=lens)
        reps.append((model_cls(variant_sizes=self.variant_sizes, variant_topos=self.variant_topos, bins_list=bins_list,
                               all_sub_paths=sampled_sub_paths),
                     self.update_sp_to_sp_id_dict(sampled_sub_paths)))
    summary = bootstrap.do_subsampling(
        reps, dict(variant_paths=self.variant_paths, variant_subpath_counters=self.variant_subpath_counters,
                   repr_to_merged_variants=self.repr_to_merged_variants, be_unidentifiable_to=self.be_unidentifiable_to,
                   loglevel=self.loglevel),
        full_fit=self.max_like_fit, raw_result=(self.variant_proportions, self.res_loglike, self.res_criterion),
        criterion=criterion, n_replicate=n_replicate, threshold=threshold, user_fixed_ids=self.user_variant_fixed_ids)
    self.variant_proportions_reps = summary.variant_proportions_reps
    self.bs_eligible = summary.bs_eligible
    self.vp_unique_results, self.vp_unique_results_sorted = summary.vp_unique_results, summary.vp_unique_results_sorted
    self._cid_sorter = summary.cid_sorter
    if summary.variant_proportions_best is not None:
        self.variant_proportions_best = summary.variant_proportions_best


def install(module=None, batch_bootstrap=True, read_paths=True, subpaths=True, parser=True):
    """Swap this package's classes into the reference's orchestrator module (default: ``traversome.traversome``,
    imported here).  Returns the module.  Idempotent."""
    if module is None:
        import importlib
        module = importlib.import_module("traversome.traversome")
    if _SAVED.get("module") is module:
        return module
    if _SAVED.get("module") is not None:
        uninstall()
    trv_cls = module.Traversome
    _SAVED.update(module=module, attrs={name: getattr(module, name) for name in ("ModelFitMaxLike", "VariantSubPathsGenerator")
                                        if hasattr(module, name)},
                  methods={name: trv_cls.__dict__[name] for name in ("generate_read_paths", "do_subsampling") if name in trv_cls.__dict__})
    if hasattr(module, "PathMultinomialModel"):
        _SAVED["PathMultinomialModel_of_module"] = module.PathMultinomialModel
    module.ModelFitMaxLike = ModelFitMaxLike                                   # swap 1
    if subpaths:
        module.VariantSubPathsGenerator = VariantSubPathsGenerator              # swap 2
    if read_paths:
        trv_cls.generate_read_paths = _generate_read_paths                      # swap 3
    if batch_bootstrap:
        trv_cls.do_subsampling = _do_subsampling                                # swap 4
    if parser and hasattr(module, "GraphAlignRecords"):
        from . import gaf
        gaf.install_on_records(module.GraphAlignRecords)                        # swap 5: alignment file -> raw_records
    return module


def uninstall():
    """Put the reference's own objects back."""
    module = _SAVED.pop("module", None)
    if module is None:
        return
    from . import gaf
    gaf.uninstall_from_records()
    trv_cls = module.Traversome
    for name in ("ModelFitMaxLike", "VariantSubPathsGenerator"):
        if name in _SAVED.get("attrs", {}):
            setattr(module, name, _SAVED["attrs"][name])
        elif hasattr(module, name):
            delattr(module, name)
    for name in ("generate_read_paths", "do_subsampling"):
        if name in _SAVED.get("methods", {}):
            setattr(trv_cls, name, _SAVED["methods"][name])
        elif name in trv_cls.__dict__:
            delattr(trv_cls, name)
    _SAVED.clear()


def main(argv=None):
    """``python -m traversome_b200 <reference CLI arguments>``: install the swaps, then run the reference's typer app
    (``traversome.__main__:app``, the console script ``traversome``)."""
    import importlib
    import sys
    install()
    app = importlib.import_module("traversome.__main__").app
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    return app()
