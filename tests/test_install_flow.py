"""``traversome_b200.install()``: the swaps of INTEGRATION.md performed on an orchestrator module, and the reference's
ML + bootstrap call sequence (traversome.py:383-392, 501-617) run through them.

The orchestrator here is tests/flow_standin.py (reference-free, so that the GPU variant runs on the B200 box, where
/root/reference does not exist); tests/test_dropin_reference_flow.py runs the REAL reference module through install()
in the build container.  CPU variant: oracle-backed engine; GPU variant (-m gpu): libtvs.so."""
import numpy as np
import pytest

import traversome_b200
from tests import flow_standin
from tests.helpers import OracleEngine, fit_context, model_from_tables
from traversome_b200 import _cabi, lanes


def _flow(plastome, n_rep):
    L = plastome["variant_sizes"]
    models = [model_from_tables(L, m["bins_list"]) for m in plastome["models"][:n_rep + 1]]
    kw0 = fit_context(models[0], plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
    reps = []
    for m in models[1:]:
        kw = fit_context(m, plastome["be_unidentifiable_to"], plastome["repr_to_merged_variants"])
        reps.append((m.all_sub_paths, m.bins_list, kw["sbp_to_sbp_id"]))
    return flow_standin.Traversome(L, kw0["variant_paths"], kw0["variant_subpath_counters"], kw0["repr_to_merged_variants"],
                                   kw0["be_unidentifiable_to"], models[0], kw0["sbp_to_sbp_id"], reps, bootstrap=n_rep)


def _check(trv, plastome, n_rep, tol):
    assert abs(trv.variant_proportions[0] - 23 / 35.) < tol and abs(trv.variant_proportions[1] - 12 / 35.) < tol
    assert abs(trv.res_loglike - plastome["final"]["res_loglike"]) <= 1e-9 * abs(trv.res_loglike)
    assert len(trv.variant_proportions_reps) == n_rep and trv.bs_eligible
    assert trv.n_serial_fits == 0                       # the serial loop was replaced
    ctx = trv.max_like_fit._ensure_lane()
    assert ctx.n_batches <= 8                           # whole-data selection + a handful of lock-step rounds, not 100+ fits
    for v_prop, fit in zip(trv.variant_proportions_reps, plastome["fits"][1:]):
        assert sorted(v_prop) == sorted(c for c, _ in fit["prop"])
        assert all(abs(v_prop[c] - p) < 6e-4 for c, p in fit["prop"])      # the reference's SLSQP stops within ~3e-4
    if n_rep == 100:
        support = [[list(map(int, k)), float(v["support"])] for k, v in trv.vp_unique_results.items()]
        assert support == plastome["final"]["support"]


def test_install_swaps_and_restores():
    saved = (flow_standin.ModelFitMaxLike, flow_standin.Traversome.do_subsampling, flow_standin.Traversome.generate_read_paths)
    mod = traversome_b200.install(flow_standin)
    try:
        assert mod is flow_standin
        assert flow_standin.ModelFitMaxLike is traversome_b200.ModelFitMaxLike.ModelFitMaxLike
        assert flow_standin.VariantSubPathsGenerator is traversome_b200.subpaths.VariantSubPathsGenerator
        assert flow_standin.Traversome.do_subsampling is not saved[1]
        assert flow_standin.Traversome.generate_read_paths is not saved[2]
        assert traversome_b200.install(flow_standin) is flow_standin           # idempotent
    finally:
        traversome_b200.uninstall()
    assert (flow_standin.ModelFitMaxLike, flow_standin.Traversome.do_subsampling, flow_standin.Traversome.generate_read_paths) == saved


def test_installed_flow_on_the_cpu_stand_in(plastome, monkeypatch):
    monkeypatch.setattr(lanes._cabi, "Engine", OracleEngine)
    traversome_b200.install(flow_standin)
    try:
        trv = _flow(plastome, 12)
        trv.run()
        _check(trv, plastome, 12, 1e-7)
    finally:
        traversome_b200.uninstall()


@pytest.mark.gpu
def test_installed_flow_on_the_gpu_engine(plastome, lib_built):
    """The same flow -- whole-data selection + all 100 bootstrap replicates of the golden run -- on libtvs.so."""
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    traversome_b200.install(flow_standin)
    try:
        trv = _flow(plastome, 100)
        trv.run()
        _check(trv, plastome, 100, 1e-7)
    finally:
        traversome_b200.uninstall()
