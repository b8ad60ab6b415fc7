"""CPU, build container only: the UNMODIFIED reference driver (`traversome.traversome.Traversome.run`, the `thorough`
flow of BASELINE configs[0]) after `traversome_b200.install(traversome.traversome)` -- the shipped entry point that
performs the swaps of INTEGRATION.md:

    traversome.traversome.ModelFitMaxLike          <- traversome_b200.ModelFitMaxLike.ModelFitMaxLike
    traversome.traversome.VariantSubPathsGenerator <- traversome_b200.subpaths.VariantSubPathsGenerator
    Traversome.generate_read_paths                 <- traversome_b200.readpaths.generate_read_paths
    Traversome.do_subsampling                      <- all replicates in lock step (traversome_b200.bootstrap)

There is no GPU here, so the two device calls are replaced by their test stand-ins (the oracle-backed engine of
tests/helpers.py and the contract emulation of tests/subpath_helpers.py); everything above the C ABI -- constructor
arguments, call order, return shapes, log-free operation, the output tables the reference then writes -- is the real
code on both sides.  Skipped where /root/reference does not exist (the GPU box)."""
import gzip
import os
import shutil
import sys

import numpy as np
import pytest

REFERENCE = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "traversome")),
                                reason="the reference tree is only present in the build container")


def _gunzip(src, dst):
    with gzip.open(src, "rb") as i, open(dst, "wb") as o:
        shutil.copyfileobj(i, o)


def test_reference_thorough_flow_with_the_swapped_classes(plastome, tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, REFERENCE)
    import symengine_standin
    symengine_standin.install()                      # only so that `import traversome.traversome` succeeds
    import traversome.traversome as tvm
    from traversome.utils import Criterion
    from loguru import logger

    from tests.helpers import OracleEngine
    from tests.test_subpaths_oracle import EmulatedGenerator
    from traversome_b200 import lanes
    from traversome_b200.ModelFitMaxLike import ModelFitMaxLike as Mirror

    created = {"fitters": 0, "generators": 0}

    class CountingMirror(Mirror):
        def __init__(self, *a, **k):
            created["fitters"] += 1
            super().__init__(*a, **k)

    class CountingGenerator(EmulatedGenerator):
        def __init__(self, *a, **k):
            created["generators"] += 1
            super().__init__(*a, **k)

    import traversome_b200
    from traversome_b200 import dropin
    created["read_path_calls"] = 0
    installed_grp = dropin._generate_read_paths

    def generate_read_paths(self, graph_alignment, filter_by_graph=True, min_alignment_counts=1):
        created["read_path_calls"] += 1
        assert not self.read_paths                                         # the reference starts from an empty table
        installed_grp(self, graph_alignment, filter_by_graph, min_alignment_counts)

    monkeypatch.setattr(dropin, "_generate_read_paths", generate_read_paths)
    traversome_b200.install(tvm)                                           # the four swaps (INTEGRATION.md sections 1, 2, 4, 6)
    monkeypatch.setattr(lanes._cabi, "Engine", OracleEngine)               # no GPU in this container: device calls -> stand-ins
    monkeypatch.setattr(tvm, "ModelFitMaxLike", CountingMirror)            # the installed class, counted
    monkeypatch.setattr(tvm, "VariantSubPathsGenerator", CountingGenerator)  # contract emulation of tvs_subpaths_count

    gfa, gaf = str(tmp_path / "plastome.gfa"), str(tmp_path / "sim.gaf")
    _gunzip(os.path.join(GOLDEN, "plastome.gfa.gz"), gfa)
    _gunzip(os.path.join(GOLDEN, "plastome_sim.gaf.gz"), gaf)
    var = os.path.join(GOLDEN, "plastome_variants.txt")
    outdir = str(tmp_path / "out")
    os.makedirs(outdir)
    np.random.seed(20240607)
    try:
        trv = tvm.Traversome(
            graph=gfa, alignment=gaf, reads_file=None, outdir=outdir, var_fixed=None, var_candidate=var,
            model_criterion=Criterion.BIC, bootstrap=100, bs_threshold=0.95, jackknife=0,
            max_valid_search=0, num_processes=1, force_circular=True, uni_chromosome=False,
            n_generations=0, random_seed=12345, loglevel="ERROR", keep_temp=True,
            min_read_identity_cutoff=0.992, min_record_identity_cutoff=0.99, min_alignment_len_cutoff=5000,
            min_alignment_counts="auto", graph_component_selection=0, ignore_conflicts=True,
            keep_graph_redundancy=False)
        trv.run()
    finally:
        logger.remove()
        traversome_b200.uninstall()

    assert created["generators"] == 1 and created["fitters"] >= 1          # the whole-data fitter comes from the swapped name
    assert created["read_path_calls"] >= 1
    ctx = trv.max_like_fit._ensure_lane()
    assert ctx.n_batches <= 8 and ctx.n_lanes_fitted >= 101                # 100 replicates in a handful of lock-step batches
    # what the reference's own objects hold afterwards
    assert [int(s) for s in trv.variant_sizes] == plastome["variant_sizes"]
    assert [[int(k), int(v)] for k, v in trv.be_unidentifiable_to.items()] == plastome["be_unidentifiable_to"]
    got_rp = [{"path": [[str(n), bool(e)] for n, e in p], "inner_len": int(i.inner_len), "full_len": int(i.full_len),
               "from_variants": [[int(k), int(c)] for k, c in i.from_variants.items()],
               "n_records": len(i.mapped_records)} for p, i in trv.all_sub_paths.items()]
    want_rp = [{k: rp[k] for k in ("path", "inner_len", "full_len", "from_variants", "n_records")}
               for rp in plastome["read_paths"]]
    assert got_rp == want_rp                                               # A, in the reference's row order
    assert abs(trv.variant_proportions[0] - 23 / 35.) < 1e-7 and abs(trv.variant_proportions[1] - 12 / 35.) < 1e-7
    assert abs(trv.res_loglike - plastome["final"]["res_loglike"]) <= 1e-9 * abs(trv.res_loglike)
    assert abs(trv.res_criterion - plastome["final"]["res_criterion"]) <= 1e-9 * abs(trv.res_criterion)
    assert trv.bs_eligible and len(trv.variant_proportions_reps) == 100
    support = [[list(map(int, k)), float(v["support"])] for k, v in trv.vp_unique_results.items()]
    assert support == plastome["final"]["support"]

    # the tables the reference wrote from those results
    want = plastome["output_tables"]
    for name in ("variants.info.tab", "readpath.information.tab"):
        assert open(os.path.join(outdir, name)).read() == want[name], name
    mine = open(os.path.join(outdir, "bootstraps.replicates.tab")).read().split("\n")
    ref = want["bootstraps.replicates.tab"].split("\n")
    assert mine[0] == ref[0] and len(mine) == len(ref)
    for a, b in zip(mine[1:], ref[1:]):
        if not b:
            continue
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[0] == fb[0]
        # same replicates (the draws replay with --rs 12345); the reference's SLSQP stops within ~3e-4 of the optimum
        assert all(abs(float(x) - float(y)) <= 6e-4 for x, y in zip(fa[1:], fb[1:])), (a, b)
    a, b = open(os.path.join(outdir, "final.result.tab")).read().split("\n"), want["final.result.tab"].split("\n")
    assert a[0] == b[0] and a[1].split("\t")[:2] == b[1].split("\t")[:2]
    assert all(abs(float(x) - float(y)) <= 1.01e-4 for x, y in zip(a[1].split("\t")[4:6], b[1].split("\t")[4:6]))
