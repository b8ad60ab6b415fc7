#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generates the committed fixtures under tests/golden/.

Runs the UNMODIFIED reference (`/root/reference/traversome`, v0.1.7.1) inside this
container and records its integer tables and its numeric answers:

  plastome_cfg1.json   BASELINE.json configs[0]: the reference's own `simulate` +
                       `thorough --topo circular --vc variants.txt -N 0 --bs 100` flow on
                       a 3-contig plastome graph (SURVEY.md section 8c recipe).
  synthetic_small.json reference model objects built directly from seeded arrays
                       (Bins/BinInfo/SubPathInfo/PathMultinomialModel), fitted through
                       the reference's ModelFitMaxLike (point_estimate and
                       reverse_model_selection).

`symengine` is absent from this image, so the reference's `import symengine` is
served by `oracle/symengine_standin.py` (sympy-backed).  Everything that passes
through it is labelled `"via": "reference + sympy stand-in for symengine"`.  The
integer tables (from_variants, num_matched, variant_sizes, record->path maps) and
the known-answer log-likelihoods (`get_like_formula(floats, math.log)`,
ModelGenerator.py:110-178) do NOT pass through the stand-in: they are the
reference's own arithmetic.

/root/reference does not exist on the GPU box, so this script is run HERE, once,
and its outputs are committed.  Usage:  python oracle/make_golden.py [plastome|synthetic|all]
"""
import gzip
import json
import math
import os
import random
import shutil
import sys
import tempfile
from collections import OrderedDict
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = "/root/reference"
VIA = "reference + sympy stand-in for symengine"


def _import_reference():
    sys.path.insert(0, HERE)
    sys.path.insert(0, REFERENCE)
    import symengine_standin
    mod = symengine_standin.install()
    import traversome  # noqa: F401
    return getattr(mod, "_TVS_STANDIN", False)


# --------------------------------------------------------------------------- dumps
def dump_bins_list(bins_list):
    """Integer/float tables of a reference bins_list (utils.py:193-226)."""
    out = []
    for bins in bins_list:
        out.append({
            "min_len": int(bins.min_len), "max_len": int(bins.max_len),
            "min_id": int(bins.min_id), "max_id": int(bins.max_id),
            "rp_bins": [{
                "from_variants": [[int(k), int(v)] for k, v in b.from_variants.items()],
                "num_matched": int(b.num_matched),
                # float.hex keeps the reference's fp64 value bit-exact in JSON
                "num_possible_X": float(b.num_possible_X).hex(),
            } for b in bins.rp_bins],
        })
    return out


def _log0(v):
    """math.log that answers -inf at 0 like the LLVM-compiled objective does (ModelFitMaxLike.py:16)."""
    return math.log(v) if v > 0 else float("-inf")


def known_answer_ll(model, points, subsets):
    """LL known answers straight from the reference (ModelGenerator.py:110-178).
    `model` is deep-copied first because get_like_formula prunes bins in place."""
    res = []
    for theta, subset in zip(points, subsets):
        m = deepcopy(model)
        vp = [float(t) for t in theta]
        info = m.get_like_formula(vp, _log0, set(subset) if subset is not None else None)
        res.append({"theta": [float(t).hex() for t in theta],
                    "subset": sorted(subset) if subset is not None else None,
                    "loglike": float(info.loglike_expression).hex(),
                    "variable_size": int(info.variable_size),
                    "sample_size": int(info.sample_size)})
    return res


# ------------------------------------------------------------------------ plastome
def write_plastome(tmp):
    rnd = random.Random(20240607)
    lens = OrderedDict([("LSC", 86000), ("IR", 26000), ("SSC", 18000)])
    gfa = os.path.join(tmp, "plastome.gfa")
    with open(gfa, "w") as h:
        h.write("H\tVN:Z:1.0\n")
        for name, ln in lens.items():
            h.write("S\t%s\t%s\n" % (name, "".join(rnd.choice("ACGT") for _ in range(ln))))
        for a, ao, b, bo in (("LSC", "+", "IR", "+"), ("LSC", "-", "IR", "+"),
                             ("IR", "+", "SSC", "+"), ("IR", "+", "SSC", "-")):
            h.write("L\t%s\t%s\t%s\t%s\t2M\n" % (a, ao, b, bo))
    var = os.path.join(tmp, "variants.txt")
    with open(var, "w") as h:
        h.write("LSC+,IR+,SSC+,IR-(circular)\n")
        h.write("LSC+,IR+,SSC-,IR-(circular)\n")
    return gfa, var


def make_plastome():
    from traversome.Assembly import Assembly
    from traversome.Simulator import SimpleSimulator
    from traversome.utils import user_paths_reader, Criterion
    import traversome.traversome as tvm
    import traversome.ModelFitMaxLike as mfl
    from traversome.ModelGenerator import PathMultinomialModel

    tmp = tempfile.mkdtemp(prefix="tvs_golden_")
    gfa, var = write_plastome(tmp)
    gaf = os.path.join(tmp, "sim.gaf")
    SimpleSimulator(graph_obj=Assembly(graph_file=gfa), variants=user_paths_reader(var),
                    variant_proportions=[0.6, 0.4], length_distribution="hifi",
                    data_size=30000000, out_gaf=gaf, out_fasta=None, random_seed=12345).run()

    # -- hooks: record every model the reference builds and every fit it performs
    models = []          # one entry per PathMultinomialModel construction
    fits = []            # one entry per reverse_model_selection / point_estimate call
    orig_init = PathMultinomialModel.__init__

    def rec_init(self, variant_sizes, variant_topos, bins_list, all_sub_paths):
        orig_init(self, variant_sizes, variant_topos, bins_list, all_sub_paths)
        models.append({
            "variant_sizes": [int(s) for s in variant_sizes],
            "bins_list": dump_bins_list(bins_list),
            "n_sub_paths": len(all_sub_paths),
            "observed_sub_path_ids": [go for go, spi in enumerate(all_sub_paths.values())
                                      if spi.mapped_records],
        })
        self._golden_id = len(models) - 1
    PathMultinomialModel.__init__ = rec_init

    orig_rms = mfl.ModelFitMaxLike.reverse_model_selection
    orig_pe = mfl.ModelFitMaxLike.point_estimate

    def rec_rms(self, n_proc, criterion=Criterion.BIC, chosen_ids=None, random_size=20, user_fixed_ids=None):
        res = orig_rms(self, n_proc, criterion=criterion, chosen_ids=chosen_ids,
                       random_size=random_size, user_fixed_ids=user_fixed_ids)
        fits.append({"kind": "reverse_model_selection", "model": self.model._golden_id,
                     "criterion": str(getattr(criterion, "value", criterion)),
                     "bootstrap_mode": self.bootstrap_mode if self.bootstrap_mode else "",
                     "prop": [[int(k), float(v)] for k, v in res[0].items()],
                     "loglike": float(res[1]), "criterion_value": float(res[2]), "via": VIA})
        return res

    def rec_pe(self, chosen_ids=None, criterion=Criterion.BIC):
        res = orig_pe(self, chosen_ids=chosen_ids, criterion=criterion)
        fits.append({"kind": "point_estimate", "model": self.model._golden_id,
                     "criterion": str(getattr(criterion, "value", criterion)),
                     "chosen_ids": sorted(chosen_ids) if chosen_ids else None,
                     "prop": [[int(k), float(v)] for k, v in res[0].items()],
                     "loglike": float(res[1]), "criterion_value": float(res[2]), "via": VIA})
        return res
    mfl.ModelFitMaxLike.reverse_model_selection = rec_rms
    mfl.ModelFitMaxLike.point_estimate = rec_pe

    outdir = os.path.join(tmp, "out")
    os.makedirs(outdir)
    np.random.seed(20240607)   # the reference never seeds np.random on this flow: pinned here
    trv = tvm.Traversome(
        graph=gfa, alignment=gaf, reads_file=None, outdir=outdir, var_fixed=None, var_candidate=var,
        model_criterion=Criterion.BIC, bootstrap=100, bs_threshold=0.95, jackknife=0,
        max_valid_search=0, num_processes=1, force_circular=True, uni_chromosome=False,
        n_generations=0, random_seed=12345, loglevel="INFO", keep_temp=True,
        min_read_identity_cutoff=0.992, min_record_identity_cutoff=0.99, min_alignment_len_cutoff=5000,
        min_alignment_counts="auto", graph_component_selection=0, ignore_conflicts=True,
        keep_graph_redundancy=False)
    trv.run()

    # -- record pool for replaying the reference's resampling (traversome.py:1622-1699)
    sub_paths = list(trv.all_sub_paths.items())
    record_rows = {}
    for go, (_p, spi) in enumerate(sub_paths):
        for rid in spi.mapped_records:
            record_rows[int(rid)] = go
    pool_sorted = sorted(record_rows)
    read_paths = []
    for path, spi in sub_paths:
        entry = {"path": [[str(n), bool(e)] for n, e in path],
                 "inner_len": int(spi.inner_len), "full_len": int(spi.full_len),
                 "from_variants": [[int(k), int(v)] for k, v in spi.from_variants.items()],
                 "n_records": len(spi.mapped_records)}
        if len(path) > 1:  # end statistics used by get_num_of_possible_alignment_start_points (:1558-1567)
            (ln1, le1), (ln2, le2) = path[0], path[1]
            (rn1, re1), (rn2, re2) = path[-1], path[-2]
            li, ri = trv.graph.vertex_info[ln1], trv.graph.vertex_info[rn1]
            entry.update(left_len=int(li.len), left_overlap=int(li.connections[le1][(ln2, not le2)]),
                         right_len=int(ri.len), right_overlap=int(ri.connections[not re1][(rn2, re2)]))
        read_paths.append(entry)

    full_model = models[0]
    pts = [[0.6, 0.4], [0.5, 0.5], [23 / 35., 12 / 35.], [0.9, 0.1], [1.0, 0.0], [0.0, 1.0]]
    subs = [None, None, None, None, {0}, {1}]
    # rebuild a live model object for the known answers
    ka = known_answer_ll(trv.model, pts, subs)

    golden = {
        "what": "BASELINE.json configs[0]; SURVEY.md 8(c) plastome recipe",
        "reference_version": "0.1.7.1",
        "symengine": VIA,
        "np_random_seed_pinned": 20240607,
        "variant_sizes": [int(s) for s in trv.variant_sizes],
        "num_valid_records": int(trv.num_valid_records),
        "be_unidentifiable_to": [[int(k), int(v)] for k, v in trv.be_unidentifiable_to.items()],
        "repr_to_merged_variants": [[int(k), [int(x) for x in v]] for k, v in trv.repr_to_merged_variants.items()],
        "read_paths": read_paths,
        "records": {"pool_sorted": pool_sorted,
                    "row_of_record": [record_rows[r] for r in pool_sorted],
                    "align_len": [int(trv.align_len_at_path_map[r]) for r in pool_sorted]},
        "models": models,
        "fits": fits,
        "known_answer_ll": ka,
        "final": {"variant_proportions": [[int(k), float(v)] for k, v in trv.variant_proportions.items()],
                  "res_loglike": float(trv.res_loglike), "res_criterion": float(trv.res_criterion),
                  "variant_proportions_best": [[int(k), float(v)] for k, v in trv.variant_proportions_best.items()],
                  "bs_eligible": bool(trv.bs_eligible),
                  "support": [[list(map(int, k)), float(v["support"])] for k, v in trv.vp_unique_results.items()]},
    }
    tables = {}
    for fn in ("final.result.tab", "bootstraps.replicates.tab", "variants.info.tab", "readpath.information.tab"):
        fp = os.path.join(outdir, fn)
        if os.path.exists(fp):
            with open(fp) as h:
                tables[fn] = h.read()
    golden["output_tables"] = tables
    with gzip.open(os.path.join(GOLDEN, "plastome_cfg1.json.gz"), "wt") as h:
        json.dump(golden, h, separators=(",", ":"))
    # inputs of the flow, so the whole `thorough` call can be replayed by tests
    shutil.copy(var, os.path.join(GOLDEN, "plastome_variants.txt"))
    with open(gfa, "rb") as i, gzip.open(os.path.join(GOLDEN, "plastome.gfa.gz"), "wb") as o:
        o.write(i.read())
    with open(gaf, "rb") as i, gzip.open(os.path.join(GOLDEN, "plastome_sim.gaf.gz"), "wb") as o:
        o.write(i.read())
    PathMultinomialModel.__init__ = orig_init
    mfl.ModelFitMaxLike.reverse_model_selection = orig_rms
    mfl.ModelFitMaxLike.point_estimate = orig_pe
    print("plastome: %d models, %d fits, final %s LL %.10f BIC %.10f" % (
        len(models), len(fits), golden["final"]["variant_proportions"], trv.res_loglike, trv.res_criterion))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    standin = _import_reference()
    print("symengine stand-in active:", standin)
    os.makedirs(GOLDEN, exist_ok=True)
    if what in ("plastome", "all"):
        make_plastome()
    if what in ("synthetic", "all"):
        from make_golden_synthetic import make_synthetic
        make_synthetic()
