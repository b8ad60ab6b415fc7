"""CPU: the result tables are byte-identical to the files the reference wrote in the plastome
run of BASELINE configs[0] (tests/golden/plastome_cfg1.json.gz: output_tables), when fed the reference's own results;
other shapes of the tally (unsupported raw result, solutions that are not re-evaluated, jackknife naming, no
replicates) are checked against the format rules of traversome.py:669-781."""
from collections import OrderedDict

from oracle import subpaths as osp
from traversome_b200 import bootstrap, output
from traversome_b200.utils import Criterion
from tests.test_subpaths_oracle import plastome_inputs


def plastome_run(plastome):
    reps = [OrderedDict((int(c), p) for c, p in fit["prop"]) for fit in plastome["fits"][1:]]
    raw = OrderedDict((int(c), p) for c, p in plastome["final"]["variant_proportions"])
    summary = bootstrap.summarize_replicates(reps, raw, plastome["final"]["res_loglike"],
                                             plastome["final"]["res_criterion"])
    return summary, raw


def test_plastome_tables_are_byte_identical(plastome):
    want = plastome["output_tables"]
    summary, raw = plastome_run(plastome)
    graph, variants, read_paths, mal = plastome_inputs(plastome)
    counters, be, _, _ = osp.informative_sub_paths(graph, variants, plastome["variant_sizes"], read_paths, mal)
    counters = {v: c for v, c in zip(variants, counters)}
    # the reference lists read paths in order of first alignment record; the table itself carries that order
    n_records = {osp.as_path(rp["path"]): rp["n_records"] for rp in plastome["read_paths"]}
    by_gaf = {output.path_to_gaf_str(p): p for p in n_records}
    order = [by_gaf[line.split("\t")[-2]] for line in want["readpath.information.tab"].strip().split("\n")[1:]]
    reads = OrderedDict((p, list(range(n_records[p]))) for p in order)
    tables = output.all_tables(summary, raw, plastome["final"]["res_loglike"], plastome["final"]["res_criterion"],
                               variants, plastome["variant_sizes"], be, reads, counters, Criterion.BIC, bootstrap=True)
    for name in ("variants.info.tab", "readpath.information.tab", "final.result.tab", "bootstraps.replicates.tab"):
        assert tables[name] == want[name], name
    assert tables["readpath.record_ids.tab"].split("\n")[1].startswith("0\t0,1,2,")


def test_format_rules_of_the_other_tally_shapes():
    reps = [OrderedDict([(2, 0.7), (0, 0.3)])] * 3 + [OrderedDict([(2, 1.0)])] * 2 + [OrderedDict([(1, 1.0)])] * 15
    raw = OrderedDict([(0, 0.25), (3, 0.75)])                       # never seen in a replicate
    refits = []

    def refit(ids):
        refits.append(ids)
        return OrderedDict((c, 1.0 / len(ids)) for c in sorted(ids)), -10.5, 30.25

    s = bootstrap.summarize_replicates(reps, raw, -9.0, 25.0, refit=refit)
    cid_to_fid = output.cid_to_fid_table(s.vp_unique_results_sorted, raw)
    assert list(cid_to_fid.items()) == [(1, 1), (2, 2), (0, 3), (3, 4)]      # order of appearance, raw result last
    text = output.final_result_tab(s, raw, -9.0, 25.0, cid_to_fid, Criterion.AIC, bootstrap=False)
    lines = text.split("\n")
    assert lines[0] == "SOLUTIONS\tJACKKNIFE_SUPPORT\tLOGLIKELIHOOD\tAIC\tfid_1\tfid_2\tfid_3\tfid_4\tFREQ_STD_OF_JACKKNIFE"
    assert lines[1] == "1\t0.75\t-10.5\t30.25\t1.0000\t-\t-\t-\t0.0000"           # best supported, re-evaluated
    assert lines[2] == "2\t0.15\t-10.5\t30.25\t-\t0.5000\t0.5000\t-\t0.0000,0.0000"
    assert lines[3] == "3\t0.1\t-10.5\t30.25\t-\t1.0000\t-\t-\t0.0000"            # support > 0.05: re-evaluated
    assert lines[4] == "4\t0\t-9.0\t25.0\t-\t-\t0.2500\t0.7500\t-"               # the unsupported raw result
    assert refits == [{1}, {2, 0}, {2}]
    assert s.variant_proportions_best == {1: 1.0}
    rep = output.replicates_tab(reps, cid_to_fid, bootstrap=False)
    assert rep.split("\n")[0] == "JACKKNIFE ID\tfid_1\tfid_2\tfid_3\tfid_4"
    assert rep.split("\n")[1] == "jackknife_00\t-\t0.7000\t0.3000\t-"
    assert rep.split("\n")[20] == "jackknife_19\t1.0000\t-\t-\t-"

    # low support, not re-evaluated: '*' placeholders and replicate means are not printed as proportions
    reps = [OrderedDict([(0, 1.0)])] * 39 + [OrderedDict([(1, 0.4), (0, 0.6)])]
    s = bootstrap.summarize_replicates(reps, OrderedDict([(0, 1.0)]), -1.0, 2.0, refit=refit)
    cid_to_fid = output.cid_to_fid_table(s.vp_unique_results_sorted, {0: 1.0})
    lines = output.final_result_tab(s, {0: 1.0}, -1.0, 2.0, cid_to_fid).split("\n")
    assert lines[1] == "1\t0.975\t-1.0\t2.0\t1.0000\t-\t" + ",".join(["0.0000"])
    assert lines[2] == "2\t0.025\t*\t*\t0.6000\t0.4000\t*,*"

    # no replicates at all
    empty = bootstrap.summarize_replicates([], OrderedDict([(5, 1.0)]), -3.0, 8.0)
    cid_to_fid = output.cid_to_fid_table(empty.vp_unique_results_sorted, {5: 1.0})
    assert output.final_result_tab(empty, {5: 1.0}, -3.0, 8.0, cid_to_fid).split("\n")[1] == "1\t1.0\t-\t-\t1.0000\t-"
    assert output.replicates_tab([], cid_to_fid) == ""


def test_variant_table_marks_unidentifiable_variants():
    paths = [(("a", True), ("b", False)), (("a", True), ("c", True), ("b", False))]
    text = output.variants_info_tab(OrderedDict([(1, 1), (0, 2)]), {0: 0, 1: 0}, paths, [100, 130])
    assert text.split("\n")[1:3] == ["1\t1\t0\t3\t130\t>a>c<b", "2\t0\t-\t2\t100\t>a<b"]
