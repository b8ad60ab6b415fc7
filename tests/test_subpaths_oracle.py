"""CPU: (1) the oracle restatement of the read-path x variant counting reproduces the reference's outputs
(tests/golden/subpaths.json.gz, written by oracle/make_golden_subpaths.py from the unmodified reference);
(2) the host side of the product (graph tables, variant and key encoders, result tables) reproduces them too when the
device call is replaced by a plain-Python statement of its contract."""
from collections import OrderedDict

import numpy as np
import pytest

from oracle import subpaths as osp
from traversome_b200 import subpaths as sp
from tests.subpath_helpers import emulate_count, load_case, random_graph_case


def case_ids(golden):
    return [c["label"] for c in golden["cases"]]


def expected_counters(case):
    return [OrderedDict((osp.as_path(p), c) for p, c in counter) for counter in case["expected"]["counters"]]


def test_oracle_matches_the_reference(subpaths_golden):
    assert len(subpaths_golden["cases"]) >= 8
    for case in subpaths_golden["cases"]:
        graph, variants, read_paths = load_case(case)
        assert [osp.closes(graph, v) for v in variants] == case["variant_circular"], case["label"]
        counters, be, merged, table = osp.informative_sub_paths(graph, variants, case["variant_sizes"], read_paths,
                                                                case["max_alignment_len"])
        exp = expected_counters(case)
        for got, want in zip(counters, exp):
            assert list(got.items()) == list(want.items()), case["label"]          # counts AND dict order
        assert [[k, v] for k, v in be.items()] == case["expected"]["be_unidentifiable_to"]
        assert [[k, v] for k, v in merged.items()] == case["expected"]["repr_to_merged_variants"]
        want_table = case["expected"]["all_sub_paths"]
        assert [p for p in table] == [osp.as_path(e["path"]) for e in want_table]
        for (p, got), want in zip(table.items(), want_table):
            assert got["inner_len"] == want["inner_len"] and got["full_len"] == want["full_len"]
            assert [[k, c] for k, c in got["from_variants"].items()] == want["from_variants"]


class EmulatedGenerator(sp.VariantSubPathsGenerator):
    """The product's generator with the device call swapped for the contract emulation (CPU tests only)."""

    def count(self, variant_paths):
        rows, (key_ptr, key_tok, key_row_circ, key_row_lin) = self._prepare()
        variant_paths = [tuple(p) for p in variant_paths]
        var_ptr, var_tok, var_step, var_circ = sp.encode_variants(self.tables, variant_paths)
        out = emulate_count(var_ptr, var_tok, var_step, var_circ, len(rows), key_ptr, key_tok, key_row_circ, key_row_lin,
                            int(self.max_alignment_len) - 2)
        return sp.SubPathCounts(rows, len(variant_paths), *out)


def check_against_golden(case, generator_cls):
    graph, variants, read_paths = load_case(case)
    gen = generator_cls(graph, 10, case["max_alignment_len"], read_paths)
    counters, be, merged, table, counts = sp.gen_all_informative_sub_paths(gen, variants, case["variant_sizes"])
    exp = expected_counters(case)
    for v, want in zip(variants, exp):
        assert list(counters[v].items()) == list(want.items()), case["label"]
        assert list(gen.gen_subpaths(v).items()) == list(want.items())
    assert [[k, v] for k, v in be.items()] == case["expected"]["be_unidentifiable_to"], case["label"]
    assert [[k, v] for k, v in merged.items()] == case["expected"]["repr_to_merged_variants"]
    want_table = case["expected"]["all_sub_paths"]
    assert list(table) == [osp.as_path(e["path"]) for e in want_table], case["label"]
    for (p, info), want in zip(table.items(), want_table):
        assert (info.inner_len, info.full_len) == (want["inner_len"], want["full_len"])
        assert [[k, c] for k, c in info.from_variants.items()] == want["from_variants"]
    # the CSR handed to tvs_create: rows in all_sub_paths order, columns increasing
    row_ptr, col, val, rows = counts.csr_in_reference_order()
    assert row_ptr.size == len(want_table) + 1
    for i, want in enumerate(want_table):
        assert sorted(want["from_variants"]) == [[int(c), int(x)] for c, x in
                                                 zip(col[row_ptr[i]:row_ptr[i + 1]], val[row_ptr[i]:row_ptr[i + 1]])]
    return counts


def test_host_encoders_match_the_reference(subpaths_golden):
    for case in subpaths_golden["cases"]:
        check_against_golden(case, EmulatedGenerator)


def test_variant_encoding_details(subpaths_golden):
    case = subpaths_golden["cases"][0]
    graph, variants, _ = load_case(case)
    tables = sp.GraphTables(graph)
    ptr, tok, step, circ = sp.encode_variants(tables, variants)
    assert circ.tolist() == [int(c) for c in case["variant_circular"]]
    v = variants[0]
    for j in range(len(v)):
        name, strand = v[j]
        prev = v[j - 1]
        assert step[ptr[0] + j] == graph.vertex_info[name].len - graph.vertex_info[name].connections[not strand][prev]
    # a step over a link the graph does not have fails like the reference's dict lookup (utils.py:454)
    missing = None
    for a in graph.vertex_info:
        for b in graph.vertex_info:
            if (b, False) not in graph.vertex_info[a].connections[True]:
                missing = ((a, True), (b, True))
                break
        if missing:
            break
    with pytest.raises(KeyError):
        sp.encode_variants(tables, [variants[1], missing + missing])


def test_unstandardised_read_paths_are_never_hit():
    graph, variants, read_paths, mal = random_graph_case(5, 10, 3, (6, 10), True, 10, 3, 500)
    flipped = [p for p in (osp._revcomp(graph, r) for r in read_paths) if p not in set(read_paths)]
    gen = EmulatedGenerator(graph, 10, mal, read_paths + flipped)
    counts = gen.count(variants)
    dense = counts.dense()
    assert dense[:len(read_paths)].sum() > 0
    assert dense[len(read_paths):].sum() == 0                       # like `not in read_paths_hashed` in the reference
    want = [osp.count_read_paths_in_variant(graph, v, set(read_paths + flipped), mal) for v in variants]
    assert [list(c.items()) for c in counts.counters()] == [list(c.items()) for c in want]


def test_emulation_equals_oracle_on_random_graphs():
    for seed, circular in ((1, True), (2, False), (3, True)):
        graph, variants, read_paths, mal = random_graph_case(seed, 30, 12, (5, 40), circular, 25, 6, 700)
        gen = EmulatedGenerator(graph, 10, mal, read_paths)
        got = gen.count(variants).counters()
        want = [osp.count_read_paths_in_variant(graph, v, set(read_paths), mal) for v in variants]
        assert [list(c.items()) for c in got] == [list(c.items()) for c in want]


def test_array_key_encoder_equals_the_general_one(subpaths_golden):
    def as_set(k):
        return {(tuple(k[1][k[0][i]:k[0][i + 1]].tolist()), int(k[2][i]), int(k[3][i])) for i in range(len(k[2]))}

    n_closed = 0
    for case in subpaths_golden["cases"]:
        graph, _, read_paths = load_case(case)
        tables = sp.GraphTables(graph)
        if tables.palindromic:
            continue
        n_closed += sum(len(r) > 1 and tables.is_circular(r) for r in read_paths)
        assert as_set(sp.encode_read_paths(tables, read_paths)) == as_set(sp._encode_read_paths_general(tables, read_paths))
    assert n_closed > 0                                   # the closed-walk branch was exercised
    graph, _, read_paths, _ = random_graph_case(9, 50, 20, (10, 60), True, 40, 6, 700)
    tables = sp.GraphTables(graph)
    # non-standard and duplicated-by-reversal read paths are handled alike
    extra = [p for p in (tables.reverse(r) for r in read_paths[:40]) if p not in set(read_paths)]
    assert extra
    assert as_set(sp.encode_read_paths(tables, read_paths + extra)) == \
        as_set(sp._encode_read_paths_general(tables, read_paths + extra))


def plastome_inputs(plastome):
    """BASELINE configs[0]: the processed plastome graph (contig lengths as the fixture's full_len values imply,
    overlaps trimmed to 0 by the reference's graph processing), the two isomer variants, the observed read paths."""
    lens = {}
    for rp in plastome["read_paths"]:
        if len(rp["path"]) == 1:
            lens[rp["path"][0][0]] = rp["full_len"]
    links = []
    for a, ae, b, be in (("LSC", True, "IR", True), ("LSC", False, "IR", True), ("IR", True, "SSC", True),
                         ("IR", True, "SSC", False)):
        links.append([a, ae, b, not be, 0])
        links.append([b, not be, a, ae, 0])
    graph = osp.PlainGraph([[n, lens[n]] for n in ("LSC", "IR", "SSC")], links)
    variants = [tuple((tok[:-1], tok[-1] == "+") for tok in line.replace("(circular)", "").split(","))
                for line in ("LSC+,IR+,SSC+,IR-", "LSC+,IR+,SSC-,IR-")]
    # the reference standardises user variants first (Assembly.get_standardized_variant, Assembly.py:1229-1260:
    # smallest rotation over both strands)
    variants = [osp.canonical(graph, v, rotations=True) for v in variants]
    read_paths = [osp.as_path(rp["path"]) for rp in plastome["read_paths"]]
    return graph, variants, read_paths[::-1], max(plastome["records"]["align_len"])


def check_plastome(plastome, generator_cls):
    graph, variants, read_paths, mal = plastome_inputs(plastome)
    gen = generator_cls(graph, min(plastome["records"]["align_len"]), mal, read_paths)
    _, be, merged, table, counts = sp.gen_all_informative_sub_paths(gen, variants, plastome["variant_sizes"])
    want = plastome["read_paths"]
    assert list(table) == [osp.as_path(rp["path"]) for rp in want]             # the reference's all_sub_paths order
    for info, rp in zip(table.values(), want):
        assert (info.inner_len, info.full_len) == (rp["inner_len"], rp["full_len"])
        assert [[v, c] for v, c in info.from_variants.items()] == rp["from_variants"]
    assert [[k, v] for k, v in be.items()] == [[0, 0], [1, 1]]
    return counts


def test_plastome_tables_from_graph_and_variants(plastome):
    graph, variants, read_paths, mal = plastome_inputs(plastome)
    _, _, _, table = osp.informative_sub_paths(graph, variants, plastome["variant_sizes"], read_paths, mal)
    assert list(table) == [osp.as_path(rp["path"]) for rp in plastome["read_paths"]]
    check_plastome(plastome, EmulatedGenerator)
