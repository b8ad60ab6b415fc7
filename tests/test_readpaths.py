"""Read-path construction from alignment records: the oracle and the array-form product against
outputs of the unmodified reference (tests/golden/readpaths.json.gz).  Host-side integer work: no GPU involved."""
import numpy as np
import pytest

from oracle import readpaths as orp
from oracle.subpaths import PlainGraph, as_path
from traversome_b200 import readpaths as rp


class Rec(object):
    def __init__(self, path, p_align_len):
        self.path, self.p_align_len = path, p_align_len


@pytest.fixture(scope="module")
def golden():
    import gzip, json, os
    here = os.path.dirname(os.path.abspath(__file__))
    with gzip.open(os.path.join(here, "golden", "readpaths.json.gz"), "rt") as h:
        return json.load(h)


def runs_of(golden):
    for case in golden["cases"]:
        graph = PlainGraph.from_json(case["graph"])
        for run in case["runs"]:
            recs = case["records"] if run["record_subset"] is None else [case["records"][i] for i in run["record_subset"]]
            yield case["label"], graph, [as_path(p) for p, _ in recs], [int(l) for _, l in recs], run


def check(run, table, lengths):
    assert [[[[n, bool(e)] for n, e in p], list(ids)] for p, ids in table.items()] == run["read_paths"]
    assert sorted([int(k), int(v)] for k, v in lengths.items()) == run["align_len_at_path_map"]


def test_oracle_matches_the_reference(golden):
    n = 0
    for label, graph, paths, lens, run in runs_of(golden):
        table, lengths = orp.read_paths_from_records(graph, paths, lens, run["filter_by_graph"], run["min_alignment_counts"])
        check(run, table, lengths)
        n += 1
    assert n == 15


def product(graph, paths, lens, run):
    records = [Rec(p, l) for p, l in zip(paths, lens)]
    out = rp.generate_read_paths(graph, records, filter_by_graph=run["filter_by_graph"],
                                 min_alignment_counts=run["min_alignment_counts"])
    check(run, out.read_paths, out.align_len_at_path_map)
    assert out.num_valid_records == run["num_valid_records"]
    assert (out.min_alignment_length, out.max_alignment_length) == (run["min_alignment_length"], run["max_alignment_length"])
    assert out.max_read_path_size == run["max_read_path_size"]


def test_array_form_matches_the_reference(golden):
    for label, graph, paths, lens, run in runs_of(golden):
        product(graph, paths, lens, run)


def test_nothing_left_exits_like_the_reference(golden):
    case = golden["cases"][0]
    graph = PlainGraph.from_json(case["graph"])
    with pytest.raises(SystemExit):
        rp.generate_read_paths(graph, [Rec((("ghost", True),), 100)])
    with pytest.raises(SystemExit):
        rp.generate_read_paths(graph, [])


def test_array_form_equals_oracle_on_a_larger_random_record_set():
    import random
    from tests.subpath_helpers import random_graph_case
    graph, variants, _, _ = random_graph_case(77, 60, 24, (20, 60), True, 5, 6, 700)
    rnd = random.Random(3)
    paths, lens = [], []
    for _ in range(6000):
        v = variants[rnd.randrange(len(variants))]
        s0, k = int(len(v) * rnd.random() ** 2) % len(v), rnd.randint(1, 7)
        w = tuple(v[(s0 + i) % len(v)] for i in range(k))
        if rnd.random() < 0.5:
            w = tuple((n, not e) for n, e in reversed(w))
        paths.append(w)
        lens.append(rnd.randint(100, 9000))
    for mac in (1, 2, 5):
        table, lengths = orp.read_paths_from_records(graph, paths, lens, True, mac)
        out = rp.generate_read_paths(graph, [Rec(p, l) for p, l in zip(paths, lens)], min_alignment_counts=mac)
        assert list(out.read_paths.items()) == list(table.items()) and out.align_len_at_path_map == lengths


def test_hash_collisions_fall_back_to_exact_comparison(monkeypatch):
    """The hashes are lookup devices only: with a degenerate hash (every permutation of a path collides) the grouping and
    the window support still equal the oracle's."""
    import random
    from tests.subpath_helpers import random_graph_case
    monkeypatch.setattr(rp, "_HASH_MUL", 1)
    graph, variants, _, _ = random_graph_case(78, 20, 8, (10, 30), True, 5, 6, 700)
    rnd = random.Random(4)
    paths, lens = [], []
    for _ in range(1500):
        v = variants[rnd.randrange(len(variants))]
        s0, k = rnd.randrange(len(v)), rnd.randint(1, 5)
        w = tuple(v[(s0 + i) % len(v)] for i in range(k))
        if rnd.random() < 0.5:
            w = tuple((n, not e) for n, e in reversed(w))
        paths.append(w)
        lens.append(rnd.randint(100, 900))
    for mac in (1, 3):
        table, lengths = orp.read_paths_from_records(graph, paths, lens, True, mac)
        out = rp.generate_read_paths(graph, [Rec(p, l) for p, l in zip(paths, lens)], min_alignment_counts=mac)
        assert list(out.read_paths.items()) == list(table.items()) and out.align_len_at_path_map == lengths
