"""GPU: tvs_subpaths_count (include/tvs_subpaths.h) against the reference's outputs (tests/golden/subpaths.json.gz),
against the oracle on seeded random graphs, and against the plain-Python statement of its contract."""
import numpy as np
import pytest

from oracle import subpaths as osp
from traversome_b200 import _cabi
from traversome_b200 import subpaths as sp
from tests.subpath_helpers import emulate_count, random_graph_case
from tests.test_subpaths_oracle import check_against_golden, check_plastome

pytestmark = pytest.mark.gpu


def test_golden_cases_through_the_device(subpaths_golden, lib_built):
    for case in subpaths_golden["cases"]:
        counts = check_against_golden(case, sp.VariantSubPathsGenerator)
        assert counts.stats["n_windows"] > 0 and counts.stats["kernel_ms"] > 0


def test_plastome_tables_through_the_device(plastome, lib_built):
    """BASELINE configs[0]: from the graph, the two isomers and the observed read paths to the reference's
    all_sub_paths table (order, lengths, from_variants)."""
    check_plastome(plastome, sp.VariantSubPathsGenerator)


@pytest.mark.parametrize("seed,circular,n_contigs,n_variants,var_len", [
    (21, True, 40, 24, (5, 60)), (22, False, 40, 24, (5, 60)), (23, True, 300, 64, (50, 200)),
    (24, False, 8, 40, (1, 12)), (25, True, 5, 16, (1, 4)),
])
def test_counts_equal_oracle_on_random_graphs(seed, circular, n_contigs, n_variants, var_len, lib_built):
    graph, variants, read_paths, mal = random_graph_case(seed, n_contigs, n_variants, var_len, circular, 30, 7, 900)
    gen = sp.VariantSubPathsGenerator(graph, 10, mal, read_paths)
    counts = gen.count(variants)
    want = [osp.count_read_paths_in_variant(graph, v, set(read_paths), mal) for v in variants]
    got = counts.counters()
    assert [list(c.items()) for c in got] == [list(c.items()) for c in want]        # counts and dict order
    # the raw arrays equal the contract emulation bit for bit
    rows, keys = gen._prepare()
    enc = sp.encode_variants(gen.tables, variants)
    row_ptr, col, val, first, stats = emulate_count(*enc, len(rows), *keys, mal - 2)
    assert np.array_equal(row_ptr, counts.row_ptr) and np.array_equal(col, counts.col)
    assert np.array_equal(val, counts.val) and np.array_equal(first, counts.first)
    assert stats["n_windows"] == counts.stats["n_windows"] and stats["nnz"] == counts.stats["nnz"]


def test_repeat_calls_are_identical_and_single_variant_form_agrees(lib_built):
    graph, variants, read_paths, mal = random_graph_case(31, 60, 32, (20, 80), True, 40, 6, 800)
    gen = sp.VariantSubPathsGenerator(graph, 10, mal, read_paths)
    a, b = gen.count(variants), gen.count(variants)
    for x, y in ((a.row_ptr, b.row_ptr), (a.col, b.col), (a.val, b.val), (a.first, b.first)):
        assert np.array_equal(x, y)
    batch = a.counters()
    for v, want in zip(variants[:6], batch[:6]):
        assert list(gen.gen_subpaths(v).items()) == list(want.items())


def test_edge_shapes(lib_built):
    graph, variants, read_paths, mal = random_graph_case(41, 12, 6, (3, 9), True, 10, 3, 300)
    # no read paths at all / read paths that never occur
    empty = sp.VariantSubPathsGenerator(graph, 10, mal, []).count(variants)
    assert empty.stats["nnz"] == 0 and empty.row_ptr.tolist() == [0]
    assert all(len(c) == 0 for c in empty.counters())
    # no variants
    none = sp.VariantSubPathsGenerator(graph, 10, mal, read_paths).count([])
    assert none.stats["nnz"] == 0 and none.row_order().size == 0
    # max_alignment_len so small that only single contigs and pairs are examined
    tiny = sp.VariantSubPathsGenerator(graph, 10, 2, read_paths).count(variants)
    want = [osp.count_read_paths_in_variant(graph, v, set(read_paths), 2) for v in variants]
    assert [list(c.items()) for c in tiny.counters()] == [list(c.items()) for c in want]
    assert all(len(p) <= 2 for c in want for p in c)


def test_argument_validation(lib_built):
    one = np.array([0, 1], dtype=np.int64)
    with pytest.raises(_cabi.TvsError) as e:
        _cabi.count_subpaths(np.array([0, 0], dtype=np.int64), [], [], [0], 1, one, [0], [0], [0], 10)
    assert e.value.code == _cabi.TVS_EINVAL
    with pytest.raises(ValueError):
        _cabi.count_subpaths(one, [0], [0], [1], 1, one, [0], [5], [0], 10)


def test_counts_feed_the_fitter(lib_built):
    """The CSR of the counting is what tvs_create takes: fit a two-variant mixture on it."""
    graph, variants, read_paths, mal = random_graph_case(51, 30, 2, (30, 40), True, 60, 5, 700)
    gen = sp.VariantSubPathsGenerator(graph, 10, mal, read_paths)
    counts = gen.count(variants)
    row_ptr, col, val, rows = counts.csr_in_reference_order()
    L = np.array([osp.full_length(graph, v) for v in variants], dtype=np.int64)
    eng = _cabi.Engine(row_ptr, col, val, L)
    R = rows.size
    dense = counts.dense()[rows].astype(np.float64)
    p = dense @ np.array([0.7, 0.3])                     # P(read path) is proportional to sum_v A[r,v] theta_v
    w = np.round(50000 * p / p.sum()).astype(np.int32).reshape(R, 1)
    eng.set_lanes(w, np.ones((2, 1), dtype=np.uint8), np.zeros(1))
    res = eng.fit()
    assert res["status"][0] == _cabi.CONVERGED and res["kkt"][0, 0] <= 1e-9
    best = res["loglike"][0]
    t0 = res["theta"][0, 0]
    assert 0.0 < t0 < 1.0
    # the read paths are a sample of the variants' windows, so the optimum need not be 0.7; it must be a maximum
    for t in (t0 - 0.05, t0 - 1e-3, t0 + 1e-3, t0 + 0.05, 0.7):
        if 0.0 < t < 1.0:
            assert eng.loglik(np.array([[t], [1.0 - t]]))[0] <= best + 1e-9 * abs(best)


def test_large_variant_by_read_path_product(lib_built):
    """20,000 variants x 100,000 read paths: 16 GB as a dense scratch, a few MB as a hit list.  Raw C-ABI arrays,
    compared with the plain-Python statement of the contract."""
    rng = np.random.default_rng(11)
    n_var, var_len, n_tok_kinds = 20000, 6, 4000
    var_tok = rng.integers(0, n_tok_kinds, size=n_var * var_len).astype(np.int32)
    var_ptr = np.arange(0, n_var * var_len + 1, var_len, dtype=np.int64)
    var_step = rng.integers(50, 400, size=var_tok.size).astype(np.int32)
    var_circ = (rng.random(n_var) < 0.5).astype(np.uint8)
    keys = {}
    while len(keys) < 100000:
        ln = int(rng.integers(1, 4))
        if rng.random() < 0.7:                              # most keys are real windows
            v = int(rng.integers(0, n_var))
            s = int(rng.integers(0, var_len - ln + 1))
            seq = tuple(var_tok[v * var_len + s:v * var_len + s + ln].tolist())
        else:
            seq = tuple(rng.integers(0, n_tok_kinds, size=ln).tolist())
        keys.setdefault(seq, len(keys))
    seqs = list(keys)
    key_ptr = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(q) for q in seqs], out=key_ptr[1:])
    key_tok = np.fromiter((t for q in seqs for t in q), dtype=np.int32, count=int(key_ptr[-1]))
    key_row = np.arange(len(seqs), dtype=np.int32)
    lin_row = np.where(rng.random(len(seqs)) < 0.9, key_row, -1).astype(np.int32)
    got = _cabi.count_subpaths(var_ptr, var_tok, var_step, var_circ, len(seqs), key_ptr, key_tok, key_row, lin_row, 600)
    want = emulate_count(var_ptr, var_tok, var_step, var_circ, len(seqs), key_ptr, key_tok, key_row, lin_row, 600)
    for a, b in zip(got[:4], want[:4]):
        assert np.array_equal(a, b)
    assert got[4]["nnz"] == want[4]["nnz"] > 50000 and got[4]["n_windows"] == want[4]["n_windows"]
