"""CPU: replicate construction in array form (traversome_b200.resample) against the reference's own replicates.

The golden run (oracle/make_golden.py) recorded the bins of every PathMultinomialModel the reference built: model 0 =
whole data, models 1..100 = its bootstrap replicates, drawn by ``random.choices`` from the module-level RNG seeded with
12345 (traversome.py:142-143,1652).  Replaying those draws and binning with RecordPool must reproduce every range, id
range, count and num_possible_X (bit for bit: float.hex in the fixture)."""
import random

import numpy as np

from tests.helpers import hexf, model_from_tables
from traversome_b200 import csr
from traversome_b200.resample import RecordPool


def _sig(fv):
    return tuple(sorted((int(v), int(c)) for v, c in fv))


def _check_bins(pool, fixture, got, ref_bins_list):
    assert len(got) == len(ref_bins_list)
    row_sig = [_sig(r["from_variants"]) for r in fixture["read_paths"]]
    for g, ref in zip(got, ref_bins_list):
        assert (g["min_len"], g["max_len"], g["min_id"], g["max_id"]) == (ref["min_len"], ref["max_len"], ref["min_id"], ref["max_id"])
        assert len(g["rows"]) == len(ref["rp_bins"])
        # the plastome fixture has read paths with identical from_variants, so match bins by (signature, count, X)
        mine = sorted((row_sig[int(r)], int(n), float(x).hex()) for r, n, x in zip(g["rows"], g["num_matched"], g["X"]))
        theirs = sorted((_sig(b["from_variants"]), int(b["num_matched"]), b["num_possible_X"]) for b in ref["rp_bins"])
        assert mine == theirs


def test_whole_data_bins_bit_exact(plastome):
    pool = RecordPool.from_tables(plastome)
    bins, observed = pool.bins_of(np.arange(pool.n_records))
    _check_bins(pool, plastome, bins, plastome["models"][0]["bins_list"])
    assert sorted(observed.tolist()) == plastome["models"][0]["observed_sub_path_ids"]


def test_bootstrap_replay_reproduces_every_reference_replicate(plastome):
    pool = RecordPool.from_tables(plastome)
    rng = random.Random(12345)                       # traversome.py:142-143 with the CLI default --rs 12345
    for k in range(1, len(plastome["models"])):
        sample = pool.sample_bootstrap(rng, plastome["num_valid_records"])
        bins, observed = pool.bins_of(sample)
        _check_bins(pool, plastome, bins, plastome["models"][k]["bins_list"])
        assert len(observed) == plastome["models"][k]["n_sub_paths"]


def test_lane_of_equals_lane_from_reference_bins(plastome):
    """(w, C, N) from the arrays == (w, C, N) from the reference's bin objects (csr.lane_from_bins)."""
    pool = RecordPool.from_tables(plastome)
    rng = random.Random(12345)
    L = plastome["variant_sizes"]
    row_sig = [_sig(r["from_variants"]) for r in plastome["read_paths"]]
    for k in range(1, 12):
        sample = pool.sample_bootstrap(rng, plastome["num_valid_records"])
        w, C, N, observed = pool.lane_of(sample)
        model = model_from_tables(L, plastome["models"][k]["bins_list"])
        matrix = csr.ReadPathMatrix(len(L))
        w_map, C_ref, N_ref = csr.lane_from_bins(model.bins_list, matrix)
        assert N == N_ref and abs(C - C_ref) <= 1e-9 * abs(C_ref)
        mine = {}
        for r in np.nonzero(w)[0]:
            mine[row_sig[r]] = mine.get(row_sig[r], 0) + int(w[r])
        theirs = {matrix._sigs[r]: n for r, n in w_map.items()}
        assert mine == theirs


def test_jackknife_draw_replays_random_sample(plastome):
    pool = RecordPool.from_tables(plastome)
    a = pool.sample_jackknife(random.Random(7), 17)
    b = random.Random(7).sample(range(pool.n_records), k=pool.n_records - 17)
    assert a.tolist() == b and len(set(a.tolist())) == pool.n_records - 17
