"""GPU (-m gpu): per-replicate multinomial bin statistics on the device (include/tvs_pool.h, csrc/tvs_pool.cu) against
the array-form host implementation (traversome_b200.resample.RecordPool.lane_of), which tests/test_resample_golden.py pins
bit for bit against the reference's own replicates (traversome.py:1305-1384, 1386-1462, 1492-1579, 1636-1699).
Integer outputs (weights, N, observed read paths) must be identical; C agrees to rounding."""
import random

import numpy as np
import pytest

from traversome_b200 import _cabi
from traversome_b200.resample import RecordPool

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(lib_built):
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return True


def _compare(pool, samples, masked=()):
    w, C, N, obs = pool.lanes_on_device(samples, masked_rows=masked)
    for b, s in enumerate(samples):
        w_h, C_h, N_h, obs_h = pool.lane_of(np.asarray(s), masked)
        assert np.array_equal(w[:, b], w_h), b
        assert int(N[b]) == int(N_h), b
        assert np.array_equal(obs[:, b], obs_h), b
        assert abs(C[b] - C_h) <= 1e-12 * max(1.0, abs(C_h)), (b, C[b], C_h)


def test_golden_replicates_of_the_reference_run(gpu, plastome):
    """The 100 bootstrap replicates of the golden run (random.Random(12345) replay) + the whole data, one device call."""
    pool = RecordPool.from_tables(plastome)
    rng = random.Random(12345)
    samples = [np.arange(pool.n_records)]
    for _ in range(1, len(plastome["models"])):
        samples.append(pool.sample_bootstrap(rng, plastome["num_valid_records"]))
    _compare(pool, samples)
    w, C, N, obs = pool.lanes_on_device(samples)
    assert 0 < int(N[0]) <= plastome["num_valid_records"] and w.shape == (pool.n_rows, len(samples))


def _random_pool(rng, n_rows, n_records):
    multi = rng.random(n_rows) < 0.6
    inner = np.where(multi, rng.integers(0, 4000, n_rows), 0)
    full = inner + rng.integers(1, 9000, n_rows) + np.where(multi, 2, 0)
    full[rng.random(n_rows) < 0.05] = 0                                   # a few illegal read paths (min_len > max_len)
    lroom = rng.integers(1, 5000, n_rows)
    rroom = rng.integers(1, 5000, n_rows)
    row_of = rng.integers(0, n_rows, n_records)
    lo = np.where(multi, inner + 2, 1)[row_of]
    hi = np.maximum(full[row_of], lo)
    align = lo + (rng.random(n_records) * (hi - lo + 1)).astype(np.int64)
    at_lo, at_hi = rng.random(n_records) < 0.1, rng.random(n_records) < 0.05      # records at the very ends of their read path's range
    align = np.where(at_lo, lo, np.where(at_hi, hi, align))
    rec_ids = np.cumsum(rng.integers(1, 4, n_records))
    return RecordPool(rec_ids, row_of, align, multi, inner, full, lroom, rroom)


@pytest.mark.parametrize("n_rows,n_records,B", [(7, 300, 9), (60, 5000, 33), (400, 20000, 12)])
def test_random_pools_bootstrap_jackknife_and_masks(gpu, n_rows, n_records, B):
    rng = np.random.default_rng(n_rows * 13 + B)
    pool = _random_pool(rng, n_rows, n_records)
    pr = random.Random(n_records)
    samples = [pool.sample_bootstrap(pr, n_records) for _ in range(B - 3)]
    samples += [pool.sample_jackknife(pr, n_records // 7) for _ in range(2)]          # shorter samples: padded with -1
    samples.append(np.asarray(pr.choices(range(n_records), k=5)))                     # a nearly empty replicate
    _compare(pool, samples)
    masked = sorted(set(rng.integers(0, n_rows, max(1, n_rows // 5)).tolist()))
    _compare(pool, samples[:5], masked)
