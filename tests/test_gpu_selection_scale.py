"""GPU (-m gpu): backward model selection at BASELINE configs[2] scale -- 256 candidate variants of which 32 carry reads,
chunks of random_size = 20 candidates with a PINNED permutation (the reference shuffles with the unseeded
process-global np.random, ModelFitMaxLike.py:203) -- against a replay of the same decision logic on the exact CPU solver
(oracle/exact.py behind tests/helpers.OracleEngine), drop for drop; and lock step (all replicates' candidates in one GPU
batch per round, traversome_b200.bootstrap.run_selections) against the serial loop of the reference
(traversome.py:515-556: one replicate after the other).  Reference: ModelFitMaxLike.py:140-322, 349-370."""
import numpy as np
import pytest
from loguru import logger

from tests.helpers import OracleEngine
from traversome_b200 import _cabi, bootstrap, synthetic
from traversome_b200.ModelFitMaxLike import ModelFitMaxLike
from traversome_b200.lanes import LaneContext

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(lib_built):
    if _cabi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return True


def _fitters(wl, W, engine_factory, seeds):
    ctx = LaneContext.from_csr(wl.V, wl.L, wl.row_ptr, wl.col, wl.val)
    if engine_factory is not None:
        ctx.engine_factory = engine_factory
    pattern = np.zeros((wl.R, wl.V), dtype=bool)
    pattern[np.repeat(np.arange(wl.R), np.diff(wl.row_ptr)), wl.col] = True
    groups = {v: [v] for v in range(wl.V)}
    ident = {v: v for v in range(wl.V)}
    out = []
    for b, seed in enumerate(seeds):
        col = ctx.add_weights(W[:, b])
        f = ModelFitMaxLike.from_lane(ctx, (col, 0.0, int(W[:, b].sum())), pattern, np.nonzero(W[:, b])[0], groups, ident)
        f._shuffle = np.random.default_rng(seed).shuffle            # pinned per replicate
        out.append(f)
    return ctx, out


def _drops(run):
    """The 'Drop cid_x' lines a selection writes (INFO outside bootstrap mode), in order."""
    seen = []
    sink = logger.add(lambda m: seen.append(m.record["message"]), level="INFO", filter=lambda r: r["message"].startswith("Drop "))
    try:
        res = run()
    finally:
        logger.remove(sink)
    return res, seen


def test_chunked_selection_from_256_candidates_equals_the_exact_replay(gpu):
    wl = synthetic.make_workload(256, 6000, 0.10, seed=1003, n_true=32)
    W = wl.n_obs.reshape(-1, 1).astype(np.int32)
    ctx_g, (fg,) = _fitters(wl, W, None, [11])
    ctx_o, (fo,) = _fitters(wl, W, OracleEngine, [11])
    (pg, lg, cg), drops_g = _drops(lambda: fg.reverse_model_selection(n_proc=1, criterion="BIC", random_size=20))
    (po, lo, co), drops_o = _drops(lambda: fo.reverse_model_selection(n_proc=1, criterion="BIC", random_size=20))
    assert len(drops_g) > 100                                   # chunking was live: > 200 candidates left after the first fit
    assert drops_g == drops_o                                   # the same variant dropped at every step
    assert sorted(pg) == sorted(po)
    assert all(abs(pg[k] - po[k]) < 1e-6 for k in pg)
    assert abs(lg - lo) <= 1e-9 * abs(lo) and abs(cg - co) <= 1e-9 * abs(co)
    true = set(np.nonzero(wl.theta_true)[0].tolist())
    # what BIC keeps besides the read-carrying variants are variants whose removal costs more log-likelihood than ln(N)/2
    # each: the greedy backward search is exact here (equal to the replay), it does not recover the 32 alone
    assert true & set(pg) == {v for v in true if v in pg} and len(pg) >= 1
    ctx_g.close()


def test_lock_step_equals_serial_selection_with_more_than_20_candidates(gpu):
    wl = synthetic.make_workload(48, 2500, 0.15, seed=1004, n_true=12)
    n_rep = 6
    W = synthetic.bootstrap_weights(wl.n_obs, n_rep, seed=5, keep_lane0=True)
    seeds = [100 + b for b in range(n_rep)]
    ctx_a, fa = _fitters(wl, W, None, seeds)
    lock = bootstrap.run_selections(fa, criterion="BIC", random_size=20)
    ctx_b, fb = _fitters(wl, W, None, seeds)
    serial = [f.reverse_model_selection(n_proc=1, criterion="BIC", random_size=20) for f in fb]
    assert ctx_a.n_batches < ctx_b.n_batches                    # the lock-step run shares its batches
    for a, b in zip(lock, serial):
        assert sorted(a[0]) == sorted(b[0])
        assert all(abs(a[0][k] - b[0][k]) < 1e-7 for k in a[0])
        assert abs(a[1] - b[1]) <= 1e-10 * abs(b[1]) and abs(a[2] - b[2]) <= 1e-10 * abs(b[2])
    ctx_a.close()
    ctx_b.close()


def test_replicate_groups_select_the_same_models_as_one_group(gpu):
    """run_selections(groups=2 / 3): the replicates advance on several host threads whose solver calls alternate; the
    selected models, proportions and criteria are those of one lock-step group."""
    wl = synthetic.make_workload(48, 2500, 0.15, seed=1005, n_true=12)
    n_rep = 9
    W = synthetic.bootstrap_weights(wl.n_obs, n_rep, seed=6, keep_lane0=True)
    seeds = [200 + b for b in range(n_rep)]
    ctx_a, fa = _fitters(wl, W, None, seeds)
    one = bootstrap.run_selections(fa, criterion="BIC", random_size=20, groups=1)
    for g in (2, 3):
        ctx_b, fb = _fitters(wl, W, None, seeds)
        many = bootstrap.run_selections(fb, criterion="BIC", random_size=20, groups=g)
        assert ctx_b.n_batches > ctx_a.n_batches and ctx_b.n_lanes_fitted == ctx_a.n_lanes_fitted
        for a, b in zip(one, many):
            assert sorted(a[0]) == sorted(b[0])
            assert all(abs(a[0][k] - b[0][k]) < 1e-7 for k in a[0])
            assert abs(a[1] - b[1]) <= 1e-10 * abs(b[1]) and abs(a[2] - b[2]) <= 1e-10 * abs(b[2])
        ctx_b.close()
    ctx_a.close()
