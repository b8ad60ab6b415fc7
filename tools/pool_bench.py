#!/usr/bin/env python
"""Developer tool (GPU box): per-replicate bin statistics -- the device batch call (tvs_pool_lanes)
against the host array form (RecordPool.lane_of), on the plastome fixture and on a synthetic pool.
Usage: tools/pool_bench.py [n_rows] [n_records] [replicates]"""
import gzip, json, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200.resample import RecordPool  # noqa: E402
from tests.test_gpu_pool import _random_pool  # noqa: E402


def run(pool, n_rep, label, host_reps):
    pr = random.Random(1)
    samples = [pool.sample_bootstrap(pr, pool.n_records) for _ in range(n_rep)]
    pool.lanes_on_device(samples[:2])                       # creates the device pool, loads the kernels
    t0 = time.perf_counter()
    w, C, N, obs = pool.lanes_on_device(samples)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    for s in samples[:host_reps]:
        pool.lane_of(s)
    t_host = (time.perf_counter() - t0) / host_reps
    w_h = pool.lane_of(samples[0])[0]
    print(json.dumps({"pool": label, "records": pool.n_records, "read_paths": pool.n_rows, "replicates": n_rep,
                      "device_ms_per_replicate": round(1e3 * t_dev / n_rep, 4), "device_call_ms": round(1e3 * t_dev, 2),
                      "host_array_ms_per_replicate": round(1e3 * t_host, 3), "identical_weights": bool(np.array_equal(w[:, 0], w_h))}))


with gzip.open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "plastome_cfg1.json.gz"), "rt") as h:
    run(RecordPool.from_tables(json.load(h)), 100, "plastome (cfg1)", 20)
n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_rec = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
n_rep = int(sys.argv[3]) if len(sys.argv) > 3 else 256
run(_random_pool(np.random.default_rng(3), n_rows, n_rec), n_rep, "synthetic", 3)
