#!/usr/bin/env python
"""Developer tool (GPU box): two handles (two CUDA streams) driven by two host threads on one GPU versus one handle:
does the tail of one batch overlap the wide phase of another?  Usage: tools/two_handles.py [lanes_total] [n_handles]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traversome_b200 import _cabi, synthetic  # noqa: E402
wl = synthetic.make_config("cfg2")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = 12


def make(seed, lanes):
    e = _cabi.Engine(wl.row_ptr, wl.col, wl.val, wl.L)
    e.resample(lanes, wl.n_obs, seed=seed, keep_lane0=False)
    e.fit()
    return e


def run(e, n):
    for _ in range(n):
        e.fit()


whole = make(1, B)
t = time.perf_counter(); run(whole, N); one = (time.perf_counter() - t) / N
print("%d lanes, one handle: %.2f ms per batch" % (B, 1e3 * one))
for k in (2, 3, 4):
    parts = [make(10 + i, B // k) for i in range(k)]
    th = [threading.Thread(target=run, args=(e, N)) for e in parts]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    many = (time.perf_counter() - t) / N
    print("%d lanes as %d concurrent handles of %d: %.2f ms per batch (%.2fx)" % (B, k, B // k, 1e3 * many, one / many))
    for e in parts: e.close()
