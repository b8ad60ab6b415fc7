"""CPU: the blocked sparse layout of the wide-candidate-set likelihood pass (blk_kernel, csrc/tvs_blk.cuh).

The host code that tvs_create runs (csrc/tvs_blk_host.h) is compiled with g++ and checked against numpy: the forward
product A x and the transposed product A^T r computed by walking the segments in the kernel's order must reproduce the
CSR, including counts that have to be split into several 16-bit-representable entries.
Mirrors ModelGenerator.py:153-174 (the bins x variants sum the blocked matrix encodes)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("blk") / "blk_host_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "blk_host_harness.cpp")])
    lib = ctypes.CDLL(out)
    lib.blk_host_products.restype = ctypes.c_int
    return lib


def random_csr(rng, R, V, density, big_counts=False):
    pat = rng.random((R, V)) < density
    pat[np.arange(R), rng.integers(V, size=R)] = True
    vals = rng.choice([1, 2, 3], p=[0.85, 0.12, 0.03], size=(R, V))
    if big_counts:
        vals = np.where(rng.random((R, V)) < 0.2, rng.integers(1, 65536, size=(R, V)), vals)
    a = (pat * vals).astype(np.int64)
    rr, cc = np.nonzero(a)
    row_ptr = np.zeros(R + 1, dtype=np.int32)
    np.cumsum(np.bincount(rr, minlength=R), out=row_ptr[1:])
    return a, row_ptr, cc.astype(np.int32), a[rr, cc].astype(np.int32)


def products(lib, a, row_ptr, col, val, x, r):
    R, V = a.shape
    p = np.empty(R)
    t = np.empty(V)
    info = np.zeros(8, dtype=np.int64)
    rc = lib.blk_host_products(ctypes.c_int64(R), ctypes.c_int64(V), ctypes.c_int64(col.size), row_ptr.ctypes.data_as(ctypes.c_void_p),
                               col.ctypes.data_as(ctypes.c_void_p), val.ctypes.data_as(ctypes.c_void_p),
                               x.ctypes.data_as(ctypes.c_void_p), p.ctypes.data_as(ctypes.c_void_p),
                               r.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p),
                               info.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return p, t, info


@pytest.mark.parametrize("R,V,density,big", [(700, 300, 0.05, False), (1000, 2048, 0.05, False), (257, 129, 0.3, False),
                                             (513, 640, 0.1, True), (40, 4096, 0.02, False), (1, 1, 1.0, True)])
def test_blocked_products_reproduce_the_csr(harness, R, V, density, big):
    rng = np.random.default_rng(R * 7 + V)
    a, row_ptr, col, val = random_csr(rng, R, V, density, big)
    # integer-valued operands: every partial sum is exact in fp64, so the comparison is exact whatever the order
    x = rng.integers(1, 1000, size=V).astype(np.float64)
    r = rng.integers(1, 1000, size=R).astype(np.float64)
    p, t, info = products(harness, a, row_ptr, col, val, x, r)
    assert np.array_equal(p, (a @ x.astype(np.int64)).astype(np.float64))
    assert np.array_equal(t, (a.T @ r.astype(np.int64)).astype(np.float64))
    assert info[0] == (R + 239) // 240 and info[1] == (V + 127) // 128     # chunks of 240 major x panels of 128 minor indices
    assert info[4] == (V + 239) // 240 and info[5] == (R + 127) // 128
    assert info[2] % 16 == 0 and info[6] % 16 == 0


def test_blocked_products_with_real_operands(harness):
    rng = np.random.default_rng(5)
    a, row_ptr, col, val = random_csr(rng, 900, 500, 0.08)
    x = rng.random(500)
    r = rng.random(900)
    p, t, _ = products(harness, a, row_ptr, col, val, x, r)
    np.testing.assert_allclose(p, a @ x, rtol=1e-13)
    np.testing.assert_allclose(t, a.T @ r, rtol=1e-13)
