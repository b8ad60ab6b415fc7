"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and fails LOUDLY without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from traversome_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    names = set()
    for header in ("tvs.h", "tvs_subpaths.h", "tvs_pool.h", "tvs_gaf.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(tvs_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_cabi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert _cabi.load().tvs_version() >= 100


def test_argument_validation_needs_no_gpu(lib_built):
    lib = _cabi.load()
    h = ctypes.c_void_p(None)
    rp = np.array([0, 1, 2], dtype=np.int32)
    col = np.array([0, 5], dtype=np.int32)          # column 5 out of range for V=2
    val = np.array([1, 1], dtype=np.int32)
    L = np.array([10, 10], dtype=np.int64)
    rc = lib.tvs_create(ctypes.byref(h), 0, 2, 2, 2, rp.ctypes.data, col.ctypes.data, val.ctypes.data, L.ctypes.data)
    assert rc == _cabi.TVS_EINVAL and b"out of range" in lib.tvs_last_error()
    rc = lib.tvs_create(ctypes.byref(h), 0, 2, 70000, 2, rp.ctypes.data, col.ctypes.data, val.ctypes.data, L.ctypes.data)
    assert rc == _cabi.TVS_ELIMIT
    assert lib.tvs_fit(None, None, None, None, None, None, None, None) == _cabi.TVS_EINVAL
    assert lib.tvs_destroy(None) == _cabi.TVS_OK


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device the product path must raise, not compute on the CPU."""
    if _cabi.device_count() > 0:
        pytest.skip("a GPU is present")
    rp = np.array([0, 1, 2], dtype=np.int32)
    with pytest.raises(_cabi.TvsError) as e:
        _cabi.Engine(rp, np.array([0, 1], np.int32), np.array([1, 1], np.int32), np.array([10, 10], np.int64))
    assert e.value.code == _cabi.TVS_ECUDA


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "traversome_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "/root/reference" not in txt, fn
