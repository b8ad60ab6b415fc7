"""ctypes binding of ``libtvs.so`` (C ABI declared in ``include/tvs.h``).

There is deliberately no CPU fallback: if the shared library is missing, or no CUDA device is
usable, every entry point raises.  Build the library with ``python __graft_entry__.py build`` or
``traversome_b200/csrc/build.sh``.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtvs.so")

# every symbol include/tvs.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "tvs_version", "tvs_last_error", "tvs_device_count", "tvs_create", "tvs_destroy", "tvs_set_lanes",
    "tvs_set_lanes_u16", "tvs_set_lanes_indexed", "tvs_columns_reserve", "tvs_columns_upload", "tvs_columns_upload_u16",
    "tvs_columns_resample", "tvs_set_lanes_columns", "tvs_resample", "tvs_get_weights", "tvs_loglik", "tvs_fit", "tvs_last_fit_stats", "tvs_dims",
    "tvs_measure_fp64_peak", "tvs_measure_smem_peak", "tvs_selftest_math", "tvs_selftest_table_log",
    # include/tvs_subpaths.h
    "tvs_subpaths_count", "tvs_subpaths_fetch", "tvs_subpaths_destroy",
    # include/tvs_pool.h
    "tvs_pool_create", "tvs_pool_lanes", "tvs_pool_destroy",
    # include/tvs_gaf.h (host code: alignment-file parser)
    "tvs_gaf_parse", "tvs_gaf_parse_spa", "tvs_gaf_dims", "tvs_gaf_columns", "tvs_gaf_steps", "tvs_gaf_segments", "tvs_gaf_segment_ids",
    "tvs_gaf_destroy",
)

TVS_OK, TVS_EINVAL, TVS_ECUDA, TVS_ENOMEM, TVS_ELIMIT, TVS_ESTATE = 0, -1, -2, -3, -4, -5
CONVERGED, MAXITER, STALLED, INFEASIBLE = 0, 1, 2, 3
_ERRNAMES = {-1: "TVS_EINVAL", -2: "TVS_ECUDA", -3: "TVS_ENOMEM", -4: "TVS_ELIMIT", -5: "TVS_ESTATE"}


class TvsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s: %s" % (_ERRNAMES.get(code, str(code)), message))
        self.code = code


class FitOptions(ctypes.Structure):
    _fields_ = [("tol", c_double), ("max_iter", c_int32), ("history", c_int32), ("check_every", c_int32),
                ("reserved", c_int32)]


class FitStats(ctypes.Structure):
    _fields_ = [("passes", c_int64), ("kernel_launches", c_int64), ("pass_ms", c_double), ("total_ms", c_double),
                ("lane_passes", c_int64), ("update_ms", c_double), ("compactions", c_int64)]


_lib = None


def load():
    """Load libtvs.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("traversome_b200: %s not found -- build it with `python __graft_entry__.py build` "
                          "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.tvs_version.restype = c_int
    lib.tvs_last_error.restype = c_char_p
    lib.tvs_device_count.argtypes = [POINTER(c_int)]
    lib.tvs_create.argtypes = [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.tvs_destroy.argtypes = [c_void_p]
    lib.tvs_set_lanes.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.tvs_set_lanes_u16.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.tvs_set_lanes_indexed.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.tvs_columns_reserve.argtypes = [c_void_p, c_int64]
    lib.tvs_columns_upload.argtypes = [c_void_p, c_int64, c_int64, c_void_p]
    lib.tvs_columns_upload_u16.argtypes = [c_void_p, c_int64, c_int64, c_void_p]
    lib.tvs_columns_resample.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_uint64, c_uint64]
    lib.tvs_set_lanes_columns.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.tvs_resample.argtypes = [c_void_p, c_int64, c_void_p, c_uint64, c_int, c_void_p, c_void_p]
    lib.tvs_get_weights.argtypes = [c_void_p, c_void_p]
    lib.tvs_loglik.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.tvs_fit.argtypes = [c_void_p, c_void_p, POINTER(FitOptions), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.tvs_last_fit_stats.argtypes = [c_void_p, POINTER(FitStats)]
    lib.tvs_dims.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    lib.tvs_measure_fp64_peak.argtypes = [c_int, POINTER(c_double)]
    lib.tvs_measure_smem_peak.argtypes = [c_int, POINTER(c_double)]
    lib.tvs_selftest_math.argtypes = [c_int, c_int64, c_void_p, c_void_p, c_void_p]
    lib.tvs_selftest_table_log.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
    lib.tvs_subpaths_count.argtypes = [POINTER(c_void_p), c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                       POINTER(c_int64), POINTER(c_int64), POINTER(c_double)]
    lib.tvs_subpaths_fetch.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.tvs_subpaths_destroy.argtypes = [c_void_p]
    lib.tvs_pool_create.argtypes = [POINTER(c_void_p), c_int, c_int64, c_int64] + [c_void_p] * 8
    lib.tvs_pool_lanes.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.tvs_pool_destroy.argtypes = [c_void_p]
    lib.tvs_gaf_parse.argtypes = [c_char_p, c_int64, c_double, c_int, c_int, POINTER(c_void_p)]
    lib.tvs_gaf_parse_spa.argtypes = [c_char_p, c_int64, c_int, POINTER(c_void_p)]
    lib.tvs_gaf_dims.argtypes = [c_void_p] + [POINTER(c_int64)] * 6
    lib.tvs_gaf_columns.argtypes = [c_void_p] + [c_void_p] * 17
    lib.tvs_gaf_steps.argtypes = [c_void_p] + [c_void_p] * 6
    lib.tvs_gaf_segments.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64)]
    lib.tvs_gaf_segment_ids.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.tvs_gaf_destroy.argtypes = [c_void_p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("tvs_last_error",):
            fn.restype = c_int
    _lib = lib
    return lib


def _check(code):
    if code != TVS_OK:
        raise TvsError(code, load().tvs_last_error().decode("utf-8", "replace"))


def device_count():
    n = c_int(0)
    lib = load()
    rc = lib.tvs_device_count(ctypes.byref(n))
    if rc != TVS_OK:
        return 0
    return n.value


def _ptr(a):
    """host numpy array / torch tensor (cpu or cuda) / raw int address / None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return c_void_p(a)
    if isinstance(a, np.ndarray):
        return c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return c_void_p(a.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(a))


def _as(a, dtype):
    """C-contiguous host array of dtype (no copy when already so); torch tensors pass through."""
    if a is None or hasattr(a, "data_ptr") or isinstance(a, int):
        return a
    return np.ascontiguousarray(a, dtype=dtype)


class Engine(object):
    """One shared CSR on one GPU + the current batch of lanes.  Thin, typed wrapper over the C ABI."""

    def __init__(self, row_ptr, col, val, variant_sizes, n_variants=None, device=0):
        lib = load()
        self._lib = lib
        self._h = c_void_p(None)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.int32)
        self.L = np.ascontiguousarray(variant_sizes, dtype=np.int64)
        self.R = int(self.row_ptr.shape[0] - 1)
        self.V = int(self.L.shape[0] if n_variants is None else n_variants)
        self.nnz = int(self.col.shape[0])
        self.B = 0
        self.device = int(device)
        _check(lib.tvs_create(ctypes.byref(self._h), self.device, self.R, self.V, self.nnz, _ptr(self.row_ptr),
                              _ptr(self.col), _ptr(self.val), _ptr(self.L)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.tvs_destroy(self._h)
            self._h = c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- lanes
    def set_lanes(self, w, mask=None, C=None, B=None):
        """w [R,B] lane-minor, int32 -- or uint16 (numpy array / torch tensor), which halves the upload;
        mask [V,B] uint8 or None; C [B] float64 or None."""
        is_u16 = (isinstance(w, np.ndarray) and w.dtype == np.uint16) or str(getattr(w, "dtype", "")) == "torch.uint16"
        if is_u16:
            w = np.ascontiguousarray(w) if isinstance(w, np.ndarray) else w
            if B is None:
                B = int(w.shape[1]) if len(w.shape) == 2 else int(w.numel() // self.R)
            mask = _as(mask, np.uint8)
            C = _as(C, np.float64)
            self._keep = (w, mask, C)
            _check(self._lib.tvs_set_lanes_u16(self._h, int(B), _ptr(w), _ptr(mask), _ptr(C)))
            self.B = int(B)
            return
        w = _as(w, np.int32)
        if B is None:
            B = int(w.shape[1]) if w.ndim == 2 else int(w.size // self.R)
        mask = _as(mask, np.uint8)
        C = _as(C, np.float64)
        self._keep = (w, mask, C)
        _check(self._lib.tvs_set_lanes(self._h, int(B), _ptr(w), _ptr(mask), _ptr(C)))
        self.B = int(B)

    def set_lanes_indexed(self, w_cols, col_of_lane, mask=None, C=None):
        """w_cols [R,n_cols] int32 distinct weight columns; lane b uses column col_of_lane[b]."""
        w_cols = _as(w_cols, np.int32)
        col_of_lane = _as(col_of_lane, np.int32)
        B = int(col_of_lane.shape[0])
        n_cols = int(w_cols.shape[1]) if w_cols.ndim == 2 else int(w_cols.size // self.R)
        mask = _as(mask, np.uint8)
        C = _as(C, np.float64)
        _check(self._lib.tvs_set_lanes_indexed(self._h, B, n_cols, _ptr(w_cols), _ptr(col_of_lane), _ptr(mask), _ptr(C)))
        self.B = B

    # ---- device-resident weight columns
    def columns_reserve(self, n_cols):
        _check(self._lib.tvs_columns_reserve(self._h, int(n_cols)))

    def columns_upload(self, first, w_cols):
        """w_cols [R, n] int32 or uint16 (numpy / torch, host or device) -> slots first .. first+n-1."""
        is_u16 = (isinstance(w_cols, np.ndarray) and w_cols.dtype == np.uint16) or str(getattr(w_cols, "dtype", "")) == "torch.uint16"
        if not is_u16:
            w_cols = _as(w_cols, np.int32)
        elif isinstance(w_cols, np.ndarray):
            w_cols = np.ascontiguousarray(w_cols)
        n = int(w_cols.shape[1]) if len(w_cols.shape) == 2 else int((w_cols.size if isinstance(w_cols, np.ndarray) else w_cols.numel()) // self.R)
        fn = self._lib.tvs_columns_upload_u16 if is_u16 else self._lib.tvs_columns_upload
        _check(fn(self._h, int(first), n, _ptr(w_cols)))

    def columns_resample(self, first, n, n_obs, seed, first_id):
        n_obs = _as(n_obs, np.int32)
        _check(self._lib.tvs_columns_resample(self._h, int(first), int(n), _ptr(n_obs), c_uint64(int(seed)), c_uint64(int(first_id))))

    def set_lanes_columns(self, col_of_lane, mask=None, C=None):
        """lane b <- slot col_of_lane[b] of the column store (no weights cross PCIe)."""
        col_of_lane = _as(col_of_lane, np.int32)
        B = int(col_of_lane.shape[0])
        mask = _as(mask, np.uint8)
        C = _as(C, np.float64)
        self._keep = (col_of_lane, mask, C)
        _check(self._lib.tvs_set_lanes_columns(self._h, B, _ptr(col_of_lane), _ptr(mask), _ptr(C)))
        self.B = B

    def resample(self, B, n_obs, seed, keep_lane0=False, mask=None, C=None):
        n_obs = _as(n_obs, np.int32)
        mask = _as(mask, np.uint8)
        C = _as(C, np.float64)
        _check(self._lib.tvs_resample(self._h, int(B), _ptr(n_obs), c_uint64(int(seed)), int(bool(keep_lane0)),
                                      _ptr(mask), _ptr(C)))
        self.B = int(B)

    def get_weights(self):
        out = np.empty((self.R, self.B), dtype=np.int32)
        _check(self._lib.tvs_get_weights(self._h, _ptr(out)))
        return out

    # ---- evaluation / fit
    def loglik(self, theta):
        theta = _as(theta, np.float64)
        out = np.empty(self.B, dtype=np.float64)
        _check(self._lib.tvs_loglik(self._h, _ptr(theta), _ptr(out)))
        return out

    def fit(self, theta0=None, tol=0.0, max_iter=0, history=0, check_every=0, time_passes=False, out=None):
        """Returns dict(theta [V,B], loglike [B], kkt [B,2], iters [B], status [B]).
        `out` may hold preallocated (e.g. pinned) arrays under the same keys."""
        opt = FitOptions(float(tol), int(max_iter), int(history), int(check_every), 1 if time_passes else 0)
        theta0 = _as(theta0, np.float64)
        o = out if out is not None else {}
        theta = o.get("theta") if o.get("theta") is not None else np.empty((self.V, self.B), dtype=np.float64)
        ll = o.get("loglike") if o.get("loglike") is not None else np.empty(self.B, dtype=np.float64)
        kkt = o.get("kkt") if o.get("kkt") is not None else np.empty((self.B, 2), dtype=np.float64)
        iters = o.get("iters") if o.get("iters") is not None else np.empty(self.B, dtype=np.int32)
        status = o.get("status") if o.get("status") is not None else np.empty(self.B, dtype=np.int32)
        _check(self._lib.tvs_fit(self._h, _ptr(theta0), ctypes.byref(opt), _ptr(theta), _ptr(ll), _ptr(kkt),
                                 _ptr(iters), _ptr(status)))
        return {"theta": theta, "loglike": ll, "kkt": kkt, "iters": iters, "status": status}

    def selftest_table_log(self, p):
        """log(p) by the table-driven routine of the staged likelihood pass."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        lg = np.empty_like(p)
        _check(self._lib.tvs_selftest_table_log(self._h, p.size, _ptr(p), _ptr(lg)))
        return lg

    def last_fit_stats(self):
        st = FitStats()
        _check(self._lib.tvs_last_fit_stats(self._h, ctypes.byref(st)))
        return {"passes": st.passes, "kernel_launches": st.kernel_launches, "pass_ms": st.pass_ms,
                "total_ms": st.total_ms, "lane_passes": st.lane_passes, "update_ms": st.update_ms,
                "compactions": st.compactions}


def measure_fp64_peak(device=0):
    v = c_double(0.0)
    _check(load().tvs_measure_fp64_peak(int(device), ctypes.byref(v)))
    return v.value


def measure_smem_peak(device=0):
    """GB/s of conflict-free 16-byte shared-memory loads (LDS.128) over all SMs: the roof of the sparse products' operand."""
    v = c_double(0.0)
    _check(load().tvs_measure_smem_peak(int(device), ctypes.byref(v)))
    return v.value


def selftest_math(p, device=0):
    """(log(p), 1/p) computed by the device routines of the likelihood pass."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    lg = np.empty_like(p)
    rc = np.empty_like(p)
    _check(load().tvs_selftest_math(int(device), p.size, _ptr(p), _ptr(lg), _ptr(rc)))
    return lg, rc


def count_subpaths(var_ptr, var_tok, var_step, var_circular, n_rows, key_ptr, key_tok, key_row_circular,
                   key_row_linear, max_internal_len, device=0):
    """tvs_subpaths_count + tvs_subpaths_fetch: CSR (by read path) of the read-path x variant occurrence counts.

    Returns (row_ptr int64[n_rows+1], col int32[nnz], val int32[nnz], first int64[nnz], stats dict)."""
    lib = load()
    var_ptr = np.ascontiguousarray(var_ptr, dtype=np.int64)
    var_tok = np.ascontiguousarray(var_tok, dtype=np.int32)
    var_step = np.ascontiguousarray(var_step, dtype=np.int32)
    var_circular = np.ascontiguousarray(var_circular, dtype=np.uint8)
    key_ptr = np.ascontiguousarray(key_ptr, dtype=np.int64)
    key_tok = np.ascontiguousarray(key_tok, dtype=np.int32)
    key_row_circular = np.ascontiguousarray(key_row_circular, dtype=np.int32)
    key_row_linear = np.ascontiguousarray(key_row_linear, dtype=np.int32)
    n_variants, n_keys = var_ptr.size - 1, key_ptr.size - 1
    if var_tok.size != var_ptr[-1] or var_step.size != var_ptr[-1] or var_circular.size != n_variants:
        raise ValueError("variant arrays do not match var_ptr")
    if key_tok.size != key_ptr[-1] or key_row_circular.size != n_keys or key_row_linear.size != n_keys:
        raise ValueError("key arrays do not match key_ptr")
    for rows in (key_row_circular, key_row_linear):
        if n_keys and (rows.min() < -1 or rows.max() >= n_rows):
            raise ValueError("key rows out of range")
    h = c_void_p()
    nnz, n_windows, ms = c_int64(), c_int64(), c_double()
    _check(lib.tvs_subpaths_count(ctypes.byref(h), int(device), n_variants, var_ptr.ctypes.data, var_tok.ctypes.data,
                                  var_step.ctypes.data, var_circular.ctypes.data, int(n_rows), n_keys,
                                  key_ptr.ctypes.data, key_tok.ctypes.data, key_row_circular.ctypes.data,
                                  key_row_linear.ctypes.data,
                                  int(max_internal_len), ctypes.byref(nnz), ctypes.byref(n_windows), ctypes.byref(ms)))
    try:
        row_ptr = np.empty(int(n_rows) + 1, dtype=np.int64)
        col = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value, dtype=np.int32)
        first = np.empty(nnz.value, dtype=np.int64)
        _check(lib.tvs_subpaths_fetch(h, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, first.ctypes.data))
    finally:
        lib.tvs_subpaths_destroy(h)
    return row_ptr, col, val, first, {"n_windows": n_windows.value, "kernel_ms": ms.value, "nnz": nnz.value}


class DevicePool(object):
    """The record pool of one run on the device (include/tvs_pool.h): per-replicate bin statistics -> lanes, batched."""

    def __init__(self, record_ids, row_of, align_len, multi, inner_len, full_len, left_room, right_room, device=0):
        lib = load()
        self._lib = lib
        self._h = c_void_p(None)
        self._keep = (np.ascontiguousarray(record_ids, dtype=np.int64), np.ascontiguousarray(row_of, dtype=np.int32),
                      np.ascontiguousarray(align_len, dtype=np.int64), np.ascontiguousarray(multi, dtype=np.uint8),
                      np.ascontiguousarray(inner_len, dtype=np.int64), np.ascontiguousarray(full_len, dtype=np.int64),
                      np.ascontiguousarray(left_room, dtype=np.int64), np.ascontiguousarray(right_room, dtype=np.int64))
        self.n_records, self.n_rows = int(self._keep[0].size), int(self._keep[3].size)
        _check(lib.tvs_pool_create(ctypes.byref(self._h), int(device), self.n_records, self.n_rows, *[_ptr(a) for a in self._keep]))

    def lanes(self, samples, masked_rows=None):
        """samples int32 [B, n_draw] (-1 = no draw) -> (w int32 [R, B], C float64 [B], N int64 [B], observed bool [R, B])."""
        samples = np.ascontiguousarray(samples, dtype=np.int32)
        B, n_draw = samples.shape
        masked = None if masked_rows is None else np.ascontiguousarray(masked_rows, dtype=np.uint8)
        w = np.empty((self.n_rows, B), dtype=np.int32)
        obs = np.empty((self.n_rows, B), dtype=np.uint8)
        C = np.empty(B, dtype=np.float64)
        N = np.empty(B, dtype=np.int64)
        _check(self._lib.tvs_pool_lanes(self._h, B, n_draw, _ptr(samples), _ptr(masked), _ptr(w), _ptr(C), _ptr(N), _ptr(obs)))
        return w, C, N, obs.astype(bool)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.tvs_pool_destroy(self._h)
            self._h = c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
