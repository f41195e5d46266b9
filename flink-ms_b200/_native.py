"""ctypes binding of libflinkms_b200.so -- the same C-ABI (include/flinkms.h) a JVM host binds through JNI.

There is NO fallback: if the shared library is missing or no sm_100a GPU is present, loading / creating a
context raises.  Nothing in this package imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FMS_LIB") or os.path.join(_HERE, "libflinkms_b200.so")   # FMS_LIB: A/B a differently built library

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_OOM = -1, -2, -3, -4, -5
X_DENSE_F32, X_DENSE_F64, X_CSR_F64 = 0, 1, 2
RNG_PER_BLOCK, RNG_SHARED_SUBTASK = 0, 1
USERS, ITEMS = 0, 1
ALS_KERNEL_AUTO, ALS_KERNEL_GENERIC, ALS_KERNEL_TC, ALS_KERNEL_TC_TMA, ALS_KERNEL_TC_LDGSTS = 0, 1, 2, 3, 4
SVM_KERNEL_AUTO, SVM_KERNEL_GENERIC, SVM_KERNEL_WARP = 0, 1, 2

_vp = C.c_void_p
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol include/flinkms.h declares
SIGNATURES = {
    "fms_create": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int), C.c_int]),
    "fms_destroy": (None, [_vp]),
    "fms_last_error": (C.c_char_p, [_vp]),
    "fms_version": (C.c_int, []),
    "fms_stream": (_vp, [_vp]),
    "fms_synchronize": (C.c_int, [_vp]),
    "fms_launch_count": (C.c_int64, [_vp]),
    "fms_als_fit": (C.c_int, [_vp, _i64p, _i64p, _f64p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int64, _i64p, _i64p]),
    "fms_als_get_factors": (C.c_int, [_vp, C.c_int, _i64p, _f64p]),
    "fms_als_get_factor_rows": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int64, _i64p, _f64p]),
    "fms_als_sgd_step": (C.c_int, [_vp, _i64p, _i64p, _f64p, C.c_int64, C.c_double, C.c_double, C.c_double, _f64p, _f64p, C.c_int, _f64p, _f64p]),
    "fms_parse_ratings_csv": (C.c_int, [_vp, C.c_char_p, C.c_int64, C.c_char, C.c_int, _i64p, _i64p, _f64p, C.c_int64, _i64p]),
    "fms_parse_libsvm": (C.c_int, [_vp, C.c_char_p, C.c_int64, _f64p, C.c_int64, _i64p, _i32p, _f64p, C.c_int64, _i64p, _i64p, _i32p]),
    "fms_als_set_model": (C.c_int, [_vp, _i64p, _f64p, C.c_int64, _i64p, _f64p, C.c_int64, C.c_int]),
    "fms_als_predict": (C.c_int, [_vp, _i64p, _i64p, C.c_int64, _f64p, _u8p]),
    "fms_als_mse": (C.c_int, [_vp, _i64p, _i64p, _f64p, C.c_int64, _f64p, _i64p]),
    "fms_als_load": (C.c_int, [_vp, _i64p, _i64p, _f64p, C.c_int64, C.c_int, C.c_int, _i64p, _i64p]),
    "fms_als_init": (C.c_int, [_vp, C.c_int, C.c_int64]),
    "fms_als_half_step": (C.c_int, [_vp, C.c_int, C.c_double]),
    "fms_als_half_step_chunk": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_int]),
    "fms_als_factors_device": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), _i64p, _i64p, _i64p]),
    "fms_als_factors_updated": (C.c_int, [_vp, C.c_int]),
    "fms_als_planes_device": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _i64p]),
    "fms_als_planes_ipc_handle": (C.c_int, [_vp, C.c_int, _vp]),
    "fms_als_planes_open_peers": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "fms_als_planes_updated": (C.c_int, [_vp, C.c_int]),
    "fms_als_planes_peer_count": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "fms_als_planes_flag_sync": (C.c_int, [_vp, C.c_int]),
    "fms_svm_exchange_ipc_handle": (C.c_int, [_vp, C.c_int, _vp]),
    "fms_svm_exchange_open_peers": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "fms_svm_round_exchange": (C.c_int, [_vp, C.c_int64]),
    "fms_svm_exchange_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "fms_als_last_info": (C.c_int, [_vp, _i64p]),
    "fms_als_set_kernel": (C.c_int, [_vp, C.c_int]),
    "fms_svm_fit": (C.c_int, [_vp, _vp, C.c_int, _i64p, _i32p, _f64p, C.c_int64, C.c_int32, _i32p, C.c_int, C.c_int,
                              C.c_int64, C.c_double, C.c_double, C.c_int64, C.c_int, _f64p]),
    "fms_svm_predict": (C.c_int, [_vp, _vp, C.c_int, _i64p, _i32p, C.c_int64, C.c_int32, _f64p, C.c_double, C.c_int, _f64p]),
    "fms_svm_set_weights": (C.c_int, [_vp, _f64p, C.c_int32]),
    "fms_svm_get_weights": (C.c_int, [_vp, _f64p, C.c_int32]),
    "fms_svm_get_alpha": (C.c_int, [_vp, _f64p, C.c_int64]),
    "fms_als_stream_begin": (C.c_int, [_vp, C.c_int64]),
    "fms_als_stream_append": (C.c_int, [_vp, _i64p, _i64p, _f64p, C.c_int64]),
    "fms_als_stream_fit": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_int64, _i64p, _i64p]),
    "fms_svm_stream_begin": (C.c_int, [_vp, C.c_int64, C.c_int32]),
    "fms_svm_stream_append": (C.c_int, [_vp, _vp, C.c_int, _f64p, _i32p, C.c_int64]),
    "fms_svm_stream_fit": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_int64, C.c_int, _f64p]),
    "fms_svm_load": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _i64p, _i32p, _f64p, C.c_int64, C.c_int32, _i32p, C.c_int,
                               C.c_int, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int]),
    "fms_svm_sweep": (C.c_int, [_vp, C.c_int64]),
    "fms_svm_delta_device": (C.c_int, [_vp, C.POINTER(_vp), _i32p]),
    "fms_svm_apply_delta": (C.c_int, [_vp]),
    "fms_svm_set_kernel": (C.c_int, [_vp, C.c_int]),
    "fms_debug_index_stream": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int32, C.c_int64, _i32p]),
    "fms_debug_als_gramian": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int64, C.c_int64, C.POINTER(C.c_float)]),
}

_lib = None


class FlinkMSError(RuntimeError):
    """What the JNI shim raises as a RuntimeException: non-zero status + fms_last_error."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[fms status {status}] {message}")
        self.status = status


def load():
    """dlopen the in-tree shared library and type every exported symbol.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C flink-ms_b200).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One fms_ctx: one GPU, one stream.  Not thread-safe (like the reference's per-thread helpers)."""

    def __init__(self, device: int | None = None, devices=None):
        """device: one GPU (None = current).  devices=[d0, d1, ...]: a GROUP ctx -- this one process drives all of them through
        the whole-fit entry points (als_fit, svm_fit; model read-out and predict are served by d0); see include/flinkms.h."""
        self._lib = load()
        self._h = _vp()
        if devices is not None:
            devices = [int(d) for d in devices]
            ids, n = (C.c_int * len(devices))(*devices), len(devices)
        else:
            ids, n = (None if device is None else (C.c_int * 1)(device)), 1
        self.n_devices = n
        rc = self._lib.fms_create(C.byref(self._h), ids, n)
        if rc != OK:
            raise FlinkMSError(rc, (self._lib.fms_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fms_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise FlinkMSError(rc, (self._lib.fms_last_error(self._h) or b"").decode())

    # ---- misc
    @property
    def stream(self) -> int:
        return int(self._lib.fms_stream(self._h) or 0)

    def synchronize(self):
        self._check(self._lib.fms_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.fms_launch_count(self._h))

    # ---- ALS
    def als_fit(self, user, item, rating, rank, iterations, lam, seed):
        u, i, r = _i64(user), _i64(item), _f64(rating)
        if not (len(u) == len(i) == len(r)):
            raise ValueError("user/item/rating lengths differ")
        nu, ni = C.c_int64(), C.c_int64()
        self._check(self._lib.fms_als_fit(self._h, _p(u, _i64p), _p(i, _i64p), _p(r, _f64p), len(u), rank, iterations,
                                          lam, seed, C.byref(nu), C.byref(ni)))
        self._als_shape = (nu.value, ni.value, rank)
        return nu.value, ni.value

    def als_fit_stream(self, user, item, rating, rank, iterations, lam, seed, chunk=1 << 20):
        """fms_als_fit fed through the chunked-upload entry points (what a JVM host with < 2 GiB direct buffers does)."""
        u, i, r = _i64(user), _i64(item), _f64(rating)
        self._check(self._lib.fms_als_stream_begin(self._h, len(u)))
        for b in range(0, len(u), chunk):
            e = min(len(u), b + chunk)
            self._check(self._lib.fms_als_stream_append(self._h, _p(u[b:e], _i64p), _p(i[b:e], _i64p), _p(r[b:e], _f64p), e - b))
        nu, ni = C.c_int64(), C.c_int64()
        self._check(self._lib.fms_als_stream_fit(self._h, rank, iterations, lam, seed, C.byref(nu), C.byref(ni)))
        self._als_shape = (nu.value, ni.value, rank)
        return nu.value, ni.value

    def als_load(self, user, item, rating, shard_index=0, shard_count=1):
        u, i, r = _i64(user), _i64(item), _f64(rating)
        nu, ni = C.c_int64(), C.c_int64()
        self._check(self._lib.fms_als_load(self._h, _p(u, _i64p), _p(i, _i64p), _p(r, _f64p), len(u), shard_index,
                                           shard_count, C.byref(nu), C.byref(ni)))
        self._als_shape = (nu.value, ni.value, 0)
        return nu.value, ni.value

    def als_load_device(self, user_ptr, item_ptr, rating_ptr, nnz, shard_index=0, shard_count=1):
        """fms_als_load with DEVICE pointers (int64, int64, float64 arrays of nnz elements already in HBM)."""
        nu, ni = C.c_int64(), C.c_int64()
        self._check(self._lib.fms_als_load(self._h, C.cast(C.c_void_p(user_ptr), _i64p), C.cast(C.c_void_p(item_ptr), _i64p),
                                           C.cast(C.c_void_p(rating_ptr), _f64p), nnz, shard_index, shard_count, C.byref(nu), C.byref(ni)))
        self._als_shape = (nu.value, ni.value, 0)
        return nu.value, ni.value

    def als_get_factor_rows(self, which, row_begin, row_end, ids_out=None, factors_out=None):
        """Rows [row_begin, row_end) of a side (a rank's own shard): ids and float64 factors, optionally into caller buffers."""
        rank = self._als_shape[2]
        m = row_end - row_begin
        ids = np.empty(m, np.int64) if ids_out is None else ids_out
        fac = np.empty((m, rank), np.float64) if factors_out is None else factors_out
        if ids.dtype != np.int64 or fac.dtype != np.float64 or ids.size < m or fac.size < m * rank or not fac.flags.c_contiguous:
            raise ValueError("als_get_factor_rows: output buffers have the wrong dtype / size")
        self._check(self._lib.fms_als_get_factor_rows(self._h, which, row_begin, row_end, _p(ids, _i64p), _p(fac, _f64p)))
        return ids[:m], fac.reshape(-1, rank)[:m] if fac.ndim == 1 else fac[:m]

    def als_init(self, rank, seed):
        self._check(self._lib.fms_als_init(self._h, rank, seed))
        self._als_shape = (self._als_shape[0], self._als_shape[1], rank)

    def als_half_step(self, which, lam):
        self._check(self._lib.fms_als_half_step(self._h, which, lam))

    def als_half_step_chunk(self, which, lam, chunk, n_chunks):
        self._check(self._lib.fms_als_half_step_chunk(self._h, which, lam, chunk, n_chunks))

    def als_set_kernel(self, kernel):
        self._check(self._lib.fms_als_set_kernel(self._h, kernel))

    def als_last_info(self) -> int:
        v = C.c_int64()
        self._check(self._lib.fms_als_last_info(self._h, C.byref(v)))
        return v.value

    def als_factors_device(self, which):
        ptr, rows, b, e = _vp(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.fms_als_factors_device(self._h, which, C.byref(ptr), C.byref(rows), C.byref(b), C.byref(e)))
        return int(ptr.value or 0), rows.value, b.value, e.value

    def als_planes_device(self, which):
        hi, lo, rows = _vp(), _vp(), C.c_int64()
        self._check(self._lib.fms_als_planes_device(self._h, which, C.byref(hi), C.byref(lo), C.byref(rows)))
        return int(hi.value or 0), int(lo.value or 0), rows.value

    def als_planes_ipc_handle(self, which) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._lib.fms_als_planes_ipc_handle(self._h, which, buf))
        return buf.raw

    def als_planes_open_peers(self, which, handles: bytes, n: int):
        self._check(self._lib.fms_als_planes_open_peers(self._h, which, C.c_char_p(handles) if n else None, n))

    def svm_exchange_ipc_handle(self, n_ranks: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._lib.fms_svm_exchange_ipc_handle(self._h, n_ranks, buf))
        return buf.raw

    def svm_exchange_open_peers(self, handles: bytes, n_ranks: int, my_rank: int):
        assert len(handles) == 64 * n_ranks
        self._check(self._lib.fms_svm_exchange_open_peers(self._h, C.create_string_buffer(handles, len(handles)), n_ranks, my_rank))

    def svm_round_exchange(self, local_iterations: int = -1):
        self._check(self._lib.fms_svm_round_exchange(self._h, local_iterations))

    def svm_exchange_status(self):
        t = C.c_int(0)
        self._check(self._lib.fms_svm_exchange_status(self._h, C.byref(t)))

    def als_planes_flag_sync(self, enable: bool):
        self._check(self._lib.fms_als_planes_flag_sync(self._h, 1 if enable else 0))

    def als_planes_peer_count(self, which) -> int:
        n = C.c_int()
        self._check(self._lib.fms_als_planes_peer_count(self._h, which, C.byref(n)))
        return n.value

    def als_planes_updated(self, which):
        self._check(self._lib.fms_als_planes_updated(self._h, which))

    def als_factors_updated(self, which):
        self._check(self._lib.fms_als_factors_updated(self._h, which))

    def als_get_factors(self, which, ids_out=None, factors_out=None):
        """ids_out / factors_out: optional caller buffers (e.g. numpy views of pinned memory, the analogue of a
        registered direct ByteBuffer) of n int64 / n*rank float64; fresh pageable arrays otherwise."""
        nu, ni, rank = self._als_shape
        n = nu if which == USERS else ni
        ids = np.empty(n, np.int64) if ids_out is None else ids_out
        fac = np.empty((n, rank), np.float64) if factors_out is None else factors_out
        if ids.dtype != np.int64 or fac.dtype != np.float64 or ids.size < n or fac.size < n * rank or not fac.flags.c_contiguous:
            raise ValueError("als_get_factors: output buffers have the wrong dtype / size")
        self._check(self._lib.fms_als_get_factors(self._h, which, _p(ids, _i64p), _p(fac, _f64p)))
        return ids, fac

    def als_set_model(self, user_ids, user_factors, item_ids, item_factors):
        ui, ii = _i64(user_ids), _i64(item_ids)
        uf, itf = _f64(user_factors), _f64(item_factors)
        rank = uf.shape[1] if uf.ndim == 2 and uf.shape[0] else (itf.shape[1] if itf.ndim == 2 else 0)
        self._check(self._lib.fms_als_set_model(self._h, _p(ui, _i64p), _p(uf, _f64p), len(ui), _p(ii, _i64p),
                                                _p(itf, _f64p), len(ii), rank))
        self._als_shape = (len(ui), len(ii), rank)

    # ---- N4 ingest
    def parse_ratings_csv(self, text: bytes, delimiter: str = ",", ignore_first_line: bool = False):
        """env.readCsvFile[(Int, Int, Double)] on the device: bytes -> (user int64[], item int64[], rating float64[])."""
        n = C.c_int64()
        d = delimiter.encode()
        if len(d) != 1:
            raise ValueError("single-byte field delimiter expected")
        self._check(self._lib.fms_parse_ratings_csv(self._h, text, len(text), d, int(ignore_first_line), None, None, None, 0, C.byref(n)))
        u = np.empty(n.value, np.int64); i = np.empty(n.value, np.int64); r = np.empty(n.value, np.float64)
        self._check(self._lib.fms_parse_ratings_csv(self._h, text, len(text), d, int(ignore_first_line), _p(u, _i64p), _p(i, _i64p),
                                                    _p(r, _f64p), n.value, C.byref(n)))
        return u, i, r

    def parse_libsvm(self, text: bytes):
        """env.readLibSVM on the device: bytes -> (label float64[], (ptr int64[], idx int32[], val float64[]), dim)."""
        n, nnz, dim = C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self._lib.fms_parse_libsvm(self._h, text, len(text), None, 0, None, None, None, 0, C.byref(n), C.byref(nnz), C.byref(dim)))
        y = np.empty(n.value, np.float64); ptr = np.empty(n.value + 1, np.int64)
        idx = np.empty(nnz.value, np.int32); val = np.empty(nnz.value, np.float64)
        self._check(self._lib.fms_parse_libsvm(self._h, text, len(text), _p(y, _f64p), n.value, _p(ptr, _i64p), _p(idx, _i32p), _p(val, _f64p),
                                               nnz.value, C.byref(n), C.byref(nnz), C.byref(dim)))
        return y, (ptr, idx, val), dim.value

    def als_sgd_step(self, user, item, rating, learning_rate=0.1, user_reg=0.0, item_reg=0.0, user_mean=None, item_mean=None, apply=False):
        """One micro-batch of the als-ms online SGD (SGD.java:160-217): returns (updated_user_rows, updated_item_rows)."""
        u, i, r = _i64(user), _i64(item), _f64(rating)
        if not (len(u) == len(i) == len(r)):
            raise ValueError("user/item/rating lengths differ")
        rank = self._als_shape[2]
        um = None if user_mean is None else _f64(user_mean)
        im = None if item_mean is None else _f64(item_mean)
        if (um is not None and len(um) != rank) or (im is not None and len(im) != rank):
            raise ValueError("mean vectors must have `rank` entries")
        ou = np.empty((len(u), rank), np.float64)
        oi = np.empty((len(u), rank), np.float64)
        self._check(self._lib.fms_als_sgd_step(self._h, _p(u, _i64p), _p(i, _i64p), _p(r, _f64p), len(u), learning_rate, user_reg, item_reg,
                                               None if um is None else _p(um, _f64p), None if im is None else _p(im, _f64p),
                                               1 if apply else 0, _p(ou, _f64p), _p(oi, _f64p)))
        return ou, oi

    def als_predict(self, u, i, out=None, found=None):
        """Batched lookup-and-dot.  `out` (float64) / `found` (uint8) may be caller-owned buffers, e.g. pinned host memory
        (what a JVM host passes as a direct ByteBuffer); they are then returned as is (found as uint8)."""
        u, i = _i64(u), _i64(i)
        if len(u) != len(i):
            raise ValueError("u/i lengths differ")
        own = found is None
        if out is None:
            out = np.empty(len(u), np.float64)
        if found is None:
            found = np.empty(len(u), np.uint8)
        if out.dtype != np.float64 or found.dtype != np.uint8 or len(out) < len(u) or len(found) < len(u):
            raise ValueError("out must be float64 and found uint8, both at least len(u) long")
        self._check(self._lib.fms_als_predict(self._h, _p(u, _i64p), _p(i, _i64p), len(u), _p(out, _f64p), _p(found, _u8p)))
        return out, (found.astype(bool) if own else found)

    def als_mse(self, u, i, rating):
        u, i, r = _i64(u), _i64(i), _f64(rating)
        mse, used = C.c_double(), C.c_int64()
        self._check(self._lib.fms_als_mse(self._h, _p(u, _i64p), _p(i, _i64p), _p(r, _f64p), len(u), C.byref(mse), C.byref(used)))
        return mse.value, used.value

    # ---- SVM
    @staticmethod
    def _xargs(X, csr):
        if csr is not None:
            ptr, idx, val = csr
            ptr = np.ascontiguousarray(ptr, np.int64)
            idx = np.ascontiguousarray(idx, np.int32)
            val = np.ascontiguousarray(val, np.float64)
            return (ptr, idx, val), val.ctypes.data_as(_vp), X_CSR_F64, _p(ptr, _i64p), _p(idx, _i32p)
        if isinstance(X, np.ndarray) and X.dtype == np.float32:
            X = np.ascontiguousarray(X)
            return (X,), X.ctypes.data_as(_vp), X_DENSE_F32, None, None
        X = np.ascontiguousarray(X, np.float64)
        return (X,), X.ctypes.data_as(_vp), X_DENSE_F64, None, None

    def svm_fit(self, X, y, block_id, blocks, iterations, local_iterations, regularization=1.0, stepsize=1.0, seed=0,
                rng_mode=RNG_PER_BLOCK, csr=None, dim=None):
        keep, xp, dt, pp, ip = self._xargs(X, csr)
        y = _f64(y)
        dim = int(dim if dim is not None else X.shape[1])
        b = np.ascontiguousarray(block_id, np.int32)
        w = np.empty(dim, np.float64)
        self._check(self._lib.fms_svm_fit(self._h, xp, dt, pp, ip, _p(y, _f64p), len(y), dim, _p(b, _i32p), blocks,
                                          iterations, local_iterations, regularization, stepsize, seed, rng_mode, _p(w, _f64p)))
        self._svm_n = len(y)
        return w

    def svm_fit_stream(self, X, y, block_id, blocks, iterations, local_iterations, regularization=1.0, stepsize=1.0, seed=0,
                       rng_mode=RNG_PER_BLOCK, chunk_rows=1024):
        """fms_svm_fit for dense rows fed through the chunked-upload entry points."""
        X = np.ascontiguousarray(X if X.dtype == np.float32 else np.asarray(X, np.float64))
        dt = X_DENSE_F32 if X.dtype == np.float32 else X_DENSE_F64
        y = _f64(y)
        b = np.ascontiguousarray(block_id, np.int32)
        n, dim = X.shape
        self._check(self._lib.fms_svm_stream_begin(self._h, n, dim))
        for s in range(0, n, chunk_rows):
            e = min(n, s + chunk_rows)
            xs = np.ascontiguousarray(X[s:e])
            self._check(self._lib.fms_svm_stream_append(self._h, xs.ctypes.data_as(_vp), dt, _p(y[s:e], _f64p), _p(b[s:e], _i32p), e - s))
        w = np.empty(dim, np.float64)
        self._check(self._lib.fms_svm_stream_fit(self._h, blocks, iterations, local_iterations, regularization, stepsize, seed, rng_mode,
                                                 _p(w, _f64p)))
        self._svm_n = n
        return w

    def svm_load(self, X, y, block_id, blocks, shard_index=0, shard_count=1, regularization=1.0, stepsize=1.0, seed=0,
                 rng_mode=RNG_PER_BLOCK, csr=None, dim=None, device_ptr=None, n=None):
        """device_ptr: raw device address of a dense fp32 row-major X already on this ctx's GPU (used in place)."""
        y = _f64(y)
        b = np.ascontiguousarray(block_id, np.int32)
        if device_ptr is not None:
            xp, dt, pp, ip, isdev = _vp(device_ptr), X_DENSE_F32, None, None, 1
            keep = ()
        else:
            keep, xp, dt, pp, ip = self._xargs(X, csr)
            isdev = 0
            dim = int(dim if dim is not None else X.shape[1])
        self._check(self._lib.fms_svm_load(self._h, xp, dt, isdev, pp, ip, _p(y, _f64p), len(y), int(dim), _p(b, _i32p), blocks,
                                           shard_index, shard_count, regularization, stepsize, seed, rng_mode))
        self._svm_n = len(y)
        self._svm_dim = int(dim)

    def svm_sweep(self, local_iterations):
        self._check(self._lib.fms_svm_sweep(self._h, local_iterations))

    def svm_apply_delta(self):
        self._check(self._lib.fms_svm_apply_delta(self._h))

    def svm_delta_device(self):
        ptr, dim = _vp(), C.c_int32()
        self._check(self._lib.fms_svm_delta_device(self._h, C.byref(ptr), C.byref(dim)))
        return int(ptr.value or 0), dim.value

    def svm_set_kernel(self, kernel):
        self._check(self._lib.fms_svm_set_kernel(self._h, kernel))

    def svm_get_weights(self, dim):
        w = np.empty(dim, np.float64)
        self._check(self._lib.fms_svm_get_weights(self._h, _p(w, _f64p), dim))
        return w

    def svm_set_weights(self, w):
        w = _f64(w)
        self._check(self._lib.fms_svm_set_weights(self._h, _p(w, _f64p), len(w)))

    def svm_get_alpha(self):
        a = np.empty(self._svm_n, np.float64)
        self._check(self._lib.fms_svm_get_alpha(self._h, _p(a, _f64p), self._svm_n))
        return a

    def svm_predict(self, X, w=None, threshold=0.0, output_decision_function=False, csr=None, dim=None, n=None):
        keep, xp, dt, pp, ip = self._xargs(X, csr)
        n = int(n if n is not None else (len(csr[0]) - 1 if csr is not None else X.shape[0]))
        dim = int(dim if dim is not None else X.shape[1])
        wv = None if w is None else _f64(w)
        out = np.empty(n, np.float64)
        self._check(self._lib.fms_svm_predict(self._h, xp, dt, pp, ip, n, dim, _p(wv, _f64p), threshold,
                                              int(bool(output_decision_function)), _p(out, _f64p)))
        return out

    def debug_als_gramian(self, which, lam, row_begin, row_end):
        rank = self._als_shape[2]
        out = np.empty((row_end - row_begin, rank * rank + rank), np.float32)
        self._check(self._lib.fms_debug_als_gramian(self._h, which, lam, row_begin, row_end,
                                                    out.ctypes.data_as(C.POINTER(C.c_float))))
        return out[:, :rank * rank].reshape(-1, rank, rank), out[:, rank * rank:]

    def debug_index_stream(self, seed, block, bound, count):
        out = np.empty(count, np.int32)
        self._check(self._lib.fms_debug_index_stream(self._h, seed, block, bound, count, _p(out, _i32p)))
        return out
