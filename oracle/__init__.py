"""CPU oracle for the flink-ms hot paths -- TEST INFRASTRUCTURE ONLY (parity unpinned, see
flinkml_oracle.c header).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product never does.

ctypes binding of oracle/liboracle.so (built by oracle/Makefile from flinkml_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (explicit `gcc`, the image's $CC wrapper lacks libgomp.spec)."""
    src = os.path.join(_HERE, "flinkml_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_als_create.restype = C.c_void_p
        L.orc_als_create.argtypes = [_i64p, _i64p, _f64p, C.c_int64]
        L.orc_als_destroy.argtypes = [C.c_void_p]
        L.orc_als_n_users.restype = C.c_int64
        L.orc_als_n_users.argtypes = [C.c_void_p]
        L.orc_als_n_items.restype = C.c_int64
        L.orc_als_n_items.argtypes = [C.c_void_p]
        L.orc_als_init.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        L.orc_als_half_step.restype = C.c_int64
        L.orc_als_half_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.orc_als_half_step_range.restype = C.c_int64
        L.orc_als_half_step_range.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_int]
        L.orc_als_fit.restype = C.c_int64
        L.orc_als_fit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int]
        L.orc_als_get_factors.argtypes = [C.c_void_p, C.c_int, _i64p, _f64p]
        L.orc_als_set_factors.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p]
        L.orc_als_predict.argtypes = [C.c_void_p, _i64p, _i64p, C.c_int64, _f64p, _u8p]
        L.orc_als_risk.restype = C.c_double
        L.orc_als_risk.argtypes = [C.c_void_p, C.c_double]
        L.orc_svm_fit.restype = C.c_int
        L.orc_svm_fit.argtypes = [C.c_void_p, C.c_int, _i64p, _i32p, _f64p, C.c_int64, C.c_int32, _i32p, C.c_int,
                                  C.c_int, C.c_int64, C.c_double, C.c_double, C.c_int64, C.c_int, C.c_int, _f64p, _f64p]
        L.orc_svm_index_stream.argtypes = [C.c_int64, C.c_int, C.c_int32, C.c_int64, _i32p]
        L.orc_svm_predict.argtypes = [C.c_void_p, C.c_int, _i64p, _i32p, C.c_int64, C.c_int32, _f64p, C.c_double,
                                      C.c_int, _f64p]
        L.orc_svm_objectives.argtypes = [C.c_void_p, C.c_int, _i64p, _i32p, _f64p, C.c_int64, C.c_int32, _f64p, _f64p,
                                         C.c_double, _f64p, _f64p]
        L.orc_jrand_ints.argtypes = [C.c_int64, C.c_int, _i32p]
        L.orc_jrand_bounded.argtypes = [C.c_int64, C.c_int32, C.c_int, _i32p]
        L.orc_jrand_doubles.argtypes = [C.c_int64, C.c_int, _f64p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_switch.argtypes = [C.c_char_p, C.c_int]
        L.orc_get_switch.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# ---------------------------------------------------------------- java.util.Random
def jrand_ints(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    lib().orc_jrand_ints(seed, n, _p(out, _i32p))
    return out


def jrand_bounded(seed: int, bound: int, n: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    lib().orc_jrand_bounded(seed, bound, n, _p(out, _i32p))
    return out


def jrand_doubles(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float64)
    lib().orc_jrand_doubles(seed, n, _p(out, _f64p))
    return out


SWITCHES = ("als_lambda_times_n", "als_final_user_step", "svm_scaling_configured_blocks", "svm_alpha_scale_stepsize")


def set_switch(name: str, value: int) -> None:
    """Risk-register switch of the restatement (flinkml_oracle.c header): recollected details of FlinkML 1.3.1 that change
    results.  Process-global; tests restore the default (1) after use."""
    if lib().orc_set_switch(name.encode(), int(value)) != 0:
        raise KeyError(name)


def get_switch(name: str) -> int:
    v = lib().orc_get_switch(name.encode())
    if v < 0:
        raise KeyError(name)
    return v


def num_threads() -> int:
    return int(lib().orc_num_threads())


# ---------------------------------------------------------------- ALS
class OracleALS:
    """fp64 FlinkML-ALS restatement (handle over orc_als_*)."""

    def __init__(self, user, item, rating):
        self._u = np.ascontiguousarray(user, np.int64)
        self._i = np.ascontiguousarray(item, np.int64)
        self._r = np.ascontiguousarray(rating, np.float64)
        self._h = lib().orc_als_create(_p(self._u, _i64p), _p(self._i, _i64p), _p(self._r, _f64p), len(self._u))
        self.n_users = int(lib().orc_als_n_users(self._h))
        self.n_items = int(lib().orc_als_n_items(self._h))
        self.rank = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_als_destroy(self._h)
            self._h = None

    def init(self, rank, seed):
        self.rank = rank
        lib().orc_als_init(self._h, rank, seed)

    def half_step(self, which, lam, nthreads=0, rows=None):
        if rows is None:
            return int(lib().orc_als_half_step(self._h, which, lam, nthreads))
        return int(lib().orc_als_half_step_range(self._h, which, lam, rows[0], rows[1], nthreads))

    def fit(self, rank, iterations, lam, seed, nthreads=0):
        self.rank = rank
        return int(lib().orc_als_fit(self._h, rank, iterations, lam, seed, nthreads))

    def factors(self, which):
        n = self.n_users if which == 0 else self.n_items
        ids = np.empty(n, np.int64)
        f = np.empty((n, self.rank), np.float64)
        lib().orc_als_get_factors(self._h, which, _p(ids, _i64p), _p(f, _f64p))
        return ids, f

    def set_factors(self, which, f):
        f = np.ascontiguousarray(f, np.float64)
        self.rank = f.shape[1]
        lib().orc_als_set_factors(self._h, which, self.rank, _p(f, _f64p))

    def predict(self, u, i):
        u = np.ascontiguousarray(u, np.int64)
        i = np.ascontiguousarray(i, np.int64)
        out = np.empty(len(u), np.float64)
        found = np.empty(len(u), np.uint8)
        lib().orc_als_predict(self._h, _p(u, _i64p), _p(i, _i64p), len(u), _p(out, _f64p), _p(found, _u8p))
        return out, found.astype(bool)

    def risk(self, lam):
        return float(lib().orc_als_risk(self._h, lam))


# ---------------------------------------------------------------- SVM
def _xargs(X, csr):
    if csr is not None:
        ptr, idx, val = csr
        ptr = np.ascontiguousarray(ptr, np.int64)
        idx = np.ascontiguousarray(idx, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        return (val, ptr, idx), val.ctypes.data_as(C.c_void_p), 2, _p(ptr, _i64p), _p(idx, _i32p)
    if X.dtype == np.float32:
        X = np.ascontiguousarray(X)
        return (X,), X.ctypes.data_as(C.c_void_p), 0, None, None
    X = np.ascontiguousarray(X, np.float64)
    return (X,), X.ctypes.data_as(C.c_void_p), 1, None, None


def svm_fit(X, y, block_id, blocks, iterations, local_iterations, regularization=1.0, stepsize=1.0, seed=0,
            rng_mode=0, nthreads=0, csr=None, dim=None, return_alpha=False):
    keep, xp, dt, pp, ip = _xargs(X, csr)
    y = np.ascontiguousarray(y, np.float64)
    n = len(y)
    dim = int(dim if dim is not None else X.shape[1])
    b = np.ascontiguousarray(block_id, np.int32)
    w = np.empty(dim, np.float64)
    alpha = np.empty(n, np.float64) if return_alpha else None
    rc = lib().orc_svm_fit(xp, dt, pp, ip, _p(y, _f64p), n, dim, _p(b, _i32p), blocks, iterations, local_iterations,
                           regularization, stepsize, seed, rng_mode, nthreads, _p(w, _f64p), _p(alpha, _f64p))
    if rc != 0:
        raise ValueError("orc_svm_fit: bad arguments")
    return (w, alpha) if return_alpha else w


def svm_index_stream(seed, block, nk, count):
    out = np.empty(count, np.int32)
    lib().orc_svm_index_stream(seed, block, nk, count, _p(out, _i32p))
    return out


def svm_predict(X, w, threshold=0.0, output_decision_function=False, csr=None, dim=None, n=None):
    keep, xp, dt, pp, ip = _xargs(X, csr)
    w = np.ascontiguousarray(w, np.float64)
    n = int(n if n is not None else (len(csr[0]) - 1 if csr is not None else X.shape[0]))
    out = np.empty(n, np.float64)
    lib().orc_svm_predict(xp, dt, pp, ip, n, len(w), _p(w, _f64p), threshold, int(output_decision_function),
                          _p(out, _f64p))
    return out


def svm_objectives(X, y, w, alpha, lam, csr=None):
    keep, xp, dt, pp, ip = _xargs(X, csr)
    y = np.ascontiguousarray(y, np.float64)
    w = np.ascontiguousarray(w, np.float64)
    alpha = np.ascontiguousarray(alpha, np.float64)
    p, d = C.c_double(), C.c_double()
    lib().orc_svm_objectives(xp, dt, pp, ip, _p(y, _f64p), len(y), len(w), _p(w, _f64p), _p(alpha, _f64p), lam,
                             C.byref(p), C.byref(d))
    return p.value, d.value
