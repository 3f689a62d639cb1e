"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/wrmf_oracle.cpp header).

ctypes front-ends to
  * `liboracle_wrmf.so`      -- our CPU restatement of the reference's ALS half-iteration
  * `_ref/libref_wrmf.so`    -- the reference's own headers compiled against oracle/mini_arma
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.

Array conventions (same as the reference's arma views): a factor matrix is `k x n`
column-major, i.e. a C-contiguous numpy array of shape (n, k); the sparse matrix is CSC whose
*columns are the rows being solved for* (`ptr` int32[nc+1], `idx` int32[nnz], `val` float64[nnz]).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CHOLESKY, CONJUGATE_GRADIENT, NNLS = 0, 1, 2

_c_int_p = C.POINTER(C.c_int)
_c_dbl_p = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "liboracle_wrmf.so")
    src = os.path.join(_HERE, "wrmf_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, os.path.join(_HERE, "liboracle_wrmf.so")])
    ref = os.path.join(_HERE, "_ref", "libref_wrmf.so")
    if os.path.isdir("/root/reference/inst/include") and (force or not os.path.exists(ref)):
        subprocess.check_call(["sh", os.path.join(_HERE, "build_ref.sh")])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "liboracle_wrmf.so"))
        for name in ("oracle_als_implicit_f32", "oracle_als_implicit_f64",
                     "oracle_als_explicit_f32", "oracle_als_explicit_f64"):
            getattr(_lib, name).restype = C.c_double
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_wrmf.so"))


def ref():
    global _ref
    if _ref is None:
        build()
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libref_wrmf.so"))
        for name in ("ref_als_implicit_f32", "ref_als_implicit_f64",
                     "ref_als_explicit_f32", "ref_als_explicit_f64"):
            getattr(_ref, name).restype = C.c_double
    return _ref


def max_threads():
    return int(lib().oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(ptr, idx, val, X, Y):
    assert ptr.dtype == np.int32 and idx.dtype == np.int32 and val.dtype == np.float64
    assert X.flags.c_contiguous and Y.flags.c_contiguous and X.dtype == Y.dtype
    assert X.shape[1] == Y.shape[1] and Y.shape[0] == len(ptr) - 1
    return X.shape[1], X.shape[0], Y.shape[0], "f32" if X.dtype == np.float32 else "f64"


def gram(X, lam, n_threads=None):
    """XtX = tcrossprod(X) + lambda*I (R/model_WRMF.R:474-486)."""
    n_threads = n_threads or max_threads()
    k = X.shape[1]
    out = np.empty((k, k), dtype=X.dtype)
    fn = lib().oracle_gram_f32 if X.dtype == np.float32 else lib().oracle_gram_f64
    fn(_p(X), C.c_int(k), C.c_int(X.shape[0]), C.c_double(lam), _p(out), C.c_int(n_threads))
    return out


def als_implicit(ptr, idx, val, X, Y, XtX, lam, solver, cg_steps=3, n_threads=1, impl="oracle"):
    """One implicit half-iteration; Y updated in place; returns loss."""
    k, n_src, nc, sfx = _check(ptr, idx, val, X, Y)
    assert XtX.shape == (k, k) and XtX.dtype == X.dtype
    if impl == "oracle":
        fn = getattr(lib(), "oracle_als_implicit_" + sfx)
        return fn(C.c_int(nc), C.c_size_t(len(idx)), _p(ptr), _p(idx), _p(val), _p(X), C.c_int(k),
                  C.c_int(n_src), _p(Y), _p(XtX), C.c_double(lam), C.c_int(n_threads),
                  C.c_int(solver), C.c_int(cg_steps))
    fn = getattr(ref(), "ref_als_implicit_" + sfx)
    gbb = np.zeros(0, dtype=X.dtype)
    return fn(C.c_int(n_src), C.c_int(nc), C.c_size_t(len(idx)), _p(idx), _p(ptr), _p(val), _p(X),
              C.c_int(k), C.c_int(n_src), _p(Y), _p(np.ascontiguousarray(XtX)), C.c_int(k),
              C.c_double(lam), C.c_int(n_threads), C.c_uint(solver), C.c_uint(cg_steps), C.c_int(0),
              C.c_int(0), C.c_double(0.0), _p(gbb), C.c_int(0), C.c_int(0))


def als_explicit(ptr, idx, val, X, Y, cnt_X, lam, solver, cg_steps=3, dynamic_lambda=True,
                 n_threads=1, impl="oracle"):
    k, n_src, nc, sfx = _check(ptr, idx, val, X, Y)
    if cnt_X is None:
        cnt_X = np.zeros(n_src, dtype=X.dtype)
    cnt_X = np.ascontiguousarray(cnt_X, dtype=X.dtype)
    if impl == "oracle":
        fn = getattr(lib(), "oracle_als_explicit_" + sfx)
        return fn(C.c_int(nc), C.c_size_t(len(idx)), _p(ptr), _p(idx), _p(val), _p(X), C.c_int(k),
                  C.c_int(n_src), _p(Y), _p(cnt_X), C.c_double(lam), C.c_int(n_threads),
                  C.c_int(solver), C.c_int(cg_steps), C.c_int(int(dynamic_lambda)))
    fn = getattr(ref(), "ref_als_explicit_" + sfx)
    return fn(C.c_int(n_src), C.c_int(nc), C.c_size_t(len(idx)), _p(idx), _p(ptr), _p(val), _p(X),
              C.c_int(k), C.c_int(n_src), _p(Y), _p(cnt_X), C.c_int(len(cnt_X)), C.c_double(lam),
              C.c_int(n_threads), C.c_uint(solver), C.c_uint(cg_steps), C.c_int(int(dynamic_lambda)),
              C.c_int(0), C.c_int(0))
