"""Host-side mirrors of the reference's R wrappers `als_implicit()` / `als_explicit()`
(R/model_WRMF.R:456-515) over the stateless C-ABI calls.  Argument meaning and in-place update of
`Y` are the reference's; matrices are numpy (n, rank) C-contiguous == rank x n column-major."""
import ctypes as C

import numpy as np

from . import _lib as L


def _dense(a, dt, name):
    if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.ndim == 2):
        raise ValueError("%s must be a C-contiguous (n, rank) %s array" % (name, np.dtype(dt).name))
    return a


def gram(X, lambda_):
    """XtX = tcrossprod(X) + lambda*I on the GPU (R/model_WRMF.R:474-486)."""
    X = _dense(X, np.float32, "X")
    out = np.empty((X.shape[1], X.shape[1]), np.float32)
    L.check(L.lib().b200als_gram_float(L.vp(X), X.shape[1], X.shape[0], float(lambda_), L.vp(out)))
    return out


def als_implicit(ptr, idx, val, X, Y, lambda_, solver_code, cg_steps=3, XtX=None, n_threads=1,
                 with_user_item_bias=False, is_bias_last_row=False, global_bias=0.0):
    """One implicit-feedback half-iteration (als_implicit_{float,double}, src/wrmf_implicit.cpp:5-31).
    `Y` is modified in place; returns the loss."""
    dt = X.dtype.type
    if dt not in (np.float32, np.float64):
        raise ValueError("X must be float32 or float64")
    X = _dense(X, dt, "X")
    Y = _dense(Y, dt, "Y")
    if X.shape[1] != Y.shape[1] or Y.shape[0] != len(ptr) - 1:
        raise ValueError("shape mismatch")
    csc, keep = L.make_csc(X.shape[0], ptr, idx, val)
    if XtX is not None:
        XtX = np.ascontiguousarray(XtX, dtype=dt)
    loss = C.c_double(0.0)
    fn = L.lib().b200als_als_implicit_float if dt == np.float32 else L.lib().b200als_als_implicit_double
    L.check(fn(C.byref(csc), X.shape[1], L.vp(X), L.vp(Y), L.vp(XtX), float(lambda_), int(n_threads), int(solver_code),
               int(cg_steps), int(with_user_item_bias), int(is_bias_last_row), float(global_bias), None, 0,
               C.byref(loss)))
    del keep
    return loss.value


def als_explicit(ptr, idx, val, X, Y, cnt_X, lambda_, solver_code, cg_steps=3, dynamic_lambda=True, n_threads=1,
                 with_user_item_bias=False, is_bias_last_row=False):
    """One explicit-feedback half-iteration (als_explicit_{float,double}, src/wrmf_explicit.cpp:5-27)."""
    dt = X.dtype.type
    X = _dense(X, dt, "X")
    Y = _dense(Y, dt, "Y")
    if X.shape[1] != Y.shape[1] or Y.shape[0] != len(ptr) - 1:
        raise ValueError("shape mismatch")
    csc, keep = L.make_csc(X.shape[0], ptr, idx, val)
    if cnt_X is not None:
        cnt_X = np.ascontiguousarray(cnt_X, dtype=dt)
    loss = C.c_double(0.0)
    fn = L.lib().b200als_als_explicit_float if dt == np.float32 else L.lib().b200als_als_explicit_double
    L.check(fn(C.byref(csc), X.shape[1], L.vp(X), L.vp(Y), L.vp(cnt_X), float(lambda_), int(n_threads), int(solver_code),
               int(cg_steps), int(bool(dynamic_lambda)), int(with_user_item_bias), int(is_bias_last_row),
               C.byref(loss)))
    del keep
    return loss.value
