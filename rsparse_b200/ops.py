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
                 with_user_item_bias=False, is_bias_last_row=False, global_bias=0.0, global_bias_base=None,
                 initialize_bias_base=True):
    """One implicit-feedback half-iteration (als_implicit_{float,double}, src/wrmf_implicit.cpp:5-31; R wrapper
    R/model_WRMF.R:456-497).  `Y` is modified in place; returns the loss.  With `with_user_item_bias`, X and Y have
    rank+2 columns and XtX is (rank+1) x (rank+1); `global_bias_base` (length = side of XtX, dtype of X) is rewritten
    when `initialize_bias_base`, as in the reference."""
    dt = X.dtype.type
    if dt not in (np.float32, np.float64):
        raise ValueError("X must be float32 or float64")
    X = _dense(X, dt, "X")
    Y = _dense(Y, dt, "Y")
    if X.shape[1] != Y.shape[1] or Y.shape[0] != len(ptr) - 1:
        raise ValueError("shape mismatch")
    csc, keep = L.make_csc(X.shape[0], ptr, idx, val)
    ks = X.shape[1] - int(bool(with_user_item_bias))
    if XtX is not None:
        XtX = np.ascontiguousarray(XtX, dtype=dt)
        if XtX.shape != (ks, ks):
            raise ValueError("XtX must be %d x %d" % (ks, ks))
    if global_bias_base is None:
        global_bias_base = np.zeros(ks, dt)
    if not (isinstance(global_bias_base, np.ndarray) and global_bias_base.dtype == dt and global_bias_base.flags.c_contiguous
            and (with_user_item_bias or len(global_bias_base) == ks)):
        raise ValueError("global_bias_base must be a contiguous %s vector of length %d" % (np.dtype(dt).name, ks))
    loss = C.c_double(0.0)
    fn = L.lib().b200als_als_implicit_float if dt == np.float32 else L.lib().b200als_als_implicit_double
    L.check(fn(C.byref(csc), X.shape[1], L.vp(X), L.vp(Y), L.vp(XtX), float(lambda_), int(n_threads), int(solver_code),
               int(cg_steps), int(bool(with_user_item_bias)), int(bool(is_bias_last_row)), float(global_bias),
               L.vp(global_bias_base), int(bool(initialize_bias_base)), C.byref(loss)))
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


def initialize_biases(c_ui, c_iu, user_bias, item_bias, lambda_, dynamic_lambda, non_negative, calculate_global_bias,
                      is_explicit_feedback):
    """initialize_biases_{double,float} (src/wrmf_init.cpp:6-34 -> inst/include/wrmf_utils.hpp:170-183) on the GPU.
    c_ui = (ptr[n_item+1], idx, val) the user x item matrix by item, c_iu = (ptr[n_user+1], idx, val) by user; the
    float64 `val` arrays are shifted IN PLACE for explicit feedback with calculate_global_bias; the bias vectors are
    filled in place.  Returns global_bias."""
    cp, ci, cv = c_ui
    rp, ri, rv = c_iu
    dt = user_bias.dtype.type
    for a in (cp, ci, rp, ri):
        if a.dtype != np.int32 or not a.flags.c_contiguous:
            raise ValueError("index arrays must be contiguous int32")
    for a in (cv, rv):
        if a.dtype != np.float64 or not a.flags.c_contiguous:
            raise ValueError("values must be contiguous float64")
    n_item, n_user = len(cp) - 1, len(rp) - 1
    if item_bias.dtype != dt or len(user_bias) != n_user or len(item_bias) != n_item or dt not in (np.float32, np.float64):
        raise ValueError("bias vectors do not match the matrix")
    g = C.c_double(0.0)
    fn = L.lib().b200als_initialize_biases_float if dt == np.float32 else L.lib().b200als_initialize_biases_double
    L.check(fn(n_user, n_item, len(ci), L.vp(cp), L.vp(ci), L.vp(cv), L.vp(rp), L.vp(ri), L.vp(rv), L.vp(user_bias),
               L.vp(item_bias), float(lambda_), int(bool(dynamic_lambda)), int(bool(non_negative)),
               int(bool(calculate_global_bias)), int(bool(is_explicit_feedback)), C.byref(g)))
    return g.value


def top_product(user_emb, item_emb, k, not_recommend=None, exclude=(), glob_mean=0.0):
    """Top-k items per user (mirror of find_top_product / top_product, R/utils.R:31-59,
    src/matrix_top_product.cpp:20-102) on the GPU.  `user_emb` (n_user, rank) and `item_emb` (n_item, rank) float32;
    `not_recommend`: scipy sparse (n_user x n_item) or None; `exclude`: 0-based item ids excluded for everyone.
    Returns (indices int32 (n_user, k), 0-based, -1 = NA; scores float64 (n_user, k), NaN = NA)."""
    import scipy.sparse as sp
    user_emb = _dense(np.ascontiguousarray(user_emb, dtype=np.float32), np.float32, "user_emb")
    item_emb = _dense(np.ascontiguousarray(item_emb, dtype=np.float32), np.float32, "item_emb")
    n_user, rank = user_emb.shape
    n_item = item_emb.shape[0]
    if item_emb.shape[1] != rank:
        raise ValueError("ncol(x) == nrow(y) is not TRUE")                 # R/utils.R:50
    ptr = idx = None
    if not_recommend is not None:
        nr = sp.csr_matrix(not_recommend)
        if nr.shape != (n_user, n_item):
            raise ValueError("not_recommend must be n_user x n_item")       # R/utils.R:55-56
        nr.sort_indices()
        ptr, idx = nr.indptr.astype(np.int32), nr.indices.astype(np.int32)
    ex = np.ascontiguousarray(np.unique(np.asarray(exclude, dtype=np.int64)) + 1, dtype=np.int32)   # R is 1-based
    if len(ex) and ex.max() > n_item:
        raise ValueError("some of items_exclude indices are bigger than number of items")
    out_idx = np.empty((k, n_user), np.int32)      # column-major n_user x k
    out_sc = np.empty((k, n_user), np.float64)
    L.check(L.lib().b200als_top_product(L.vp(user_emb), n_user, L.vp(item_emb), n_item, rank, int(k), L.vp(ptr), L.vp(idx),
                                        L.vp(ex) if len(ex) else None, len(ex), float(glob_mean), L.vp(out_idx),
                                        L.vp(out_sc)))
    idx0 = out_idx.T.astype(np.int64)
    idx0 = np.where(idx0 == -2147483648, -1, idx0 - 1).astype(np.int32)
    return idx0, np.ascontiguousarray(out_sc.T)
