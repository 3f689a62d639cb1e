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
    ref2 = os.path.join(_HERE, "_ref", "libref_topk.so")
    if os.path.isdir("/root/reference/inst/include") and (force or not os.path.exists(ref) or not os.path.exists(ref2)):
        subprocess.check_call(["sh", os.path.join(_HERE, "build_ref.sh")])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "liboracle_wrmf.so"))
        for name in ("oracle_als_implicit_f32", "oracle_als_implicit_f64",
                     "oracle_als_explicit_f32", "oracle_als_explicit_f64",
                     "oracle_als_implicit_bias_f32", "oracle_als_implicit_bias_f64",
                     "oracle_als_explicit_bias_f32", "oracle_als_explicit_bias_f64",
                     "oracle_initialize_biases_f32", "oracle_initialize_biases_f64"):
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
                     "ref_als_explicit_f32", "ref_als_explicit_f64",
                     "ref_initialize_biases_f32", "ref_initialize_biases_f64"):
            getattr(_ref, name).restype = C.c_double
    return _ref


_ref_topk = None


def ref_topk_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_topk.so"))


def ref_top_product(x, y, k, nr_ptr=None, nr_idx=None, exclude_1based=(), glob_mean=0.0, n_threads=1):
    """The reference's own `top_product` (src/matrix_top_product.cpp:20-102, compiled in place by build_ref.sh).
    x: (n_user, rank), y: (n_item, rank), rows = embeddings; same return convention as oracle.topk.top_product:
    (idx 1-based int32 (n_user, k) with NA_integer_, scores float64 (n_user, k) with NaN)."""
    global _ref_topk
    if _ref_topk is None:
        build()
        _ref_topk = C.CDLL(os.path.join(_HERE, "_ref", "libref_topk.so"))
    x64 = np.asfortranarray(np.asarray(x, np.float64))                   # n_user x rank, column-major
    y64 = np.ascontiguousarray(np.asarray(y, np.float64))                # (n_item, rank) C-order == rank x n_item column-major
    n_user, rank = x64.shape
    n_item = y64.shape[0]
    if nr_ptr is None:
        nr_ptr, nr_idx = np.zeros(n_user + 1, np.int32), np.zeros(0, np.int32)
    nr_ptr = np.ascontiguousarray(nr_ptr, np.int32)
    nr_idx = np.ascontiguousarray(nr_idx, np.int32)
    nr_x = np.ones(len(nr_idx), np.float64)
    ex = np.ascontiguousarray(np.asarray(list(exclude_1based), dtype=np.int32))
    out_idx = np.empty((k, n_user), np.int32)        # column-major n_user x k
    out_sc = np.empty((k, n_user), np.float64)
    _ref_topk.ref_top_product(_p(x64), C.c_int(n_user), C.c_int(rank), _p(y64), C.c_int(n_item), C.c_uint(k),
                              C.c_uint(n_threads), _p(nr_ptr), _p(nr_idx), _p(nr_x), C.c_size_t(len(nr_idx)), _p(ex),
                              C.c_int(len(ex)), C.c_double(glob_mean), _p(out_idx), _p(out_sc))
    return np.ascontiguousarray(out_idx.T), np.ascontiguousarray(out_sc.T)


def max_threads():
    """The reference's omp_thread_count() (src/utils.cpp:84-91): honours OMP_NUM_THREADS / OMP_THREAD_LIMIT."""
    return int(lib().oracle_max_threads())


def host_threads():
    """Threads the CPU arms of bench.py use: every core this process may run on, NOT omp_get_max_threads() --
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would silently make the baseline
    single-threaded.  The half-iteration entry points take n_threads explicitly (`num_threads(n)` clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def synth_csr(n_rows, n_cols, nnz_per_row, seed=42, explicit_values=False, row_offset=0, with_values=True, n_threads=None):
    """bench.py's synthetic CSR, generated on the host by the oracle library (same entries as the product's
    b200als_synth_csr_host; see tests/test_abi.py)."""
    ptr = np.empty(n_rows + 1, np.int32)
    idx = np.empty(n_rows * nnz_per_row, np.int32)
    val = np.empty(n_rows * nnz_per_row, np.float64) if with_values else None
    rc = lib().oracle_synth_csr(C.c_int(n_rows), C.c_int(n_cols), C.c_int(nnz_per_row), C.c_ulonglong(seed),
                                C.c_int(int(explicit_values)), C.c_longlong(row_offset), _p(ptr), _p(idx),
                                _p(val) if with_values else None, C.c_int(n_threads or host_threads()))
    if rc != 0:
        raise ValueError("oracle_synth_csr: bad shape (rc %d)" % rc)
    return ptr, idx, val


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


def als_implicit_bias(ptr, idx, val, X, Y, XtX, lam, solver, cg_steps=3, with_biases=False, is_bias_last_row=False,
                      global_bias=0.0, global_bias_base=None, initialize_bias_base=True, n_threads=1, impl="oracle"):
    """Implicit half-iteration with user/item and/or global bias (wrmf_implicit.hpp:90-305, every branch).
    X, Y have rank+2 columns when with_biases; XtX is (rank+1)^2 then.  global_bias_base (length = XtX side) is
    updated in place when initialize_bias_base.  Returns loss."""
    k, n_src, nc, sfx = _check(ptr, idx, val, X, Y)
    ks = k - int(bool(with_biases))
    assert XtX.shape == (ks, ks) and XtX.dtype == X.dtype
    if global_bias_base is None:
        global_bias_base = np.zeros(ks, dtype=X.dtype)
    assert global_bias_base.dtype == X.dtype and len(global_bias_base) >= (0 if with_biases else ks)
    if impl == "oracle":
        fn = getattr(lib(), "oracle_als_implicit_bias_" + sfx)
        return fn(C.c_int(nc), C.c_size_t(len(idx)), _p(ptr), _p(idx), _p(val), _p(X), C.c_int(k), C.c_int(n_src),
                  _p(Y), _p(XtX), C.c_double(lam), C.c_int(n_threads), C.c_int(solver), C.c_int(cg_steps),
                  C.c_int(int(with_biases)), C.c_int(int(is_bias_last_row)), C.c_double(global_bias),
                  _p(global_bias_base), C.c_int(int(initialize_bias_base)))
    fn = getattr(ref(), "ref_als_implicit_" + sfx)
    return fn(C.c_int(n_src), C.c_int(nc), C.c_size_t(len(idx)), _p(idx), _p(ptr), _p(val), _p(X),
              C.c_int(k), C.c_int(n_src), _p(Y), _p(np.ascontiguousarray(XtX)), C.c_int(ks),
              C.c_double(lam), C.c_int(n_threads), C.c_uint(solver), C.c_uint(cg_steps), C.c_int(int(with_biases)),
              C.c_int(int(is_bias_last_row)), C.c_double(global_bias), _p(global_bias_base),
              C.c_int(len(global_bias_base)), C.c_int(int(initialize_bias_base)))


def als_explicit_bias(ptr, idx, val, X, Y, cnt_X, lam, solver, cg_steps=3, dynamic_lambda=True, with_biases=False,
                      is_bias_last_row=False, n_threads=1, impl="oracle"):
    """Explicit half-iteration with user/item biases (wrmf_explicit.hpp:33-174, every branch)."""
    k, n_src, nc, sfx = _check(ptr, idx, val, X, Y)
    if cnt_X is None:
        cnt_X = np.zeros(n_src, dtype=X.dtype)
    cnt_X = np.ascontiguousarray(cnt_X, dtype=X.dtype)
    if impl == "oracle":
        fn = getattr(lib(), "oracle_als_explicit_bias_" + sfx)
        return fn(C.c_int(nc), C.c_size_t(len(idx)), _p(ptr), _p(idx), _p(val), _p(X), C.c_int(k), C.c_int(n_src),
                  _p(Y), _p(cnt_X), C.c_double(lam), C.c_int(n_threads), C.c_int(solver), C.c_int(cg_steps),
                  C.c_int(int(dynamic_lambda)), C.c_int(int(with_biases)), C.c_int(int(is_bias_last_row)))
    fn = getattr(ref(), "ref_als_explicit_" + sfx)
    return fn(C.c_int(n_src), C.c_int(nc), C.c_size_t(len(idx)), _p(idx), _p(ptr), _p(val), _p(X),
              C.c_int(k), C.c_int(n_src), _p(Y), _p(cnt_X), C.c_int(len(cnt_X)), C.c_double(lam),
              C.c_int(n_threads), C.c_uint(solver), C.c_uint(cg_steps), C.c_int(int(dynamic_lambda)),
              C.c_int(int(with_biases)), C.c_int(int(is_bias_last_row)))


def initialize_biases(csc, csr, user_bias, item_bias, lam, dynamic_lambda, non_negative, calculate_global_bias,
                      is_explicit, impl="oracle"):
    """initialize_biases<T> (wrmf_utils.hpp:170-183).  csc = (ptr[n_item+1], idx(users), val) of the user x item
    matrix, csr = (ptr[n_user+1], idx(items), val) of the same entries by user; val arrays (float64) are modified in
    place for explicit feedback with calculate_global_bias.  Biases are filled in place; returns global_bias."""
    cp, ci, cv = csc
    rp, ri, rv = csr
    n_items, n_users = len(cp) - 1, len(rp) - 1
    assert user_bias.dtype == item_bias.dtype and len(user_bias) == n_users and len(item_bias) == n_items
    sfx = "f32" if user_bias.dtype == np.float32 else "f64"
    tail = (_p(user_bias), _p(item_bias), C.c_double(lam), C.c_int(int(dynamic_lambda)), C.c_int(int(non_negative)),
            C.c_int(int(calculate_global_bias)), C.c_int(int(is_explicit)))
    if impl == "oracle":
        fn = getattr(lib(), "oracle_initialize_biases_" + sfx)
        return fn(C.c_int(n_items), C.c_int(n_users), C.c_size_t(len(ci)), _p(cp), _p(ci), _p(cv), _p(rp), _p(ri),
                  _p(rv), *tail)
    fn = getattr(ref(), "ref_initialize_biases_" + sfx)
    return fn(C.c_int(n_users), C.c_int(n_items), C.c_size_t(len(ci)), _p(ci), _p(cp), _p(cv), _p(ri), _p(rp), _p(rv),
              *tail)
