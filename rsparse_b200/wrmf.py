"""Host-side mirror of the reference's R6 class `WRMF` (R/model_WRMF.R:35-454) re-bound to
libb200als.so.  Same constructor arguments, `fit_transform` / `transform` / `components` semantics
and error behaviour; the ALS loop of R/model_WRMF.R:318-338 runs either through the stateless
calls (precision="double": R's default, fp64 kernels, one call per half-iteration exactly like
the reference) or inside a device-resident session (precision="float": the fp32 engine).
In an image with R the same ABI is bound by R/ + src/b200als_shim.c."""
import ctypes as C
import logging

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from .ops import als_explicit, als_implicit, initialize_biases

logger = logging.getLogger("rsparse")

_SOLVER_CODES = {"cholesky": 0, "conjugate_gradient": 1, "nnls": 2}  # R/model_WRMF.R:99-100


def _targets_csc(m_targets_by_src):
    """scipy (targets x src) -> (ptr, idx, val): the CSC whose columns are the rows to solve."""
    m = sp.csr_matrix(m_targets_by_src)
    m.sort_indices()
    return m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.astype(np.float64)


class Session:
    """Thin RAII wrapper over the session API of include/b200als.h."""

    def __init__(self, c_ui, c_iu, n_user, n_item, rank, feedback, solver, cg_steps=3, dynamic_lambda=True,
                 lambda_=0.0, kernel=0, stage=0, ctas=0, gram=0):
        self._h = C.c_void_p(None)
        o = L.Options()
        L.lib().b200als_default_options(C.byref(o))
        o.feedback = L.IMPLICIT if feedback == "implicit" else L.EXPLICIT
        o.solver = int(solver)
        o.cg_steps = int(cg_steps)
        o.dynamic_lambda = int(bool(dynamic_lambda))
        o.lambda_ = float(lambda_)
        o.kernel = int(kernel)
        o.reserved[0] = int(stage)   # resident-kernel tile staging: 0 default, 1 cp.async.bulk, 2 cp.async
        o.reserved[1] = int(ctas)    # resident-kernel CTAs/SM: 0 default, 3, 4
        o.reserved[2] = int(gram)    # XtX arithmetic: 0 default, 1 bf16 tensor-core, 2 fp32 FMA, 3 3xTF32
        self.rank, self.n_user, self.n_item = int(rank), int(n_user), int(n_item)
        keep = []
        s_ui = s_iu = None
        if c_ui is not None:
            s_ui, k1 = L.make_csc(n_user, *c_ui)
            keep.append(k1)
        if c_iu is not None:
            s_iu, k2 = L.make_csc(n_item, *c_iu)
            keep.append(k2)
        L.check(L.lib().b200als_create(C.byref(self._h), C.byref(s_ui) if s_ui is not None else None,
                                       C.byref(s_iu) if s_iu is not None else None, self.n_user, self.n_item,
                                       self.rank, C.byref(o)))

    @classmethod
    def synthetic(cls, n_user_local, user_offset, n_user_global, n_item, nnz_per_row, seed, rank, feedback="implicit",
                  solver=L.CONJUGATE_GRADIENT, cg_steps=3, dynamic_lambda=True, lambda_=0.1, kernel=0, stage=0, ctas=0,
                  col_dist=0, len_dist=0, gram=0):
        self = cls.__new__(cls)
        self._h = C.c_void_p(None)
        o = L.Options()
        L.lib().b200als_default_options(C.byref(o))
        o.feedback = L.IMPLICIT if feedback == "implicit" else L.EXPLICIT
        o.solver, o.cg_steps, o.dynamic_lambda, o.lambda_, o.kernel = int(solver), int(cg_steps), int(dynamic_lambda), float(lambda_), int(kernel)
        o.reserved[0] = int(stage)
        o.reserved[1] = int(ctas)
        o.reserved[2] = int(gram)
        self.rank, self.n_user, self.n_item = int(rank), int(n_user_global), int(n_item)
        L.check(L.lib().b200als_create_synthetic_ex(C.byref(self._h), int(n_user_local), int(user_offset), int(n_user_global),
                                                    int(n_item), int(nnz_per_row), int(seed), int(rank), C.byref(o),
                                                    int(col_dist), int(len_dist)))
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().b200als_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close

    def build_missing_orientation(self):
        L.check(L.lib().b200als_build_missing_orientation(self._h))

    def get_orientation(self, which):
        nnz = C.c_int64(0)
        L.check(L.lib().b200als_get_orientation(self._h, which, None, None, None, C.byref(nnz)))
        n_cols = self.n_item if which == L.ITEMS else self.n_user
        ptr = np.empty(n_cols + 1, np.int32)
        idx = np.empty(nnz.value, np.int32)
        val = np.empty(nnz.value, np.float32)
        L.check(L.lib().b200als_get_orientation(self._h, which, L.vp(ptr), L.vp(idx), L.vp(val), None))
        return ptr, idx, val

    def set_factors(self, which, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        n = self.n_item if which == L.ITEMS else self.n_user
        assert a.shape == (n, self.rank), (a.shape, (n, self.rank))
        L.check(L.lib().b200als_set_factors(self._h, which, L.vp(a)))

    def get_factors(self, which):
        n = self.n_item if which == L.ITEMS else self.n_user
        out = np.empty((n, self.rank), np.float32)
        L.check(L.lib().b200als_get_factors(self._h, which, L.vp(out)))
        return out

    def init_factors(self, seed):
        L.check(L.lib().b200als_init_factors(self._h, int(seed)))

    def randomize_factors(self, which, seed, scale=0.01, decay=0.0):
        L.check(L.lib().b200als_randomize_factors(self._h, which, int(seed), float(scale), float(decay)))

    def set_shard(self, which, begin, end):
        L.check(L.lib().b200als_set_shard(self._h, which, int(begin), int(end)))

    def half_iteration(self, which, solver_override=-1):
        loss = C.c_double(0.0)
        L.check(L.lib().b200als_half_iteration(self._h, which, int(solver_override), C.byref(loss)))
        return loss.value

    def fit(self, n_iter, convergence_tol):
        trace = np.zeros(2 * max(1, n_iter), np.float64)
        done = C.c_int(0)
        L.check(L.lib().b200als_fit(self._h, int(n_iter), float(convergence_tol), L.vp(trace), C.byref(done)))
        return trace[:2 * done.value].copy(), done.value

    def transform(self, n_local=None):
        out = np.empty((self.n_user if n_local is None else n_local, self.rank), np.float32)
        loss = C.c_double(0.0)
        L.check(L.lib().b200als_transform(self._h, L.vp(out), C.byref(loss)))
        return out, loss.value

    def exchange_mode(self):
        """'none' (single GPU / nothing exchanged yet), 'p2p' (peer-memory pushes) or 'nccl' (broadcast fallback)."""
        m = C.c_int(0)
        L.check(L.lib().b200als_exchange_mode(self._h, C.byref(m)))
        return ("none", "p2p", "nccl")[m.value]

    def set_bias(self, with_user_item_bias, global_bias=0.0):
        L.check(L.lib().b200als_set_bias(self._h, int(bool(with_user_item_bias)), float(global_bias)))

    def row_plan(self, which):
        """Rows per kernel of the last CG half-iteration of `which` and the local number of entries."""
        counts = np.zeros(10, np.int32)
        caps = np.zeros(9, np.int32)
        nnz = C.c_int64(0)
        L.check(L.lib().b200als_row_plan(self._h, which, L.vp(counts), L.vp(caps), C.byref(nnz)))
        # "long": rows beyond every tile / cluster class -- als_cg_gram_kernel at rank 128, the streaming kernel otherwise
        names = ["resident", "tile_4cta_double", "tile_4cta_single", "tile_2cta_single", "tile_1cta_double", "cluster2", "cluster4",
                 "cluster8", "long", "empty"]
        return {"rows": dict(zip(names, [int(v) for v in counts])), "longest_row": dict(zip(names[:9], [int(v) for v in caps])),
                "nnz_local": int(nnz.value)}

    def last_timing(self):
        t = [C.c_float(0) for _ in range(4)]
        L.check(L.lib().b200als_last_timing(self._h, *[C.byref(x) for x in t]))
        return dict(gram_ms=t[0].value, prep_ms=t[1].value, solve_ms=t[2].value, comm_ms=t[3].value)


class WRMF:
    """Weighted Regularized Matrix Factorization (mirror of R/model_WRMF.R:35-454).

    Parameters are the reference's (`lambda` is spelled `lambda_`).  With `with_user_item_bias` the factor matrices
    carry rank + 2 rows ([1, ..., user_bias] / [item_bias, ..., 1], R/model_WRMF.R:162-166, :205-246) and every
    half-iteration goes through the stateless C-ABI calls, as in R; without bias terms float models run on the
    resident-factor session."""

    def __init__(self, rank=10, lambda_=0.0, dynamic_lambda=True, init=None, preprocess=None, feedback="implicit",
                 solver="conjugate_gradient", with_user_item_bias=False, with_global_bias=False, cg_steps=3,
                 precision="double", seed=None, kernel=0):
        if init is not None and not isinstance(init, np.ndarray):
            raise TypeError("init must be a matrix")                      # stopifnot(is.null(init) || is.matrix(init))
        if solver not in _SOLVER_CODES:
            raise ValueError("solver should be one of %s" % list(_SOLVER_CODES))   # match.arg
        if feedback not in ("implicit", "explicit"):
            raise ValueError("feedback should be one of ['implicit', 'explicit']")
        if precision not in ("double", "float"):
            raise ValueError("precision should be one of ['double', 'float']")
        if not isinstance(cg_steps, (int, np.integer)):
            raise TypeError("cg_steps must be an integer")               # stopifnot(is.integer(cg_steps))
        self._solver_code = _SOLVER_CODES[solver]
        self._non_negative = solver == "nnls"
        if self._non_negative and with_global_bias:                      # R/model_WRMF.R:90-93
            logger.warning("setting `with_global_bias=FALSE` for 'nnls' solver")
            with_global_bias = False
        self._with_user_item_bias = bool(with_user_item_bias)
        self._with_global_bias = bool(with_global_bias)
        self.global_bias_base = None
        self._precision = precision
        self._feedback = feedback
        self._lambda = float(lambda_)
        self._dynamic_lambda = bool(dynamic_lambda)
        self._cg_steps = int(cg_steps)
        self._rank = int(rank) + (2 if self._with_user_item_bias else 0)   # R/model_WRMF.R:162-166
        self._preprocess = preprocess if preprocess is not None else (lambda m: m)
        self._rng = np.random.default_rng(seed)
        self._kernel = kernel
        self.components = init          # rank x n_item (R layout), set by fit_transform
        self.global_bias = 0.0
        self._U = None
        self._XtX = None
        self._cnt_u = None
        self._session = None
        self._dt = np.float64 if precision == "double" else np.float32

    # ---- private$solver (R/model_WRMF.R:111-147) over the stateless C ABI -------------------------
    def _solve(self, mat, X, Y, is_bias_last_row, cnt_X=None, XtX=None, avoid_cg=False):
        solver_use = 0 if (avoid_cg and self._solver_code == 1) else self._solver_code
        if self._feedback == "implicit":
            return als_implicit(*mat, X, Y, self._lambda, solver_use, self._cg_steps, XtX=XtX,
                                with_user_item_bias=self._with_user_item_bias, is_bias_last_row=is_bias_last_row,
                                global_bias=self.global_bias, global_bias_base=self.global_bias_base,
                                initialize_bias_base=not avoid_cg)
        return als_explicit(*mat, X, Y, cnt_X, self._lambda, solver_use, self._cg_steps, self._dynamic_lambda,
                            with_user_item_bias=self._with_user_item_bias, is_bias_last_row=is_bias_last_row)

    def fit_transform(self, x, n_iter=10, convergence_tol=None):
        if convergence_tol is None:
            convergence_tol = 0.005 if self._feedback == "implicit" else 0.001
        c_ui = self._preprocess(sp.csc_matrix(x))
        n_user, n_item = c_ui.shape
        if self._feedback != "explicit" or self._non_negative:
            if c_ui.nnz and c_ui.data.min() < 0:
                raise ValueError("all(c_ui@x >= 0) is not TRUE")          # R/model_WRMF.R:195-197
        items = _targets_csc(c_ui.T)      # c_ui: columns = items, idx = users
        users = _targets_csc(c_ui)        # c_iu: columns = users, idx = items
        dt = self._dt
        k = self._rank
        wuib = self._with_user_item_bias
        U = (self._rng.standard_normal((n_user, k)) / 100.0).astype(dt)   # large_rand_matrix / flrnorm(.., 0, 0.01)
        if wuib:
            U[:, 0] = 1.0                                                  # for item biases (R/model_WRMF.R:205-216)
        if self.components is None:
            if self._solver_code == 1:                                     # R/model_WRMF.R:219-230: zeros for CG
                comp = np.zeros((n_item, k), dt)
            else:
                comp = (self._rng.standard_normal((n_item, k)) / 100.0).astype(dt)
        else:
            if self.components.shape != (k, n_item):
                raise ValueError("init must be rank x n_item")
            comp = np.ascontiguousarray(self.components.T, dtype=dt)
        if wuib and self.components is None:
            comp[:, k - 1] = 1.0                                           # for user biases (:219-243)
        if self._non_negative:                                              # R/model_WRMF.R:251-255
            comp = np.abs(comp)
            U = np.abs(U)
        self.global_bias = 0.0
        if wuib:                                                            # R/model_WRMF.R:260-279
            user_bias = np.zeros(n_user, dt)
            item_bias = np.zeros(n_item, dt)
            global_bias = initialize_biases(items, users, user_bias, item_bias, self._lambda, self._dynamic_lambda,
                                            self._non_negative, self._with_global_bias, self._feedback == "explicit")
            comp[:, 0] = item_bias
            U[:, k - 1] = user_bias
            if self._with_global_bias:
                self.global_bias = global_bias
        elif self._with_global_bias:                                        # R/model_WRMF.R:280-289
            if self._feedback == "explicit":
                self.global_bias = float(np.mean(items[2])) if len(items[2]) else 0.0
                items[2][:] -= self.global_bias
                users[2][:] -= self.global_bias
            else:
                ssum = float(np.sum(items[2]))
                self.global_bias = ssum / (ssum + float(n_user) * float(n_item) - len(items[2]))
        if self._feedback == "implicit":
            # the reference sizes this rank-1 (R/model_WRMF.R:291-297) while als_implicit<T> reads and writes `rank`
            # entries (wrmf_implicit.hpp:111-112,155-157); the C ABI takes the length the C++ code uses
            self.global_bias_base = np.zeros(0 if wuib else k, dt)
        cnt_u = np.diff(items[0]).astype(dt)   # diff(c_ui@p): nnz per item -> cnt_X of the user half (:311)
        cnt_i = np.diff(users[0]).astype(dt)   # diff(c_iu@p): nnz per user -> cnt_X of the item half (:312)
        self._cnt_u = cnt_u
        logger.info("starting factorization")
        biased = wuib or (self._feedback == "implicit" and self.global_bias != 0.0)
        if self._precision == "float":
            # device-resident session; bias terms included (b200als_set_bias)
            res = self._fit_session(items, users, n_user, n_item, U, comp, n_iter, convergence_tol, biased)
        else:
            loss_prev = np.inf
            for i in range(int(n_iter)):
                loss = self._solve(items, U, comp, True, cnt_X=cnt_i)        # R/model_WRMF.R:319-323
                logger.info("iter %d (items) loss = %.4f", i + 1, loss)
                loss = self._solve(users, comp, U, False, cnt_X=cnt_u)       # R/model_WRMF.R:326-329
                logger.info("iter %d (users) loss = %.4f", i + 1, loss)
                if loss_prev / loss - 1 < convergence_tol:
                    logger.info("Converged after %d iterations", i + 1)
                    break
                loss_prev = loss
            self._U = U
            self._set_components(comp)
            res = self._transform(users)
        return res

    def _fit_session(self, items, users, n_user, n_item, U, comp, n_iter, convergence_tol, biased=False):
        s = Session(items, users, n_user, n_item, self._rank, self._feedback, self._solver_code, self._cg_steps,
                    self._dynamic_lambda, self._lambda, self._kernel)
        try:
            if biased:
                s.set_bias(self._with_user_item_bias, self.global_bias if self._feedback == "implicit" else 0.0)
            s.set_factors(L.USERS, U)
            s.set_factors(L.ITEMS, comp)
            trace, done = s.fit(n_iter, convergence_tol)
            for i in range(done):
                logger.info("iter %d (items) loss = %.4f", i + 1, trace[2 * i])
                logger.info("iter %d (users) loss = %.4f", i + 1, trace[2 * i + 1])
            self.loss_trace = trace
            comp = s.get_factors(L.ITEMS)
            self._U = s.get_factors(L.USERS)
            res, _ = s.transform()
            if biased and self._feedback == "implicit" and not self._with_user_item_bias and len(self.global_bias_base):
                # what the last user half-iteration left in self$global_bias_base (wrmf_implicit.hpp:111-112): the session kept
                # it on the device; transform() reuses it (initialize_bias_base = FALSE, R/model_WRMF.R:130)
                self.global_bias_base[:] = (-self.global_bias * comp.astype(np.float64).sum(axis=0)).astype(self._dt)
        finally:
            s.close()
        self._set_components(comp)
        return res

    def _set_components(self, comp):
        self._comp_rows = np.ascontiguousarray(comp)              # n_item x rank
        self.components = self._comp_rows.T                       # rank x n_item, as in R
        if self._feedback == "implicit":                          # private$XtX (R/model_WRMF.R:347-353)
            c64 = self._comp_rows.astype(self._dt)
            if self._with_user_item_bias:
                c64 = c64[:, 1:]                                  # components[-1L, ]: without the item-bias row
            k = c64.shape[1]
            self._XtX = (c64.T @ c64 + self._lambda * np.eye(k, dtype=self._dt)).astype(self._dt)

    # ---- transform_ (R/model_WRMF.R:412-452) -------------------------------------------------------
    def _transform(self, users):
        res = np.zeros((len(users[0]) - 1, self._rank), self._dt)
        if self._with_user_item_bias:
            res[:, 0] = 1.0                                        # R/model_WRMF.R:429-431
        self._solve(users, self._comp_rows.astype(self._dt, copy=False), res, False, cnt_X=self._cnt_u, XtX=self._XtX,
                    avoid_cg=True)
        return res

    def transform(self, x):
        if self.components is None:
            raise RuntimeError("model is not fitted")
        x = self._preprocess(sp.csr_matrix(x))
        if x.shape[1] != self.components.shape[1]:
            raise ValueError("ncol(x) == ncol(self$components) is not TRUE")   # R/model_WRMF.R:367
        users = _targets_csc(x)
        if self.global_bias != 0.0 and self._feedback == "explicit":       # R/model_WRMF.R:381-382
            users[2][:] -= self.global_bias
        return self._transform(users)

    def predict(self, x, k, not_recommend="x", items_exclude=(), return_scores=False):
        """MatrixFactorizationRecommender$predict (R/MatrixFactorizationRecommender.R:24-78): transform(x), then the
        top-k items per user, excluding `not_recommend` (default: the interactions in x) and `items_exclude`.
        Returns 0-based item indices (n_user, k), -1 where fewer than k candidates exist."""
        from .ops import top_product
        if items_exclude is not None and len(items_exclude) and not np.issubdtype(np.asarray(items_exclude).dtype, np.integer):
            raise TypeError("items_exclude should be one of character/integer")
        emb = self.transform(x)
        nr = x if isinstance(not_recommend, str) and not_recommend == "x" else not_recommend
        idx, scores = top_product(emb.astype(np.float32), self._comp_rows.astype(np.float32), k, nr,
                                  () if items_exclude is None else items_exclude, glob_mean=self.global_bias)
        return (idx, scores) if return_scores else idx
