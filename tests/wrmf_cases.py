"""Deterministic, RNG-free input builders shared by tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py (pure integer hashing -> identical on every platform)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHOL, CG, NNLS = 0, 1, 2


def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def det_uniform(n, seed):
    """float32 uniforms in [0,1) with 24 exact bits."""
    with np.errstate(over="ignore"):
        h = splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3))
    return ((h >> np.uint64(40)).astype(np.float32)) / np.float32(1 << 24)


def det_factors(n, k, seed, scale=0.01):
    """(n, k) float32 'N(0, scale)'-like factors: sum of 4 uniforms, centred, exact in float32."""
    u = det_uniform(4 * n * k, seed).reshape(4, n * k)
    z = (u[0] + u[1] + u[2] + u[3] - np.float32(2.0)) * np.float32(scale * 1.7320508)  # var(sum)=1/3
    return np.ascontiguousarray(z.reshape(n, k).astype(np.float32))


def det_csr(n_rows, n_cols, nnz_per_row, seed, ragged=False, empty_every=0, explicit=False):
    """CSC-with-columns-as-targets triplet (ptr int32, idx int32 ascending+distinct, val float64)."""
    rows = []
    with np.errstate(over="ignore"):
        for r in range(n_rows):
            n = nnz_per_row
            if ragged:
                n = 1 + int(splitmix64(np.uint64(seed * 7919 + r))) % (2 * nnz_per_row - 1)
            if empty_every and r % empty_every == empty_every - 1:
                n = 0
            n = min(n, n_cols)
            h = splitmix64(np.arange(n_cols, dtype=np.uint64) + np.uint64((seed * 1000003 + r) * 65537))
            rows.append(np.sort(np.argsort(h, kind="stable")[:n]).astype(np.int32))
    ptr = np.zeros(n_rows + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(c) for c in rows])
    idx = np.concatenate(rows).astype(np.int32) if ptr[-1] else np.zeros(0, np.int32)
    u = det_uniform(int(ptr[-1]), seed + 12345)
    if explicit:
        val = (1 + np.floor(u * 5)).astype(np.float64)
    else:
        val = (1.0 + np.floor(10.0 * u.astype(np.float64) ** 2)).astype(np.float64)
    return ptr, idx, val


def load_movielens():
    """users x items dgCMatrix slots from the committed fixture (reference data/movielens100k.RData)."""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLDEN, "movielens100k.npz"))
    M = sp.csc_matrix((z["x"].astype(np.float64), z["i"], z["p"]), shape=tuple(z["dim"])).tocsr()
    M.sort_indices()
    return M


def targets_csc(M_targets_by_src):
    import scipy.sparse as sp
    M = sp.csr_matrix(M_targets_by_src)
    M.sort_indices()
    return M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)


def half_iteration_cases():
    """name -> dict(ptr, idx, val, X, Y0, feedback, solver, lam, cg_steps, dynamic_lambda, cnt_X)."""
    M = load_movielens()
    users, items = targets_csc(M), targets_csc(M.T)
    n_user, n_item = M.shape
    cnt_items = np.diff(items[0]).astype(np.float32)
    cnt_users = np.diff(users[0]).astype(np.float32)
    C = {}

    def add(name, mat, n_src, k, feedback, solver, lam, seed, cg_steps=3, dynamic_lambda=True, cnt_X=None, scale=0.01):
        ptr, idx, val = mat
        C[name] = dict(ptr=ptr, idx=idx, val=val, X=det_factors(n_src, k, seed, scale),
                       Y0=det_factors(len(ptr) - 1, k, seed + 1, scale), feedback=feedback, solver=solver, lam=lam,
                       cg_steps=cg_steps, dynamic_lambda=dynamic_lambda, cnt_X=cnt_X)

    add("ml100k_user_implicit_cg_k16", users, n_item, 16, "implicit", CG, 0.1, 1)      # BASELINE configs[0]
    add("ml100k_item_implicit_cg_k16", items, n_user, 16, "implicit", CG, 0.1, 2)
    add("ml100k_user_implicit_chol_k16", users, n_item, 16, "implicit", CHOL, 0.1, 3)
    add("ml100k_user_implicit_chol_k10_lam0", users, n_item, 10, "implicit", CHOL, 0.0, 4)
    add("ml100k_user_implicit_cg_k7_lam1000", users, n_item, 7, "implicit", CG, 1000.0, 5)
    add("ml100k_user_explicit_cg_k8", users, n_item, 8, "explicit", CG, 0.1, 6, cnt_X=cnt_items)
    add("ml100k_user_explicit_chol_k8", users, n_item, 8, "explicit", CHOL, 0.1, 7, cnt_X=cnt_items)
    add("ml100k_item_explicit_chol_k9_static", items, n_user, 9, "explicit", CHOL, 1000.0, 8, dynamic_lambda=False,
        cnt_X=cnt_users)
    add("synth_implicit_chol_k64", det_csr(600, 400, 50, 11), 400, 64, "implicit", CHOL, 0.1, 11)   # C2-shaped
    add("synth_implicit_cg_k128", det_csr(500, 600, 80, 12), 600, 128, "implicit", CG, 0.1, 12)      # C3-shaped
    m = det_csr(500, 600, 80, 13, explicit=True)
    add("synth_explicit_cg_k128", m, 600, 128, "explicit", CG, 0.1, 13,
        cnt_X=np.bincount(m[1], minlength=600).astype(np.float32))                                  # C4-shaped
    add("synth_implicit_cg_k256", det_csr(120, 500, 100, 14), 500, 256, "implicit", CG, 0.1, 14)     # C5-shaped
    rag = det_csr(400, 900, 40, 15, ragged=True, empty_every=9)
    add("synth_ragged_implicit_cg_k128", rag, 900, 128, "implicit", CG, 0.1, 15)
    add("synth_ragged_implicit_chol_k32", rag, 900, 32, "implicit", CHOL, 0.5, 16)
    add("synth_ragged_explicit_cg_k64", det_csr(400, 900, 40, 17, ragged=True, empty_every=7, explicit=True), 900, 64,
        "explicit", CG, 0.05, 17, cnt_X=np.ones(900, np.float32))
    add("synth_long_implicit_cg_k128", det_csr(24, 1000, 350, 18, ragged=True), 1000, 128, "implicit", CG, 0.1, 18)
    add("synth_cg_early_exit_k16", det_csr(200, 300, 10, 19), 300, 16, "implicit", CG, 10.0, 19, scale=1e-5)
    add("synth_implicit_cg5_k32", det_csr(300, 300, 20, 20), 300, 32, "implicit", CG, 0.1, 20, cg_steps=5)
    # solver = "nnls" (inst/include/nnls.hpp): non-negative warm start as R does (abs(), R/model_WRMF.R:251-255)
    add("ml100k_user_implicit_nnls_k10", users, n_item, 10, "implicit", NNLS, 0.1, 21)
    add("ml100k_item_explicit_nnls_k8", items, n_user, 8, "explicit", NNLS, 0.1, 22, cnt_X=cnt_users)
    add("synth_ragged_implicit_nnls_k32", rag, 900, 32, "implicit", NNLS, 0.5, 23)
    for name in ("ml100k_user_implicit_nnls_k10", "ml100k_item_explicit_nnls_k8", "synth_ragged_implicit_nnls_k32"):
        C[name]["X"] = np.abs(C[name]["X"])
        C[name]["Y0"] = np.abs(C[name]["Y0"])
    return C


def xtx_for(c, dt):
    """XtX as R builds it for als_implicit (R/model_WRMF.R:474-486): tcrossprod of X without its bias row + lambda I."""
    X = c["X"].astype(dt)
    if c.get("with_biases"):
        X = X[:, :-1] if c["is_last"] else X[:, 1:]
    return (X.T @ X + c["lam"] * np.eye(X.shape[1], dtype=dt)).astype(dt)


def bias_cases():
    """Half-iterations with with_user_item_bias / global_bias.  Same dict as half_iteration_cases() plus
    with_biases, is_last (is_x_bias_last_row), gbias.  X / Y0 follow the reference's layouts (wrmf_implicit.hpp:96-101):
    is_last: X = [1, ..., x_bias], Y = [y_bias, ..., 1]; otherwise X = [x_bias, ..., 1], Y = [1, ..., y_bias]."""
    M = load_movielens()
    users, items = targets_csc(M), targets_csc(M.T)
    n_user, n_item = M.shape
    cnt_items = np.diff(items[0]).astype(np.float32)
    cnt_users = np.diff(users[0]).astype(np.float32)
    rag = det_csr(400, 900, 40, 31, ragged=True, empty_every=9)
    rag_e = det_csr(400, 900, 40, 32, ragged=True, empty_every=7, explicit=True)
    C = {}

    def add(name, mat, n_src, rank, feedback, solver, lam, seed, wb, is_last, gbias=0.0, dynamic_lambda=True, cnt_X=None,
            cg_steps=3):
        ptr, idx, val = mat
        kf = rank + (2 if wb else 0)
        X = det_factors(n_src, kf, seed, 0.1)
        Y0 = det_factors(len(ptr) - 1, kf, seed + 1, 0.1)
        if solver == NNLS:
            X, Y0 = np.abs(X), np.abs(Y0)
        if wb:
            if is_last:
                X[:, 0] = 1.0
                Y0[:, -1] = 1.0
            else:
                X[:, -1] = 1.0
                Y0[:, 0] = 1.0
        C[name] = dict(ptr=ptr, idx=idx, val=val, X=X, Y0=Y0, feedback=feedback, solver=solver, lam=lam, cg_steps=cg_steps,
                       dynamic_lambda=dynamic_lambda, cnt_X=cnt_X, with_biases=wb, is_last=is_last, gbias=gbias)

    add("bias_ml100k_user_implicit_chol", users, n_item, 10, "implicit", CHOL, 0.1, 41, True, False)
    add("bias_ml100k_item_implicit_chol", items, n_user, 6, "implicit", CHOL, 0.1, 42, True, True)
    add("bias_rag_implicit_chol_global", rag, 900, 8, "implicit", CHOL, 0.1, 43, True, False, gbias=0.05)
    add("bias_rag_implicit_nnls", rag, 900, 8, "implicit", NNLS, 0.1, 44, True, True)
    add("global_rag_implicit_chol", rag, 900, 16, "implicit", CHOL, 0.1, 45, False, False, gbias=0.05)
    add("global_ml100k_user_implicit_cg", users, n_item, 16, "implicit", CG, 0.1, 46, False, False, gbias=0.19)
    add("global_rag_implicit_nnls", rag, 900, 12, "implicit", NNLS, 0.1, 47, False, True, gbias=0.05)
    add("bias_ml100k_user_explicit_cg", users, n_item, 8, "explicit", CG, 0.1, 48, True, False, cnt_X=cnt_items)
    add("bias_ml100k_item_explicit_cg_static", items, n_user, 8, "explicit", CG, 5.0, 49, True, True, dynamic_lambda=False,
        cnt_X=cnt_users)
    add("bias_rag_explicit_chol", rag_e, 900, 10, "explicit", CHOL, 0.1, 50, True, False, cnt_X=np.ones(900, np.float32))
    add("bias_rag_explicit_nnls", rag_e, 900, 10, "explicit", NNLS, 0.1, 51, True, True, cnt_X=np.ones(900, np.float32))
    return C


def topk_cases():
    """name -> dict(x, y, k, nr (scipy CSR or None), exclude (0-based), glob_mean): inputs of `top_product`
    (src/matrix_top_product.cpp:20-102).  Includes exact ties (duplicated item rows, all-zero users), NA padding
    (k > number of admissible items), per-user and global exclusions."""
    import scipy.sparse as sp
    C = {}

    def add(name, n_user, n_item, rank, k, seed, density=0.1, exclude=(), glob_mean=0.0, with_nr=True, dup=False, zero_users=0):
        x = det_factors(n_user, rank, seed, 1.0)
        y = det_factors(n_item, rank, seed + 1, 1.0)
        if dup:                                     # exact score ties: every third item repeats its predecessor
            y[2::3] = y[1::3][: len(y[2::3])]
        if zero_users:
            x[:zero_users] = 0.0
        nr = None
        if with_nr:
            u = det_uniform(n_user * n_item, seed + 2).reshape(n_user, n_item)
            nr = sp.csr_matrix((u < density).astype(np.float64))
            nr.sort_indices()
        C[name] = dict(x=x, y=y, k=k, nr=nr, exclude=list(exclude), glob_mean=glob_mean)

    add("plain_100x50_r10_k10", 100, 50, 10, 10, 150, with_nr=False)
    add("filter_70x333_r16_k7", 70, 333, 16, 7, 433)
    add("filter_excl_33x1000_r128_k40", 33, 1000, 128, 40, 1100, exclude=(1, 3, 999), glob_mean=0.25)
    add("na_padding_5x20_r4_k20", 5, 20, 4, 20, 120, density=0.3, exclude=(0, 7))
    add("ties_40x90_r8_k12", 40, 90, 8, 12, 77, dup=True, zero_users=3, exclude=(0,))
    add("ties_nofilter_9x30_r4_k30", 9, 30, 4, 30, 78, with_nr=False, dup=True, zero_users=2)
    return C
