"""Generates the committed golden fixtures under tests/golden/ -- run HERE (authoring container),
where /root/reference exists; the outputs travel with the repo, the reference does not.

Every "expected" array in the fixtures is produced by the REFERENCE'S OWN SOURCE
(/root/reference/inst/include/wrmf_implicit.hpp:90-305, wrmf_explicit.hpp:33-174) compiled
unmodified against oracle/mini_arma (oracle/build_ref.sh -> oracle/_ref/libref_wrmf.so), driven
the way R/model_WRMF.R drives it:
  * XtX = tcrossprod(X) + lambda*I in the working precision (R/model_WRMF.R:474-486)
  * Y is updated in place, the call returns loss/nnz
  * the ALS trace alternates item / user half-iterations (R/model_WRMF.R:318-338) with the
    CG zero-initialised `components` (R/model_WRMF.R:217-230) and ends with the avoid_cg
    transform (R/model_WRMF.R:412-452)
Inputs are stored in the fixture (float32 values, upcast for the f64 runs) so the tests do not
depend on any RNG implementation.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle.rdata import load_movielens100k  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CHOL, CG, NNLS = 0, 1, 2


def csc_cols_are_targets(M_targets_by_src):
    """scipy CSR (targets x src) -> (ptr, idx, val) = CSC whose columns are the targets."""
    M = sp.csr_matrix(M_targets_by_src)
    M.sort_indices()
    return M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)


def half_iteration_ref(ptr, idx, val, X32, Y32, feedback, solver, lam, cg_steps, dynamic_lambda, cnt_X, dt):
    X = X32.astype(dt)
    Y = Y32.astype(dt).copy()
    if feedback == "implicit":
        XtX = (X.T @ X + lam * np.eye(X.shape[1], dtype=dt)).astype(dt)
        loss = oracle.als_implicit(ptr, idx, val, X, Y, XtX, lam, solver, cg_steps, 1, impl="ref")
    else:
        loss = oracle.als_explicit(ptr, idx, val, X, Y, None if cnt_X is None else cnt_X.astype(dt), lam,
                                   solver, cg_steps, dynamic_lambda, 1, impl="ref")
    return Y, loss


def main():
    oracle.build()
    assert oracle.ref_available(), "oracle/_ref/libref_wrmf.so missing (needs /root/reference)"
    i, p, x, dim = load_movielens100k()
    np.savez_compressed(os.path.join(OUT, "movielens100k.npz"), i=i, p=p, x=x.astype(np.uint8), dim=dim)
    import wrmf_cases as wc
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    n_user, n_item = M.shape

    # ---- single half-iterations: inputs are rebuilt by tests/wrmf_cases.py, only outputs are stored ----
    out = {}
    for name, c in wc.half_iteration_cases().items():
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            Y, loss = half_iteration_ref(c["ptr"], c["idx"], c["val"], c["X"], c["Y0"], c["feedback"], c["solver"],
                                         c["lam"], c["cg_steps"], c["dynamic_lambda"], c["cnt_X"], dt)
            out[name + "/Y_" + tag] = Y
            out[name + "/loss_" + tag] = np.float64(loss)
        print("%-40s loss f64 %.8f f32 %.8f" % (name, out[name + "/loss_f64"], out[name + "/loss_f32"]))
    np.savez_compressed(os.path.join(OUT, "half_iterations.npz"), **out)

    # ---- bias variants (with_user_item_bias / global_bias), also straight from the reference's headers -----------
    out = {}
    for name, c in wc.bias_cases().items():
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            X = c["X"].astype(dt)
            Y = c["Y0"].astype(dt).copy()
            if c["feedback"] == "implicit":
                gbb = np.zeros(X.shape[1] - int(c["with_biases"]), dt)
                loss = oracle.als_implicit_bias(c["ptr"], c["idx"], c["val"], X, Y, wc.xtx_for(c, dt), c["lam"], c["solver"],
                                                c["cg_steps"], c["with_biases"], c["is_last"], c["gbias"], gbb, True, 1,
                                                impl="ref")
                out[name + "/gbb_" + tag] = gbb
            else:
                loss = oracle.als_explicit_bias(c["ptr"], c["idx"], c["val"], X, Y, c["cnt_X"].astype(dt), c["lam"],
                                                c["solver"], c["cg_steps"], c["dynamic_lambda"], c["with_biases"],
                                                c["is_last"], 1, impl="ref")
            out[name + "/Y_" + tag] = Y
            out[name + "/loss_" + tag] = np.float64(loss)
        print("%-40s loss f64 %.8f f32 %.8f" % (name, out[name + "/loss_f64"], out[name + "/loss_f32"]))
    # initialize_biases<T> (wrmf_utils.hpp:170-183) on movielens100k
    for is_explicit in (True, False):
        for nn in (False, True):
            for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
                csc = (items[0], items[1], items[2].copy())
                csr = (users[0], users[1], users[2].copy())
                ub, ib = np.zeros(n_user, dt), np.zeros(n_item, dt)
                g = oracle.initialize_biases(csc, csr, ub, ib, 0.1, True, nn, True, is_explicit, impl="ref")
                key = "init_%s_nn%d_%s" % ("explicit" if is_explicit else "implicit", int(nn), tag)
                out[key + "/user_bias"], out[key + "/item_bias"], out[key + "/global_bias"] = ub, ib, np.float64(g)
                print("%-40s global_bias %.8f" % (key, g))
    np.savez_compressed(os.path.join(OUT, "bias_half_iterations.npz"), **out)

    def factors(n, k, seed=[100]):
        seed[0] += 1
        return wc.det_factors(n, k, seed[0])

    # ---- ALS trace on movielens100k (fit_transform flow, R/model_WRMF.R:173-360) ---------------------
    trace = {}
    for feedback, solver, k, lam in (("implicit", CG, 16, 0.1), ("implicit", CHOL, 8, 0.1), ("explicit", CG, 8, 0.1)):
        name = "ml100k_%s_%s_k%d" % (feedback, "cg" if solver == CG else "chol", k)
        U0 = factors(n_user, k)
        I0 = np.zeros((n_item, k), np.float32) if solver == CG else factors(n_item, k)
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            U, I = U0.astype(dt), I0.astype(dt)
            cnt_u = np.diff(items[0]).astype(dt)   # diff(c_ui@p): nnz per item  (R/model_WRMF.R:311)
            cnt_i = np.diff(users[0]).astype(dt)   # diff(c_iu@p): nnz per user  (R/model_WRMF.R:312)
            losses = []
            for it in range(3):
                for (mat, Xf, Yf, cnt) in ((items, U, I, cnt_i), (users, I, U, cnt_u)):
                    if feedback == "implicit":
                        G = (Xf.T @ Xf + lam * np.eye(k, dtype=dt)).astype(dt)
                        losses.append(oracle.als_implicit(*mat, Xf, Yf, G, lam, solver, 3, 1, impl="ref"))
                    else:
                        losses.append(oracle.als_explicit(*mat, Xf, Yf, cnt, lam, solver, 3, True, 1, impl="ref"))
            # final transform_: Y = 0, stored XtX, CG -> Cholesky (R/model_WRMF.R:347-359, :412-452)
            res = np.zeros_like(U)
            if feedback == "implicit":
                G = (I.T @ I + lam * np.eye(k, dtype=dt)).astype(dt)
                oracle.als_implicit(*users, I, res, G, lam, CHOL, 3, 1, impl="ref")
            else:
                oracle.als_explicit(*users, I, res, cnt_u, lam, CHOL, 3, True, 1, impl="ref")
            trace[name + "/losses_" + tag] = np.array(losses)
            trace[name + "/components_" + tag] = I
            trace[name + "/user_emb_" + tag] = res
            print(name, tag, "losses", np.round(losses, 5))
        trace[name + "/U0"] = U0
        trace[name + "/I0"] = I0
        trace[name + "/solver"] = np.int32(solver)
        trace[name + "/feedback"] = np.array(feedback)
        trace[name + "/lam"] = np.float64(lam)
    np.savez_compressed(os.path.join(OUT, "als_traces.npz"), **trace)
    for f in ("movielens100k.npz", "half_iterations.npz", "als_traces.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
