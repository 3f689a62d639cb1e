"""GPU parity for the bias variants and initialize_biases, through the C ABI, against the golden
vectors generated from the reference's own headers (tests/golden/bias_half_iterations.npz) and the CPU oracle.
Tolerances as in test_gpu_parity.py: fp64 kernels 1e-9, fp32 engine vs the fp64 reference 1e-5 -- except
solver = nnls, an iteration stopped at a relative coordinate step of 1e-4 (inst/include/nnls.hpp:44): two correct
implementations agree to about that step, fp32 ones less when the system contains the column of ones."""
import numpy as np
import pytest

import oracle
import wrmf_cases as wc
from rsparse_b200 import WRMF, Session, als_explicit, als_implicit
from rsparse_b200 import _lib as L
from rsparse_b200.ops import initialize_biases
from test_oracle_bias import run_oracle_bias

pytestmark = pytest.mark.gpu

BIAS_CASES = sorted(wc.bias_cases().keys())


def relF(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_engine_bias(c, dt, XtX="host"):
    X = c["X"].astype(dt)
    Y = c["Y0"].astype(dt).copy()
    gbb = None
    if c["feedback"] == "implicit":
        gbb = np.zeros(X.shape[1] - int(c["with_biases"]), dt)
        G = wc.xtx_for(c, dt) if XtX == "host" else None
        loss = als_implicit(c["ptr"], c["idx"], c["val"], X, Y, c["lam"], c["solver"], c["cg_steps"], XtX=G,
                            with_user_item_bias=c["with_biases"], is_bias_last_row=c["is_last"], global_bias=c["gbias"],
                            global_bias_base=gbb, initialize_bias_base=True)
    else:
        loss = als_explicit(c["ptr"], c["idx"], c["val"], X, Y, c["cnt_X"].astype(dt), c["lam"], c["solver"], c["cg_steps"],
                            c["dynamic_lambda"], with_user_item_bias=c["with_biases"], is_bias_last_row=c["is_last"])
    return Y, loss, gbb


def tols(c, dt, ref32_err=0.0):
    """(factor tolerance, loss tolerance).  fp32: 1e-5, or three times the distance of the REFERENCE's own fp32 result
    from its fp64 result where that is larger (the systems with the column of ones are badly conditioned in fp32:
    the reference itself is off by up to 9e-5 there)."""
    if c["solver"] == wc.NNLS:
        return (1e-6, 1e-8) if dt == np.float64 else (max(5e-4, 3 * ref32_err), 1e-4)
    return (1e-9, 1e-9) if dt == np.float64 else (max(1e-5, 3 * ref32_err), 1e-5)


@pytest.mark.parametrize("name", BIAS_CASES)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("xtx", ["host", "engine"])
def test_bias_half_iteration_vs_reference_golden(name, dt, xtx, bias_cases, golden_bias):
    c = bias_cases[name]
    if xtx == "engine" and c["feedback"] != "implicit":
        pytest.skip("XtX only exists for implicit feedback")
    Y, loss, gbb = run_engine_bias(c, dt, xtx)
    ref = golden_bias[name + "/Y_f64"]
    tol_y, tol_l = tols(c, dt, relF(golden_bias[name + "/Y_f32"], ref))
    assert relF(Y, ref) < tol_y, (relF(Y, ref), relF(golden_bias[name + "/Y_f32"], ref))
    assert abs(loss - float(golden_bias[name + "/loss_f64"])) <= tol_l * abs(loss)
    if c["with_biases"]:   # the row of ones of Y is not written
        assert np.all(Y[:, -1 if c["is_last"] else 0] == 1)
    if gbb is not None and c["gbias"] and not c["with_biases"]:   # global_bias_base is an output (wrmf_implicit.hpp:111-112)
        assert relF(gbb, golden_bias[name + "/gbb_f64"]) < (1e-12 if dt == np.float64 else 1e-6)


def test_global_bias_base_is_reused_when_not_initialised(bias_cases, golden_bias):
    """initialize_bias_base = FALSE (transform, R/model_WRMF.R:438-446): the caller's base is read, not rewritten."""
    c = bias_cases["global_rag_implicit_chol"]
    X = c["X"].astype(np.float64)
    Y = c["Y0"].astype(np.float64).copy()
    gbb = golden_bias["global_rag_implicit_chol/gbb_f64"].copy()
    keep = gbb.copy()
    loss = als_implicit(c["ptr"], c["idx"], c["val"], X, Y, c["lam"], c["solver"], XtX=wc.xtx_for(c, np.float64),
                        global_bias=c["gbias"], global_bias_base=gbb, initialize_bias_base=False)
    assert np.array_equal(gbb, keep)
    assert relF(Y, golden_bias["global_rag_implicit_chol/Y_f64"]) < 1e-9
    assert abs(loss - float(golden_bias["global_rag_implicit_chol/loss_f64"])) < 1e-9 * loss
    # a global bias below sqrt(eps) counts as none (wrmf_implicit.hpp:108-109)
    Y1, Y2 = c["Y0"].astype(np.float64).copy(), c["Y0"].astype(np.float64).copy()
    l1 = als_implicit(c["ptr"], c["idx"], c["val"], X, Y1, c["lam"], c["solver"], global_bias=1e-9)
    l2 = als_implicit(c["ptr"], c["idx"], c["val"], X, Y2, c["lam"], c["solver"])
    assert np.array_equal(Y1, Y2) and abs(l1 - l2) <= 1e-13 * l2   # rows are summed in ticket order


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("is_last", [False, True])
def test_implicit_cg_with_biases_is_solved_not_rejected(dt, is_last, bias_cases):
    """The reference's implicit + conjugate_gradient + with_biases branch fails on a dimension mismatch (`init` loses
    a row at wrmf_implicit.hpp:191 and again at :199).  The engine (and the oracle) drop the row once.  Checks: equal
    to the oracle, and with enough steps CG lands on the Cholesky solution of the same system."""
    c = dict(bias_cases["bias_ml100k_item_implicit_chol" if is_last else "bias_ml100k_user_implicit_chol"])
    c["solver"], c["cg_steps"] = wc.CG, 3
    Y, loss, _ = run_engine_bias(c, dt)
    Yo, lo, _ = run_oracle_bias(c, np.float64, n_threads=oracle.max_threads())
    if dt == np.float64:
        tol_y = tol_l = 1e-9
    else:   # yardstick: how far the fp32 oracle lands from the fp64 oracle on this (badly conditioned) system
        Y32, l32, _ = run_oracle_bias(c, np.float32, n_threads=oracle.max_threads())
        tol_y, tol_l = max(1e-5, 3 * relF(Y32, Yo)), max(1e-5, 3 * abs(l32 - lo) / abs(lo))
    assert relF(Y, Yo) < tol_y
    assert abs(loss - lo) <= tol_l * abs(lo)
    if dt == np.float64:
        c["cg_steps"] = 60
        Ycg, _, _ = run_engine_bias(c, dt)
        c["solver"] = wc.CHOL
        Ych, _, _ = run_engine_bias(c, dt)
        assert relF(Ycg, Ych) < 1e-6


@pytest.mark.parametrize("is_explicit", [True, False])
@pytest.mark.parametrize("non_negative", [False, True])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_initialize_biases_vs_reference_golden(is_explicit, non_negative, dt, golden_bias):
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    csc = (items[0], items[1], items[2].copy())
    csr = (users[0], users[1], users[2].copy())
    ub, ib = np.zeros(M.shape[0], dt), np.zeros(M.shape[1], dt)
    g = initialize_biases(csc, csr, ub, ib, 0.1, True, non_negative, True, is_explicit)
    key = "init_%s_nn%d_%s" % ("explicit" if is_explicit else "implicit", int(non_negative), "f64" if dt == np.float64 else "f32")
    # per-column arithmetic is the reference's; the scalar means are tree sums instead of running means
    tol = 1e-12 if dt == np.float64 else 2e-6
    assert abs(g - float(golden_bias[key + "/global_bias"])) <= 1e-13 * abs(g)
    assert relF(ub, golden_bias[key + "/user_bias"]) < tol
    assert relF(ib, golden_bias[key + "/item_bias"]) < tol
    if non_negative:
        assert ub.min() >= 0 and ib.min() >= 0
    if is_explicit:
        assert np.allclose(csc[2], items[2] - g, rtol=0, atol=1e-12) and np.allclose(csr[2], users[2] - g, rtol=0, atol=1e-12)
    else:
        assert np.array_equal(csc[2], items[2])


def test_initialize_biases_without_global_bias_and_static_lambda():
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    for is_explicit in (True, False):
        res = []
        for fn, kw in ((initialize_biases, {}), (oracle.initialize_biases, {})):
            csc = (items[0], items[1], items[2].copy())
            csr = (users[0], users[1], users[2].copy())
            ub, ib = np.zeros(M.shape[0]), np.zeros(M.shape[1])
            g = fn(csc, csr, ub, ib, 2.5, False, False, False, is_explicit, **kw)
            res.append((ub, ib, g, csc[2]))
        assert res[0][2] == 0 and res[1][2] == 0
        assert relF(res[0][0], res[1][0]) < 1e-12 and relF(res[0][1], res[1][1]) < 1e-12
        assert np.array_equal(res[0][3], items[2])


@pytest.mark.parametrize("precision", ["float", "double"])
@pytest.mark.parametrize("feedback,solver,lam", [
    ("implicit", "cholesky", 0.1), ("implicit", "nnls", 0.1), ("implicit", "cholesky", 1000.0), ("implicit", "cholesky", 0.0),
    ("explicit", "conjugate_gradient", 0.1), ("explicit", "cholesky", 0.1), ("explicit", "nnls", 0.1),
    ("explicit", "cholesky", 1000.0), ("explicit", "conjugate_gradient", 1000.0)])
def test_wrmf_class_with_user_item_bias_like_reference_tests(precision, feedback, solver, lam):
    """tests/testthat/test-wrmf.R:9-71 with with_user_item_bias = TRUE: rank + 2 columns, fit_transform(train) ==
    transform(train), predict / transform on held-out users, non-negativity for nnls."""
    M = wc.load_movielens()
    train, cv = M[:900], M[900:]
    model = WRMF(rank=6, lambda_=lam, feedback=feedback, solver=solver, with_user_item_bias=True, precision=precision, seed=1)
    emb = model.fit_transform(train, n_iter=5, convergence_tol=-1)
    assert emb.shape == (900, 8)
    assert model.components.shape == (8, M.shape[1])
    assert np.all(emb[:, 0] == 1) and np.all(model.components[-1, :] == 1)
    emb2 = model.transform(train)
    tol = 2e-2 if solver == "nnls" else (1e-4 if precision == "float" else 1e-9)
    assert relF(emb2, emb) < tol
    emb_cv = model.transform(cv)
    assert emb_cv.shape == (cv.shape[0], 8) and np.all(np.isfinite(emb_cv))
    preds = model.predict(cv, k=7)
    assert preds.shape == (cv.shape[0], 7)
    if solver == "nnls":
        assert np.all(emb >= 0) and np.all(emb_cv >= 0) and np.all(model.components >= 0)


@pytest.mark.parametrize("precision", ["float", "double"])
@pytest.mark.parametrize("feedback,solver,wuib", [("implicit", "cholesky", False), ("implicit", "conjugate_gradient", False),
                                                  ("implicit", "cholesky", True), ("explicit", "cholesky", False),
                                                  ("explicit", "conjugate_gradient", True)])
def test_wrmf_class_with_global_bias(precision, feedback, solver, wuib):
    """with_global_bias (R/model_WRMF.R:280-289, :381-382): the model learns on centred ratings (explicit) or with the
    global-bias right-hand side (implicit); more iterations do not increase the loss."""
    M = wc.load_movielens()
    train = M[:900]
    model = WRMF(rank=6, lambda_=0.1, feedback=feedback, solver=solver, with_user_item_bias=wuib, with_global_bias=True,
                 precision=precision, seed=2)
    emb = model.fit_transform(train, n_iter=3, convergence_tol=-1)
    assert np.all(np.isfinite(emb)) and emb.shape == (900, 6 + 2 * int(wuib))
    if feedback == "explicit":
        assert abs(model.global_bias - train.data.mean()) < 1e-9
    else:
        s = train.data.sum()
        assert abs(model.global_bias - s / (s + 900.0 * M.shape[1] - train.nnz)) < 1e-12
    assert relF(model.transform(train), emb) < (1e-4 if precision == "float" else 1e-9)


@pytest.mark.parametrize("feedback,solver,gbias", [("implicit", wc.CHOL, 0.0), ("implicit", wc.CHOL, 0.05), ("explicit", wc.CG, 0.0),
                                                   ("explicit", wc.CHOL, 0.0), ("implicit", wc.NNLS, 0.0)])
def test_biased_session_matches_the_chain_of_reference_shaped_calls(feedback, solver, gbias):
    """Bias terms inside the device-resident session (b200als_set_bias): two ALS iterations (item half with
    is_bias_last_row = TRUE, user half with FALSE, R/model_WRMF.R:318-338) and the final transform_ give what the same chain
    of reference-shaped stateless calls gives -- those are checked against the reference's golden vectors above -- and what
    the fp64 oracle gives, without any host round trip between half-iterations."""
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    n_user, n_item = M.shape
    rank, lam = 6, 0.1
    kf = rank + 2
    U = wc.det_factors(n_user, kf, 81, 0.1)
    I = wc.det_factors(n_item, kf, 82, 0.1)
    if solver == wc.NNLS:
        U, I = np.abs(U), np.abs(I)
    U[:, 0] = 1.0            # users: [1, ..., user_bias]
    I[:, kf - 1] = 1.0       # items: [item_bias, ..., 1]
    cnt_i = np.diff(users[0]).astype(np.float32)      # nnz per user: cnt_X of the item half
    cnt_u = np.diff(items[0]).astype(np.float32)      # nnz per item: cnt_X of the user half
    s = Session(items, users, n_user, n_item, kf, feedback, solver, 3, True, lam)
    s.set_bias(True, gbias)
    s.set_factors(L.USERS, U)
    s.set_factors(L.ITEMS, I)
    trace, done = s.fit(2, -1.0)
    Us, Is = s.get_factors(L.USERS), s.get_factors(L.ITEMS)
    emb, _ = s.transform()
    s.close()
    assert done == 2 and np.all(Us[:, 0] == 1) and np.all(Is[:, kf - 1] == 1) and np.all(emb[:, 0] == 1)

    def chain(dt, fi, fe):
        Uc, Ic = U.astype(dt), I.astype(dt)
        losses = []
        for _ in range(2):
            for mat, X, Y, last, cnt in ((items, Uc, Ic, True, cnt_i), (users, Ic, Uc, False, cnt_u)):
                if feedback == "implicit":
                    gbb = np.zeros(kf - 1, dt)
                    losses.append(fi(mat, X, Y, last, gbb))
                else:
                    losses.append(fe(mat, X, Y, last, cnt.astype(dt)))
        return Uc, Ic, losses

    # (1) the same chain through the stateless, reference-shaped calls (fp32, GPU)
    U1, I1, l1 = chain(np.float32,
                       lambda m, X, Y, last, gbb: als_implicit(m[0], m[1], m[2], X, Y, lam, solver, 3, with_user_item_bias=True,
                                                               is_bias_last_row=last, global_bias=gbias, global_bias_base=gbb,
                                                               initialize_bias_base=True),
                       lambda m, X, Y, last, cnt: als_explicit(m[0], m[1], m[2], X, Y, cnt, lam, solver, 3, True,
                                                               with_user_item_bias=True, is_bias_last_row=last))
    tol = 5e-3 if solver == wc.NNLS else 2e-5
    assert relF(Us, U1) < tol and relF(Is, I1) < tol
    assert np.allclose(trace, l1, rtol=1e-3 if solver == wc.NNLS else 1e-5)
    # (2) the fp64 oracle (CPU restatement of every bias branch, pinned to the reference in test_oracle_bias.py)
    U2, I2, l2 = chain(np.float64,
                       lambda m, X, Y, last, gbb: oracle.als_implicit_bias(m[0], m[1], m[2], X, Y, wc.xtx_for(dict(X=X, lam=lam, with_biases=True, is_last=last), np.float64),
                                                                           lam, solver, 3, True, last, gbias, gbb, True),
                       lambda m, X, Y, last, cnt: oracle.als_explicit_bias(m[0], m[1], m[2], X, Y, cnt, lam, solver, 3, True, True, last))
    tol2 = 2e-2 if solver == wc.NNLS else 2e-4      # two chained iterations, fp32 vs fp64 (cf. the 3-iteration trace bound)
    assert relF(Us, U2) < tol2 and relF(Is, I2) < tol2
    assert np.allclose(trace, l2, rtol=1e-3 if solver == wc.NNLS else 5e-5)
