"""GPU parity tests: the CUDA path, called through the C ABI, against
  (1) the committed golden vectors generated from the reference's own source,
  (2) the CPU oracle on seeded inputs at sizes the oracle finishes in seconds,
  (3) size-independent properties at larger sizes.
Stated tolerances: fp32 engine vs fp64 reference: relative Frobenius error of the
updated factor matrix <= 1e-5 per half-iteration, loss relative error <= 1e-5; fp64 kernels: <= 1e-9
(summation order only).  CSR indexing is exact: every row's result depends only on its own indices."""
import numpy as np
import pytest

import oracle
import wrmf_cases as wc
from rsparse_b200 import WRMF, Session, als_explicit, als_implicit, gram
from rsparse_b200 import _lib as L

pytestmark = pytest.mark.gpu

CASES = sorted(wc.half_iteration_cases().keys())
TOL_F32 = 1e-5
TOL_F64 = 1e-9


def relF(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_stateless(c, dt, XtX="engine"):
    X = c["X"].astype(dt)
    Y = c["Y0"].astype(dt).copy()
    if c["feedback"] == "implicit":
        G = None
        if XtX == "host":  # what R does: tcrossprod(X) + lambda*I in the working precision (R/model_WRMF.R:474-486)
            G = (X.T @ X + c["lam"] * np.eye(X.shape[1], dtype=dt)).astype(dt)
        loss = als_implicit(c["ptr"], c["idx"], c["val"], X, Y, c["lam"], c["solver"], c["cg_steps"], XtX=G)
    else:
        loss = als_explicit(c["ptr"], c["idx"], c["val"], X, Y, c["cnt_X"], c["lam"], c["solver"], c["cg_steps"],
                            c["dynamic_lambda"])
    return Y, loss


@pytest.mark.parametrize("name", CASES)
def test_f64_kernels_vs_reference_golden(name, cases, golden_half):
    Y, loss = run_stateless(cases[name], np.float64, XtX="host")
    assert relF(Y, golden_half[name + "/Y_f64"]) < TOL_F64
    assert abs(loss - float(golden_half[name + "/loss_f64"])) <= TOL_F64 * abs(loss)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("xtx", ["host", "engine"])
def test_f32_engine_vs_reference_golden(name, xtx, cases, golden_half):
    Y, loss = run_stateless(cases[name], np.float32, XtX=xtx)
    ref = golden_half[name + "/Y_f64"]
    tol = TOL_F32
    if cases[name]["solver"] == wc.NNLS:
        # an iteration stopped at a relative coordinate step of 1e-4 (nnls.hpp:44): the reference's own fp32 run is
        # 1.1e-5 away from its fp64 run here; allow three times that
        tol = max(1e-4, 3 * relF(golden_half[name + "/Y_f32"], ref))
    assert relF(Y, ref) < tol, (relF(Y, ref), relF(golden_half[name + "/Y_f32"], ref))
    assert abs(loss - float(golden_half[name + "/loss_f64"])) <= TOL_F32 * abs(loss)
    empty = np.diff(cases[name]["ptr"]) == 0
    assert np.all(Y[empty] == 0)


def _session_for(c, kernel, solver=None, stage=0, ctas=0):
    n_src, k = c["X"].shape
    n_tgt = c["Y0"].shape[0]
    s = Session(None, (c["ptr"], c["idx"], c["val"]), n_tgt, n_src, k, c["feedback"], c["solver"] if solver is None else solver,
                c["cg_steps"], c["dynamic_lambda"], c["lam"], kernel, stage, ctas)
    s.set_factors(L.ITEMS, c["X"])
    s.set_factors(L.USERS, c["Y0"])
    return s


@pytest.mark.parametrize("name", [n for n in CASES if "k128" in n and "_cg_" in n])
@pytest.mark.parametrize("kernel,stage", [(1, 0), (2, 1), (2, 2), (3, 1), (3, 2)])
def test_cg_kernel_variants_agree_with_reference(name, kernel, stage, cases, golden_half):
    """generic streaming (1), register-resident with full XtX (2), register-resident in the eigenbasis (3);
    tile staging by cp.async.bulk (stage 1) or cp.async (stage 2)."""
    c = cases[name]
    if kernel == 3 and c["feedback"] != "implicit":
        pytest.skip("eigenbasis path: implicit feedback")
    s = _session_for(c, kernel, stage=stage)
    loss = s.half_iteration(L.USERS)
    Y = s.get_factors(L.USERS)
    X = s.get_factors(L.ITEMS)
    s.close()
    ref = golden_half[name + "/Y_f64"]
    assert relF(Y, ref) < TOL_F32
    assert abs(loss - float(golden_half[name + "/loss_f64"])) <= TOL_F32 * abs(loss)
    assert relF(X, c["X"]) < 2e-6   # the fixed matrix comes back unchanged (rotated out of the eigenbasis for kernel 3)


@pytest.mark.parametrize("name", [n for n in CASES if "_cg" in n and wc.half_iteration_cases()[n]["X"].shape[1] % 4 == 0])
@pytest.mark.parametrize("kernel", [10, 0])
def test_tile_cg_kernel_in_session(name, kernel, cases, golden_half):
    """als_cg_tile_kernel (shared-memory tile, any rank % 4 == 0 up to 256, rows binned by length) through the
    device-resident session: kernel = 10 keeps the register-resident kernel out and forces the eigenbasis for implicit
    feedback (tile kernel with `diag`; rows beyond the largest tile class take the streaming kernel with `diag`),
    kernel = 0 is the automatic choice.  Ranks 8 ... 256, ragged / empty / long rows, implicit and explicit."""
    c = cases[name]
    s = _session_for(c, kernel)
    loss = s.half_iteration(L.USERS)
    Y = s.get_factors(L.USERS)
    X = s.get_factors(L.ITEMS)
    s.close()
    ref = golden_half[name + "/Y_f64"]
    assert relF(Y, ref) < TOL_F32, relF(Y, ref)
    # the session counts cnt_X from the matrix itself; golden cases built with a stand-in cnt_X differ in the regulariser
    true_cnt = c["cnt_X"] is None or np.array_equal(c["cnt_X"], np.bincount(c["idx"], minlength=c["X"].shape[0]))
    if c["feedback"] == "implicit" or true_cnt:
        assert abs(loss - float(golden_half[name + "/loss_f64"])) <= TOL_F32 * abs(loss)
    assert relF(X, c["X"]) < 2e-6
    assert np.all(Y[np.diff(c["ptr"]) == 0] == 0)


@pytest.mark.parametrize("k,feedback,gram_rows,clusters", [(128, "implicit", "1", "1"), (128, "explicit", "1", "1"), (128, "implicit", "0", "8"),
                                                           (128, "explicit", "0", "8"), (256, "implicit", "1", "8"), (64, "implicit", "1", "8"),
                                                           (128, "implicit", "0", "1"), (256, "implicit", "1", "1"),
                                                           (128, "implicit-low", "1", "1")])
@pytest.mark.parametrize("kernel", [10, 2])
def test_long_rows_gram_and_cluster_kernels(k, feedback, gram_rows, clusters, kernel, monkeypatch):
    """Rows too long for one CTA's tile buffers.  Rank 128: als_cg_gram_kernel -- the row's 128 x 128 system is formed on
    tcgen05 in one pass over the tile (3xTF32), then the reference's CG steps run on it.  Other ranks, or with
    B200ALS_GRAM_ROWS=0: the streaming kernel, or with B200ALS_TILE_CLUSTER=8 thread-block clusters of 2 / 4 / 8 CTAs (slabs of
    the tile in each CTA's shared memory, per-sweep sums over distributed shared memory).  Ragged rows of 1 ... 2399 entries
    against the fp64 oracle; kernel = 10: eigenbasis forced for implicit feedback, kernel = 2: full XtX (rank 128)."""
    if kernel == 2 and k != 128:
        pytest.skip("kernel = 2 is a rank-128 option")
    monkeypatch.setenv("B200ALS_GRAM_ROWS", gram_rows)
    monkeypatch.setenv("B200ALS_TILE_CLUSTER", clusters)      # clusters are opt-in (profiles/r2/tile_ab.txt)
    n_rows, n_src, lam = 40, 3000, 0.1
    low = (feedback == "implicit-low")      # confidences below 1: c - 1 < 0, the symmetric Gram-rows kernel must step aside
    feedback = "implicit" if low else feedback
    ptr, idx, val = wc.det_csr(n_rows, n_src, 1200, 77 + k, ragged=True, explicit=(feedback == "explicit"))
    if low:
        val = val * 0.25
    X = np.ascontiguousarray(wc.det_factors(n_src, k, 300 + k, 0.1) * (1.0 + np.arange(k, dtype=np.float32)) ** -0.5)
    Y0 = wc.det_factors(n_rows, k, 301 + k)
    X64, Yo = X.astype(np.float64), Y0.astype(np.float64)
    if feedback == "implicit":
        lo = oracle.als_implicit(ptr, idx, val, X64, Yo, X64.T @ X64 + lam * np.eye(k), lam, wc.CG, 3, 2)
    else:
        cnt = np.bincount(idx, minlength=n_src).astype(np.float64)
        lo = oracle.als_explicit(ptr, idx, val, X64, Yo, cnt, lam, wc.CG, 3, True, 2)
    s = Session(None, (ptr, idx, val), n_rows, n_src, k, feedback, wc.CG, 3, True, lam, kernel)
    s.set_factors(L.ITEMS, X)
    s.set_factors(L.USERS, Y0)
    loss = s.half_iteration(L.USERS)
    Y = s.get_factors(L.USERS)
    plan = s.row_plan(L.USERS)["rows"]
    s.close()
    n_clu = plan["cluster2"] + plan["cluster4"] + plan["cluster8"]
    if (k == 128 and gram_rows == "1") or clusters == "1":
        assert plan["long"] > 0 and n_clu == 0, plan          # Gram-rows kernel (rank 128) or the streaming kernel
    else:
        assert n_clu > 0, plan
    assert relF(Y, Yo) < TOL_F32, (relF(Y, Yo), plan)
    assert abs(loss - lo) <= TOL_F32 * abs(lo)


def test_row_results_depend_only_on_their_own_indices(cases):
    """Bit-exact CSR indexing: permuting the order rows are presented in permutes the result rows bitwise."""
    c = cases["synth_ragged_implicit_cg_k128"]
    n = c["Y0"].shape[0]
    perm = np.argsort(wc.splitmix64(np.arange(n, dtype=np.uint64)))
    lens = np.diff(c["ptr"])
    ptr2 = np.zeros(n + 1, np.int32)
    ptr2[1:] = np.cumsum(lens[perm])
    idx2 = np.concatenate([c["idx"][c["ptr"][r]:c["ptr"][r + 1]] for r in perm]).astype(np.int32)
    val2 = np.concatenate([c["val"][c["ptr"][r]:c["ptr"][r + 1]] for r in perm])
    X = c["X"]
    G = gram(X, c["lam"])
    Y1 = c["Y0"].copy()
    als_implicit(c["ptr"], c["idx"], c["val"], X, Y1, c["lam"], wc.CG, 3, XtX=G)
    Y2 = np.ascontiguousarray(c["Y0"][perm])
    als_implicit(ptr2, idx2, val2, X, Y2, c["lam"], wc.CG, 3, XtX=G)
    assert np.array_equal(Y1[perm], Y2)


def test_gram_kernel_vs_oracle():
    for (n, k) in ((5000, 128), (777, 16), (3000, 200)):
        X = wc.det_factors(n, k, 5 + k)
        G = gram(X, 0.3)
        ref = X.astype(np.float64).T @ X.astype(np.float64) + 0.3 * np.eye(k)
        assert np.allclose(G, ref, rtol=3e-6, atol=1e-7)
        assert np.array_equal(G, G.T)
        assert np.array_equal(G, gram(X, 0.3))   # fixed-order reduction => reproducible


@pytest.mark.parametrize("name,kernel,warm", [("ml100k_implicit_cg_k16", 0, "1"), ("ml100k_implicit_chol_k8", 0, "1"),
                                              ("ml100k_explicit_cg_k8", 0, "1"), ("ml100k_implicit_cg_k16", 10, "1"),
                                              ("ml100k_implicit_cg_k16", 10, "0"), ("ml100k_implicit_cg_k16", 3, "1")])
def test_session_fit_matches_reference_trace(name, kernel, warm, golden_traces, monkeypatch):
    """b200als_fit (the loop of R/model_WRMF.R:318-338) + transform_ vs the reference-generated trace.  kernel = 10 / 3 force the
    eigenbasis of XtX in every half-iteration of this small problem (ITEM and USER halves: the basis is re-diagonalised six
    times, warm-started from each side's previous eigenvectors unless B200ALS_EIG_WARM=0)."""
    monkeypatch.setenv("B200ALS_EIG_WARM", warm)
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    U0, I0 = golden_traces[name + "/U0"], golden_traces[name + "/I0"]
    k = U0.shape[1]
    s = Session(items, users, M.shape[0], M.shape[1], k, str(golden_traces[name + "/feedback"]),
                int(golden_traces[name + "/solver"]), 3, True, float(golden_traces[name + "/lam"]), kernel)
    s.set_factors(L.USERS, U0)
    s.set_factors(L.ITEMS, I0)
    trace, done = s.fit(3, -1.0)
    assert done == 3
    ref_losses = golden_traces[name + "/losses_f64"]
    assert np.allclose(trace, ref_losses, rtol=5e-5)
    comp = s.get_factors(L.ITEMS)
    emb, _ = s.transform()
    s.close()
    # three full ALS iterations compound the per-half-iteration fp32 error
    assert relF(comp, golden_traces[name + "/components_f64"]) < 2e-4
    assert relF(emb, golden_traces[name + "/user_emb_f64"]) < 2e-4


@pytest.mark.parametrize("precision", ["float", "double"])
@pytest.mark.parametrize("feedback,solver,lam", [
    ("implicit", "conjugate_gradient", 0.1), ("implicit", "cholesky", 0.1), ("implicit", "nnls", 0.1),
    ("explicit", "conjugate_gradient", 0.1), ("explicit", "cholesky", 0.1), ("explicit", "nnls", 0.1),
    ("implicit", "cholesky", 0.0), ("implicit", "conjugate_gradient", 1000.0), ("implicit", "nnls", 1000.0),
    ("explicit", "cholesky", 1000.0), ("explicit", "nnls", 1000.0)])
def test_wrmf_class_like_reference_tests(precision, feedback, solver, lam):
    """tests/testthat/test-wrmf.R:9-71: shapes, fit_transform(train) == transform(train), transform(cv) shape,
    non-negative embeddings for solver = "nnls"."""
    M = wc.load_movielens()
    train, cv = M[:900], M[900:]
    model = WRMF(rank=8, lambda_=lam, feedback=feedback, solver=solver, precision=precision, seed=1)
    emb = model.fit_transform(train, n_iter=5, convergence_tol=-1)
    assert emb.shape == (900, 8)
    assert model.components.shape == (8, M.shape[1])
    emb2 = model.transform(train)
    # nnls is a coordinate descent stopped at a relative step of 1e-4 (inst/include/nnls.hpp:44) from a different
    # start in transform() than in the last fit sweep: equal only to that tolerance.
    tol = 2e-3 if solver == "nnls" else (3e-5 if precision == "float" else 1e-9)
    assert relF(emb2, emb) < tol
    emb_cv = model.transform(cv)
    assert emb_cv.shape == (cv.shape[0], 8)
    assert np.all(np.isfinite(emb))
    if solver == "nnls":
        assert np.all(emb >= 0) and np.all(emb_cv >= 0) and np.all(model.components >= 0)


def test_wrmf_rejects_what_the_reference_rejects():
    M = wc.load_movielens()
    with pytest.raises(ValueError):
        WRMF(solver="lbfgs")
    with pytest.raises(TypeError):
        WRMF(cg_steps=3.0)
    neg = M.copy().astype(np.float64)
    neg.data[0] = -1.0
    with pytest.raises(ValueError):
        WRMF(rank=4, feedback="implicit", precision="float").fit_transform(neg, n_iter=1)
    m = WRMF(rank=4, precision="float")
    m.fit_transform(M[:50], n_iter=1)
    with pytest.raises(ValueError):
        m.transform(M[:10, :100])


def test_larger_synthetic_vs_oracle_all_cg_paths():
    """60k x 20k, 80 nnz/row, rank 128 (a C3-shaped slice): eigenbasis kernel, full-XtX kernel and the
    streaming kernel against the fp32 oracle on the same generator output."""
    n_user, n_item, nnz, k, lam = 60000, 20000, 80, 128, 0.1
    ptr = np.zeros(n_user + 1, np.int32)
    idx = np.zeros(n_user * nnz, np.int32)
    v64 = np.zeros(n_user * nnz, np.float64)
    L.check(L.lib().b200als_synth_csr_host(n_user, n_item, nnz, 42, 0, 0, L.vp(ptr), L.vp(idx), None, L.vp(v64)))
    # trained-like item factors (decaying spectrum): every row takes all three CG steps
    X = np.ascontiguousarray(wc.det_factors(n_item, k, 901, 0.1) * (1.0 + np.arange(k, dtype=np.float32)) ** -0.5)
    Y0 = wc.det_factors(n_user, k, 902)
    G = oracle.gram(X, lam)
    Yo = Y0.copy()
    lo = oracle.als_implicit(ptr, idx, v64, X, Yo, G, lam, wc.CG, 3, oracle.max_threads())
    Y2 = Y0[:500].copy()
    oracle.als_implicit(ptr[:501], idx[:ptr[500]], v64[:ptr[500]], X, Y2, G, lam, wc.CG, 2, 1)
    assert relF(Y2, Yo[:500]) > 1e-4   # the third CG step matters on this input
    for kernel in (1, 2, 3):
        s = Session.synthetic(n_user, 0, n_user, n_item, nnz, 42, k, "implicit", L.CONJUGATE_GRADIENT, 3, True, lam, kernel)
        s.set_factors(L.ITEMS, X)
        s.set_factors(L.USERS, Y0)
        loss = s.half_iteration(L.USERS)
        Y = s.get_factors(L.USERS)
        s.close()
        assert relF(Y, Yo) < TOL_F32, kernel
        assert abs(loss - lo) <= TOL_F32 * abs(lo), kernel


def test_explicit_larger_synthetic_vs_oracle():
    n_user, n_item, nnz, k, lam = 30000, 8000, 80, 128, 0.1
    ptr = np.zeros(n_user + 1, np.int32)
    idx = np.zeros(n_user * nnz, np.int32)
    v64 = np.zeros(n_user * nnz, np.float64)
    L.check(L.lib().b200als_synth_csr_host(n_user, n_item, nnz, 7, 1, 0, L.vp(ptr), L.vp(idx), None, L.vp(v64)))
    X = wc.det_factors(n_item, k, 911)
    Y0 = wc.det_factors(n_user, k, 912)
    cnt = np.bincount(idx, minlength=n_item).astype(np.float32)
    Yo = Y0.copy()
    lo = oracle.als_explicit(ptr, idx, v64, X, Yo, cnt, lam, wc.CG, 3, True, oracle.max_threads())
    Y = Y0.copy()
    loss = als_explicit(ptr, idx, v64, X, Y, cnt, lam, wc.CG, 3, True)
    assert relF(Y, Yo) < TOL_F32
    assert abs(loss - lo) <= TOL_F32 * abs(lo)


def test_unsupported_options_fail_loudly(cases):
    c = cases["synth_cg_early_exit_k16"]
    with pytest.raises(L.B200AlsError) as e:
        als_implicit(c["ptr"], c["idx"], c["val"], c["X"], c["Y0"].copy(), 0.1, 7)
    assert e.value.code == L.EINVAL
    with pytest.raises(L.B200AlsError) as e:   # bias layouts need the two extra rows
        als_implicit(c["ptr"], c["idx"], c["val"], c["X"][:, :1].copy(), c["Y0"][:, :1].copy(), 0.1, L.CHOLESKY,
                     with_user_item_bias=True)
    assert e.value.code == L.EINVAL


@pytest.mark.parametrize("name", ["synth_ragged_implicit_cg_k128", "synth_long_implicit_cg_k128", "synth_explicit_cg_k128",
                                  "synth_implicit_cg_k128"])
@pytest.mark.parametrize("vals", ["f64", "f32"])
def test_pipelined_stateless_path(name, vals, cases, golden_half, monkeypatch):
    """The chunked H2D / solve / D2H pipeline of the stateless call (used for >= 200k rows), forced on small
    inputs with 97-row blocks: empty rows, rows longer than the register tile, explicit feedback."""
    import os
    monkeypatch.setenv("B200ALS_PIPELINE", "1")
    monkeypatch.setenv("B200ALS_PIPELINE_ROWS", "97")
    c = dict(cases[name])
    if vals == "f32":
        c["val"] = c["val"].astype(np.float32)
    for xtx in ("host", "engine"):
        Y, loss = run_stateless(c, np.float32, XtX=xtx)
        ref = golden_half[name + "/Y_f64"]
        assert relF(Y, ref) < TOL_F32, (xtx, relF(Y, ref))
        assert abs(loss - float(golden_half[name + "/loss_f64"])) <= TOL_F32 * abs(loss)
        assert np.all(Y[np.diff(c["ptr"]) == 0] == 0)


_CHOL_CASES = ["synth_implicit_cg_k128", "synth_ragged_implicit_cg_k128", "synth_explicit_cg_k128", "synth_ragged_explicit_cg_k64",
               "synth_long_implicit_cg_k128", "synth_implicit_chol_k64"]


@pytest.mark.parametrize("name,kernel,ctas", [(n, 0, 0) for n in _CHOL_CASES] + [(n, 4, 0) for n in _CHOL_CASES] +
                         [("synth_implicit_cg_k128", 4, 2), ("synth_ragged_implicit_cg_k128", 4, 2), ("synth_explicit_cg_k128", 4, 2),
                          ("synth_implicit_chol_k64", 1, 0)])
def test_tiled_cholesky_vs_oracle(name, kernel, ctas, cases):
    """The rank-64/128 Cholesky kernels for rows <= 80 nnz (longer and empty rows go through the generic kernel) against
    the fp64 oracle, implicit and explicit.  kernel = 0, the defaults: rank 128 -- row-per-thread panel Cholesky with the
    per-row Gram on tcgen05 (3xTF32, TMEM accumulator), rank 64 -- warp per system; kernel = 4: the FFMA2-Gram
    row-per-thread kernel at both ranks (also its 2-CTA/SM build at rank 128); kernel = 1: the generic kernel."""
    c = dict(cases[name])
    X64, Y64 = c["X"].astype(np.float64), c["Y0"].astype(np.float64).copy()
    if c["feedback"] == "implicit":
        G = X64.T @ X64 + c["lam"] * np.eye(X64.shape[1])
        lo = oracle.als_implicit(c["ptr"], c["idx"], c["val"], X64, Y64, G, c["lam"], wc.CHOL, 3, 2)
    else:
        # a session counts cnt_X itself (nnz per row of the fixed matrix, R/model_WRMF.R:305-315)
        cnt = np.bincount(c["idx"], minlength=X64.shape[0]).astype(np.float64)
        lo = oracle.als_explicit(c["ptr"], c["idx"], c["val"], X64, Y64, cnt, c["lam"], wc.CHOL, 3, c["dynamic_lambda"], 2)
    s = _session_for(c, kernel, solver=wc.CHOL, ctas=ctas)
    loss = s.half_iteration(L.USERS)
    Y = s.get_factors(L.USERS)
    s.close()
    assert relF(Y, Y64) < TOL_F32, relF(Y, Y64)
    assert abs(loss - lo) <= TOL_F32 * abs(lo)
    assert np.all(Y[np.diff(c["ptr"]) == 0] == 0)


def test_singular_rows_do_not_fail_the_call():
    """The reference's default explicit configuration (lambda = 0, R/model_WRMF.R:72-83) makes every row with fewer
    entries than the rank singular; its arma::solve(lhs, rhs, fast) (wrmf_explicit.hpp:108) then returns an approximate
    solution instead of failing.  Here such a row is re-factored with a small relative diagonal shift: the call succeeds,
    results are finite and every row's residual of the normal equations is small relative to its right-hand side."""
    M = wc.load_movielens()
    ptr, idx, val = wc.targets_csc(M)
    n_item = M.shape[1]
    for dt in (np.float32, np.float64):
        X = wc.det_factors(n_item, 10, 71, 0.5).astype(dt)
        Y = wc.det_factors(M.shape[0], 10, 72, 0.5).astype(dt)
        lens = np.diff(ptr)
        assert (lens < 10).sum() == 0 or True
        # make the first 40 users short (3 entries each): singular 10 x 10 systems with lambda = 0
        keep = np.ones(len(idx), bool)
        for r in range(40):
            keep[ptr[r] + 3:ptr[r + 1]] = False
        lens2 = lens.copy()
        lens2[:40] = np.minimum(lens[:40], 3)
        ptr2 = np.zeros_like(ptr)
        ptr2[1:] = np.cumsum(lens2)
        loss = als_explicit(ptr2, idx[keep], val[keep], X, Y, np.ones(n_item, dt), 0.0, wc.CHOL, 3, True)
        assert np.isfinite(loss) and np.all(np.isfinite(Y))
        X64, Y64 = X.astype(np.float64), Y.astype(np.float64)
        for r in list(range(0, 40, 7)) + [100, 500]:
            sl = slice(ptr2[r], ptr2[r + 1])
            Xn, rr = X64[idx[keep][sl]], val[keep][sl]
            res = Xn.T @ (rr - Xn @ Y64[r])
            assert np.linalg.norm(res) <= (2e-2 if dt == np.float32 else 1e-6) * np.linalg.norm(Xn.T @ rr), (r, dt)


def test_device_side_transpose_matches_scipy_and_fit_is_bit_identical():
    """a session given ONE orientation builds the other on the device (stable radix sort); the
    result equals scipy's transpose index for index, and a fit on it is bit-identical to a fit on host-built
    orientations."""
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    k = 16
    U0, I0 = wc.det_factors(M.shape[0], k, 31), wc.det_factors(M.shape[1], k, 32)
    out = []
    for mode in ("both", "users_only", "items_only"):
        s = Session(items if mode != "users_only" else None, users if mode != "items_only" else None, M.shape[0],
                    M.shape[1], k, "implicit", L.CONJUGATE_GRADIENT, 3, True, 0.1)
        s.build_missing_orientation()
        for which, ref in ((L.ITEMS, items), (L.USERS, users)):
            ptr, idx, val = s.get_orientation(which)
            assert np.array_equal(ptr, ref[0]) and np.array_equal(idx, ref[1])
            assert np.array_equal(val, ref[2].astype(np.float32))
        s.set_factors(L.USERS, U0)
        s.set_factors(L.ITEMS, I0)
        trace, _ = s.fit(2, -1.0)
        out.append((trace, s.get_factors(L.ITEMS), s.get_factors(L.USERS)))
        s.close()
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and np.array_equal(o[1], out[0][1]) and np.array_equal(o[2], out[0][2])


def test_device_side_transpose_larger_ragged():
    ptr, idx, val = wc.det_csr(5000, 3000, 30, 77, ragged=True, empty_every=11)
    import scipy.sparse as sp
    A = sp.csr_matrix((val, idx, ptr), shape=(5000, 3000))   # users x items (rows = users)
    s = Session(None, (ptr, idx, val), 5000, 3000, 8, "implicit", L.CONJUGATE_GRADIENT, 3, True, 0.1)
    s.build_missing_orientation()
    p2, i2, v2 = s.get_orientation(L.ITEMS)
    s.close()
    T = sp.csr_matrix(A.T)
    T.sort_indices()
    assert np.array_equal(p2, T.indptr) and np.array_equal(i2, T.indices) and np.array_equal(v2, T.data.astype(np.float32))


def test_tensor_core_gram_and_rotation_paths(cases, golden_half, monkeypatch):
    """tcgen05 3xTF32 kernels (XtX and the change of basis; used automatically for >= 8k / 64k rows): forced on the
    small golden case and on a 150k-row synthetic, against the reference golden / the fp32 oracle."""
    monkeypatch.setenv("B200ALS_ROTATE", "tc")
    c = cases["synth_implicit_cg_k128"]
    s = _session_for(c, 3)
    loss = s.half_iteration(L.USERS)
    Y, X = s.get_factors(L.USERS), s.get_factors(L.ITEMS)
    s.close()
    assert relF(Y, golden_half["synth_implicit_cg_k128/Y_f64"]) < TOL_F32
    assert abs(loss - float(golden_half["synth_implicit_cg_k128/loss_f64"])) <= TOL_F32 * abs(loss)
    assert relF(X, c["X"]) < 2e-6
    # larger: Gram on tensor cores too (n_item >= 8192)
    n_user, n_item, nnz, k, lam = 150000, 70000, 80, 128, 0.1
    ptr = np.zeros(n_user + 1, np.int32)
    idx = np.zeros(n_user * nnz, np.int32)
    v64 = np.zeros(n_user * nnz, np.float64)
    L.check(L.lib().b200als_synth_csr_host(n_user, n_item, nnz, 43, 0, 0, L.vp(ptr), L.vp(idx), None, L.vp(v64)))
    X = np.ascontiguousarray(wc.det_factors(n_item, k, 903, 0.1) * (1.0 + np.arange(k, dtype=np.float32)) ** -0.5)
    Y0 = wc.det_factors(n_user, k, 904)
    Yo = Y0.copy()
    lo = oracle.als_implicit(ptr, idx, v64, X, Yo, oracle.gram(X, lam), lam, wc.CG, 3, oracle.max_threads())
    res = {}
    for mode in ("tc", "ffma"):
        monkeypatch.setenv("B200ALS_ROTATE", mode)
        monkeypatch.setenv("B200ALS_GRAM", mode)
        s = Session.synthetic(n_user, 0, n_user, n_item, nnz, 43, k, "implicit", L.CONJUGATE_GRADIENT, 3, True, lam, 3)
        s.set_factors(L.ITEMS, X)
        s.set_factors(L.USERS, Y0)
        loss = s.half_iteration(L.USERS)
        res[mode] = (s.get_factors(L.USERS), loss, s.get_factors(L.ITEMS))
        s.close()
        assert relF(res[mode][0], Yo) < TOL_F32, (mode, relF(res[mode][0], Yo))
        assert abs(loss - lo) <= TOL_F32 * abs(lo), mode
        assert relF(res[mode][2], X) < 2e-6, mode


@pytest.mark.parametrize("k", [128, 256])
def test_gram_tensor_core_blocks_and_bf16(k, monkeypatch):
    """gram_tc_blocks_kernel: XtX in 128 x 128 blocks on tcgen05 (rank 256 = three blocks of the lower triangle).
    3xTF32 is fp32-grade (<= 3e-6 against fp64); the bf16-operand mode of BASELINE configs[4] is a different,
    stated accuracy class: operands rounded to 8 significant bits, relative Frobenius error of XtX ~1e-3."""
    n = 20000
    X = np.ascontiguousarray(wc.det_factors(n, k, 700 + k, 0.1) * (1.0 + np.arange(k, dtype=np.float32)) ** -0.5)
    ref = X.astype(np.float64).T @ X.astype(np.float64) + 0.3 * np.eye(k)
    # accuracy in the Frobenius norm and against the largest entry (off-diagonal entries of a Gram are sums with
    # cancellation: their elementwise relative error is not a meaningful yardstick)
    def errs(G):
        return relF(G, ref), float(np.abs(G.astype(np.float64) - ref).max() / np.abs(ref).max())
    monkeypatch.setenv("B200ALS_GRAM", "tf32x3")
    G = gram(X, 0.3)
    assert max(errs(G)) < 3e-6 and np.array_equal(G, G.T), errs(G)
    monkeypatch.setenv("B200ALS_GRAM", "ffma")
    Gf = gram(X, 0.3)
    assert max(errs(Gf)) < 3e-6, errs(Gf)
    monkeypatch.setenv("B200ALS_GRAM", "bf16")
    Gb = gram(X, 0.3)
    eb = errs(Gb)
    assert 1e-6 < eb[0] < 5e-3 and eb[1] < 5e-3, eb     # bf16 operands: visibly not fp32-grade, but a usable Gram
    assert np.array_equal(Gb, Gb.T)


def _topk_case(n_user, n_item, rank, seed, density=0.1):
    import scipy.sparse as sp
    x = wc.det_factors(n_user, rank, seed, 1.0)
    y = wc.det_factors(n_item, rank, seed + 1, 1.0)
    u = wc.det_uniform(n_user * n_item, seed + 2).reshape(n_user, n_item)
    nr = sp.csr_matrix((u < density).astype(np.float64))
    nr.sort_indices()
    return x, y, nr


@pytest.mark.parametrize("n_user,n_item,rank,k", [(100, 50, 10, 10), (70, 333, 16, 7), (33, 1000, 128, 40), (5, 20, 4, 20),
                                                (40, 700, 256, 25), (37, 300, 200, 128)])
def test_top_product_matches_reference_semantics(n_user, n_item, rank, k):
    """`top_product` (src/matrix_top_product.cpp:20-102): indices bit-exact against the pure-Python restatement,
    with per-user and global exclusions, NA padding when fewer than k candidates remain."""
    from oracle.topk import NA_INTEGER, top_product as ref_top
    from rsparse_b200 import top_product
    x, y, nr = _topk_case(n_user, n_item, rank, 100 + n_item)
    exclude0 = [1, 3, n_item - 1]
    for (nr_use, ex) in ((None, []), (nr, []), (nr, exclude0)):
        idx, sc = top_product(x, y, k, nr_use, ex, glob_mean=0.25 if ex else 0.0)
        ridx, rsc = ref_top(x, y, k, None if nr_use is None else nr_use.indptr, None if nr_use is None else nr_use.indices,
                            [e + 1 for e in ex], 0.25 if ex else 0.0)
        ref0 = np.where(ridx == NA_INTEGER, -1, ridx - 1)
        assert np.array_equal(idx, ref0)
        ok = ref0 >= 0
        # scores are sums of `rank` double products taken in a different order than numpy's: 1e-13 relative, or 1e-12
        # absolute for the near-zero scores a long list (k = 128 of 300 items) reaches
        assert np.allclose(sc[ok], rsc[ok], rtol=1e-13, atol=1e-12) and np.all(np.isnan(sc[~ok]))
    # the reference's own test (tests/testthat/test-top-product.R:3-14): equals order(scores, decreasing = TRUE)[1:k]
    idx, _ = top_product(x, y, min(k, n_item), None, [])
    full = x.astype(np.float64) @ y.astype(np.float64).T
    assert np.array_equal(idx[0], np.argsort(-full[0], kind="stable")[:min(k, n_item)])


@pytest.mark.parametrize("name", sorted(wc.topk_cases()))
def test_top_product_vs_reference_golden(name):
    """b200als_top_product against the outputs of the reference's own top_product (tests/golden/topk.npz, generated
    from src/matrix_top_product.cpp compiled in place): indices bit-exact -- exact ties, NA padding, per-user and
    global exclusions included -- scores to rounding."""
    import os
    from rsparse_b200 import top_product
    c = wc.topk_cases()[name]
    g = np.load(os.path.join(wc.GOLDEN, "topk.npz"))
    idx, sc = top_product(c["x"], c["y"], c["k"], c["nr"], c["exclude"], glob_mean=c["glob_mean"])
    gi = g[name + "/idx"]
    ref0 = np.where(gi == -2147483648, -1, gi - 1)
    assert np.array_equal(idx, ref0)
    ok = ref0 >= 0
    assert np.allclose(sc[ok], g[name + "/scores"][ok], rtol=1e-12, atol=0) and np.all(np.isnan(sc[~ok]))


def test_top_product_ties_follow_the_heap_rule():
    """All-zero user embeddings give equal scores everywhere: the heap keeps the FIRST k items and emits them by
    decreasing index (src/matrix_top_product.cpp:80-95)."""
    from oracle.topk import top_product as ref_top
    from rsparse_b200 import top_product
    x = np.zeros((3, 8), np.float32)
    y = wc.det_factors(40, 8, 5, 1.0)
    idx, _ = top_product(x, y, 6, None, [0])
    ridx, _ = ref_top(x, y, 6, None, None, [1])
    assert np.array_equal(idx, ridx - 1)
    assert list(idx[0]) == [6, 5, 4, 3, 2, 1]


def test_wrmf_predict_like_reference_test():
    """tests/testthat/test-wrmf.R:59-61: predict(cv, k) has shape (n_cv, k); recommended items exclude the seen ones."""
    M = wc.load_movielens()
    train, cv = M[:900], M[900:]
    model = WRMF(rank=8, lambda_=0.1, feedback="implicit", solver="conjugate_gradient", precision="float", seed=1)
    model.fit_transform(train, n_iter=3, convergence_tol=-1)
    K = 7
    preds = model.predict(cv, k=K)
    assert preds.shape == (cv.shape[0], K) and preds.min() >= 0 and preds.max() < M.shape[1]
    seen = cv.tolil().rows
    for u in range(cv.shape[0]):
        assert not set(preds[u]) & set(seen[u])
        assert len(set(preds[u])) == K
