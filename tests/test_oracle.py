"""CPU: pins the oracle (oracle/wrmf_oracle.cpp) against the golden vectors produced by the
reference's own source compiled against oracle/mini_arma (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle
import wrmf_cases as wc

CASES = sorted(wc.half_iteration_cases().keys())


def _run_oracle(c, dt, n_threads=1, impl="oracle"):
    X = c["X"].astype(dt)
    Y = c["Y0"].astype(dt).copy()
    if c["feedback"] == "implicit":
        G = (X.T @ X + c["lam"] * np.eye(X.shape[1], dtype=dt)).astype(dt)
        loss = oracle.als_implicit(c["ptr"], c["idx"], c["val"], X, Y, G, c["lam"], c["solver"], c["cg_steps"],
                                   n_threads, impl=impl)
    else:
        cnt = None if c["cnt_X"] is None else c["cnt_X"].astype(dt)
        loss = oracle.als_explicit(c["ptr"], c["idx"], c["val"], X, Y, cnt, c["lam"], c["solver"], c["cg_steps"],
                                   c["dynamic_lambda"], n_threads, impl=impl)
    return Y, loss


def relF(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


@pytest.mark.parametrize("name", CASES)
def test_oracle_f64_matches_reference_golden(name, cases, golden_half):
    Y, loss = _run_oracle(cases[name], np.float64)
    assert relF(Y, golden_half[name + "/Y_f64"]) < 1e-11
    assert abs(loss - float(golden_half[name + "/loss_f64"])) <= 1e-11 * abs(loss)


@pytest.mark.parametrize("name", CASES)
def test_oracle_f32_matches_reference_golden(name, cases, golden_half):
    Y, loss = _run_oracle(cases[name], np.float32)
    # same algorithm, same type, different summation order inside the dense primitives
    assert relF(Y, golden_half[name + "/Y_f32"]) < 2e-5
    assert relF(Y, golden_half[name + "/Y_f64"]) < 2e-5
    assert abs(loss - float(golden_half[name + "/loss_f64"])) <= 2e-6 * abs(loss)


def test_oracle_is_thread_count_invariant_for_Y(cases):
    c = cases["synth_ragged_implicit_cg_k128"]
    Y1, _ = _run_oracle(c, np.float32, 1)
    Y4, _ = _run_oracle(c, np.float32, 4)
    assert np.array_equal(Y1, Y4)  # rows are independent; only the loss reduction order may differ


def test_golden_was_generated_from_reference_when_available(cases, golden_half):
    """When oracle/_ref/libref_wrmf.so is present (authoring container, or shipped with the snapshot),
    re-running the reference's compiled source reproduces the committed fixture bit for bit."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_wrmf.so not present")
    for name in ("ml100k_user_implicit_cg_k16", "synth_explicit_cg_k128", "synth_ragged_implicit_chol_k32"):
        Y, loss = _run_oracle(cases[name], np.float64, impl="ref")
        assert np.array_equal(Y, golden_half[name + "/Y_f64"])
        assert loss == float(golden_half[name + "/loss_f64"])


def test_empty_rows_are_zeroed(cases, golden_half):
    c = cases["synth_ragged_implicit_cg_k128"]
    empty = np.diff(c["ptr"]) == 0
    assert empty.any()
    assert np.all(golden_half["synth_ragged_implicit_cg_k128/Y_f64"][empty] == 0)
    Y, _ = _run_oracle(c, np.float64)
    assert np.all(Y[empty] == 0)


def test_gram_matches_numpy():
    X = wc.det_factors(5000, 48, 77)
    G = oracle.gram(X, 0.25, 3)
    ref = X.astype(np.float64).T @ X.astype(np.float64) + 0.25 * np.eye(48)
    assert np.allclose(G, ref, rtol=2e-6, atol=1e-8)


def test_als_trace_oracle_vs_reference_golden(golden_traces):
    """fit_transform flow (R/model_WRMF.R:318-359) with the oracle vs the reference-generated trace."""
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    name, k, lam = "ml100k_implicit_cg_k16", 16, 0.1
    U = golden_traces[name + "/U0"].astype(np.float64)
    I = golden_traces[name + "/I0"].astype(np.float64)
    losses = []
    for it in range(3):
        for (mat, Xf, Yf) in ((items, U, I), (users, I, U)):
            G = Xf.T @ Xf + lam * np.eye(k)
            losses.append(oracle.als_implicit(*mat, Xf, Yf, G, lam, wc.CG, 3, 2))
    assert np.allclose(losses, golden_traces[name + "/losses_f64"], rtol=1e-10)
    assert relF(I, golden_traces[name + "/components_f64"]) < 1e-10
    res = np.zeros_like(U)
    G = I.T @ I + lam * np.eye(k)
    oracle.als_implicit(*users, I, res, G, lam, wc.CHOL, 3, 2)
    assert relF(res, golden_traces[name + "/user_emb_f64"]) < 1e-10


@pytest.mark.parametrize("name", sorted(wc.topk_cases()))
def test_topk_restatement_is_pinned_to_the_reference(name):
    """oracle/topk.py against (1) the committed outputs of the reference's own top_product
    (tests/golden/topk.npz, made by tests/golden/make_golden_topk.py from src/matrix_top_product.cpp compiled in
    place) and (2) that binary itself when it is present: indices identical (ties and NA padding included),
    scores equal to rounding (the score row is a BLAS-free dot product in both, summed in a different order)."""
    import os
    from oracle.topk import top_product as py_top
    c = wc.topk_cases()[name]
    nr = c["nr"]
    args = (c["x"], c["y"], c["k"], None if nr is None else nr.indptr, None if nr is None else nr.indices,
            [e + 1 for e in c["exclude"]], c["glob_mean"])
    idx, sc = py_top(*args)
    g = np.load(os.path.join(wc.GOLDEN, "topk.npz"))
    assert np.array_equal(idx, g[name + "/idx"])
    assert np.allclose(sc, g[name + "/scores"], rtol=1e-12, atol=0, equal_nan=True)
    if oracle.ref_topk_available():
        ridx, rsc = oracle.ref_top_product(*args)
        assert np.array_equal(ridx, g[name + "/idx"]) and np.array_equal(rsc, g[name + "/scores"], equal_nan=True)
