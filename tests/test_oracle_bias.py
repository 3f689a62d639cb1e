"""CPU: pins the oracle's bias variants (als_implicit_bias / als_explicit_bias / initialize_biases in
oracle/wrmf_oracle.cpp) against golden vectors produced by the reference's own headers
(tests/golden/make_golden.py -> bias_half_iterations.npz)."""
import numpy as np
import pytest

import oracle
import wrmf_cases as wc

BIAS_CASES = sorted(wc.bias_cases().keys())


def relF(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_oracle_bias(c, dt, impl="oracle", n_threads=1):
    X = c["X"].astype(dt)
    Y = c["Y0"].astype(dt).copy()
    gbb = None
    if c["feedback"] == "implicit":
        gbb = np.zeros(X.shape[1] - int(c["with_biases"]), dt)
        loss = oracle.als_implicit_bias(c["ptr"], c["idx"], c["val"], X, Y, wc.xtx_for(c, dt), c["lam"], c["solver"],
                                        c["cg_steps"], c["with_biases"], c["is_last"], c["gbias"], gbb, True, n_threads, impl)
    else:
        loss = oracle.als_explicit_bias(c["ptr"], c["idx"], c["val"], X, Y, c["cnt_X"].astype(dt), c["lam"], c["solver"],
                                        c["cg_steps"], c["dynamic_lambda"], c["with_biases"], c["is_last"], n_threads, impl)
    return Y, loss, gbb


# nnls is an iteration stopped at a relative coordinate step of 1e-4 (nnls.hpp:44): implementations that differ in
# summation order stop at slightly different iterates, fp32 ones visibly so when the system holds the column of ones
def tol_for(c, dt):
    if c["solver"] == wc.NNLS:
        return (1e-7, 1e-9) if dt == np.float64 else (5e-2, 1e-4)
    return (1e-10, 1e-11) if dt == np.float64 else (1e-4, 5e-6)


@pytest.mark.parametrize("name", BIAS_CASES)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_oracle_bias_variants_match_reference_golden(name, dt, bias_cases, golden_bias):
    c = bias_cases[name]
    Y, loss, gbb = run_oracle_bias(c, dt)
    tag = "f64" if dt == np.float64 else "f32"
    tol_y, tol_l = tol_for(c, dt)
    assert relF(Y, golden_bias[name + "/Y_" + tag]) < tol_y
    assert relF(Y, golden_bias[name + "/Y_f64"]) < max(tol_y, 1e-4 if dt == np.float32 else 0)
    assert abs(loss - float(golden_bias[name + "/loss_f64"])) <= max(tol_l, 1e-5 if dt == np.float32 else 0) * abs(loss)
    if gbb is not None and c["gbias"] and not c["with_biases"]:
        assert relF(gbb, golden_bias[name + "/gbb_f64"]) < (1e-12 if dt == np.float64 else 1e-5)
    # the untouched row of Y (the ones) and the layout
    if c["with_biases"]:
        col = -1 if c["is_last"] else 0
        assert np.all(Y[:, col] == 1)


def test_bias_golden_was_generated_from_reference_when_available(bias_cases, golden_bias):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_wrmf.so not present")
    for name in ("bias_ml100k_user_implicit_chol", "global_ml100k_user_implicit_cg", "bias_rag_explicit_chol"):
        Y, loss, _ = run_oracle_bias(bias_cases[name], np.float64, impl="ref")
        assert np.array_equal(Y, golden_bias[name + "/Y_f64"])
        assert loss == float(golden_bias[name + "/loss_f64"])


def test_rows_without_entries_are_solved_with_bias_terms(bias_cases, golden_bias):
    """wrmf_implicit.hpp:179: with biases / a global bias an empty column still gets rhs_init solved against XtX."""
    c = bias_cases["bias_rag_implicit_chol_global"]
    empty = np.diff(c["ptr"]) == 0
    assert empty.any()
    Y = golden_bias["bias_rag_implicit_chol_global/Y_f64"]
    assert np.all(np.abs(Y[empty][:, 1:]).sum(axis=1) > 0)
    c = bias_cases["bias_rag_explicit_chol"]
    empty = np.diff(c["ptr"]) == 0
    Y = golden_bias["bias_rag_explicit_chol/Y_f64"]
    assert np.all(Y[empty][:, 1:] == 0) and np.all(Y[empty][:, 0] == 1)   # wrmf_explicit.hpp:134-145


@pytest.mark.parametrize("is_explicit", [True, False])
@pytest.mark.parametrize("non_negative", [False, True])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_oracle_initialize_biases_matches_reference_golden(is_explicit, non_negative, dt, golden_bias):
    M = wc.load_movielens()
    users, items = wc.targets_csc(M), wc.targets_csc(M.T)
    csc = (items[0], items[1], items[2].copy())
    csr = (users[0], users[1], users[2].copy())
    ub, ib = np.zeros(M.shape[0], dt), np.zeros(M.shape[1], dt)
    g = oracle.initialize_biases(csc, csr, ub, ib, 0.1, True, non_negative, True, is_explicit)
    key = "init_%s_nn%d_%s" % ("explicit" if is_explicit else "implicit", int(non_negative), "f64" if dt == np.float64 else "f32")
    assert np.array_equal(ub, golden_bias[key + "/user_bias"])
    assert np.array_equal(ib, golden_bias[key + "/item_bias"])
    assert g == float(golden_bias[key + "/global_bias"])
    if is_explicit:   # values are shifted by the mean rating in place (wrmf_utils.hpp:48-51)
        assert np.allclose(csc[2], items[2] - g) and np.allclose(csr[2], users[2] - g)
