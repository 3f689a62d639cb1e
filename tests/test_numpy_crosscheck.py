"""CPU: a third, independent restatement of the half-iteration in plain numpy (its `solve`, `@` and `dot` are a real
LAPACK / BLAS) checked against the golden vectors that the reference's own headers produced on top of oracle/mini_arma.
This closes the caveat of DESIGN section 5: the dense primitives under the reference's control flow are ours -- here
they are cross-checked against LAPACK-backed arithmetic on every Cholesky and CG case (fp64, <= 1e-10).

Each function cites what it restates:
  implicit system / loss   inst/include/wrmf_implicit.hpp:207-208,231,236,259-261,286-304
  cg_solver_implicit       inst/include/wrmf_implicit.hpp:8-32
  explicit system / loss   inst/include/wrmf_explicit.hpp:78,103-108,131-132,147-173
  cg_solver_explicit       inst/include/wrmf_explicit.hpp:8-31
"""
import numpy as np
import pytest

import wrmf_cases as wc

CASES = {k: v for k, v in wc.half_iteration_cases().items() if v["solver"] in (wc.CHOL, wc.CG)}
CG_TOL = 1e-10   # inst/include/wrmf.hpp:22


def _cg(matvec, b, x, n_iter):
    r = b - matvec(x)
    p = r.copy()
    rsold = float(r @ r)
    for _ in range(n_iter):
        Ap = matvec(p)
        alpha = rsold / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rsnew = float(r @ r)
        if rsnew < CG_TOL:
            break
        p = r + p * (rsnew / rsold)
        rsold = rsnew
    return x


def numpy_half_iteration(c):
    X = c["X"].astype(np.float64)
    Y = c["Y0"].astype(np.float64).copy()
    ptr, idx, val = c["ptr"], c["idx"], c["val"].astype(np.float64)
    k = X.shape[1]
    lam = float(c["lam"])
    implicit = c["feedback"] == "implicit"
    G = X.T @ X + lam * np.eye(k) if implicit else None
    loss = 0.0
    for i in range(len(ptr) - 1):
        p1, p2 = ptr[i], ptr[i + 1]
        if p1 == p2:
            Y[i] = 0.0
            continue
        Xn = X[idx[p1:p2]]          # n x k  (X_nnz')
        cv = val[p1:p2]
        if implicit:
            if c["solver"] == wc.CHOL:
                lhs = G + (Xn.T * (cv - 1.0)) @ Xn
                y = np.linalg.solve(lhs, Xn.T @ cv)
            else:
                y = _cg(lambda v: G @ v + Xn.T @ ((cv - 1.0) * (Xn @ v)), Xn.T @ cv, Y[i].copy(), c["cg_steps"])
            loss += float(cv @ (1.0 - Xn @ y) ** 2) + lam * float(y @ y)
        else:
            lam_use = lam * (len(cv) if c["dynamic_lambda"] else 1.0)
            if c["solver"] == wc.CHOL:
                y = np.linalg.solve(Xn.T @ Xn + lam_use * np.eye(k), Xn.T @ cv)
            else:
                y = _cg(lambda v: Xn.T @ (Xn @ v) + lam_use * v, Xn.T @ cv, Y[i].copy(), c["cg_steps"])
            loss += float(np.sum((cv - Xn @ y) ** 2)) + lam_use * float(y @ y)
        Y[i] = y
    if lam > 0:
        if implicit or not c["dynamic_lambda"]:
            loss += lam * float(np.sum(X * X))
        else:
            loss += lam * float(np.sum((X * X) * c["cnt_X"].astype(np.float64)[:, None]))
    return Y, loss / len(val)


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_restatement_matches_reference_golden(name, golden_half):
    Y, loss = numpy_half_iteration(CASES[name])
    ref = golden_half[name + "/Y_f64"]
    rel = np.linalg.norm(Y - ref) / np.linalg.norm(ref)
    assert rel < 1e-10, rel
    assert abs(loss - float(golden_half[name + "/loss_f64"])) <= 1e-10 * abs(loss)
