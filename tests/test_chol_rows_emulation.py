"""CPU: lane-level emulation of the index logic of als_chol_rows_kernel (thread r owns row r, 4-column panels, the
left-shifting register update, never-loaded upper blocks, blocked back substitution) -- scripts/emulate_chol_rows.py --
against numpy's solve.  It pins the DESIGN of the kernel (the emulation poisons every register the kernel never loads
with NaN, so a wrong trip count or shift shows up as NaN); the kernel itself is checked on the GPU in test_gpu_parity."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from emulate_chol_rows import emulate  # noqa: E402


@pytest.mark.parametrize("K,n", [(64, 50), (64, 3), (128, 80), (128, 1)])
def test_row_panel_cholesky_index_logic(K, n):
    rng = np.random.default_rng(K + n)
    X = (rng.standard_normal((n, K)) * 0.1).astype(np.float32)
    w = rng.integers(1, 10, n).astype(np.float32)
    G = (rng.standard_normal((500, K)) * 0.1).astype(np.float32)
    A = (G.T @ G + 0.1 * np.eye(K) + (X.T * w) @ X).astype(np.float32)
    b = (X.T @ (w + 1)).astype(np.float32)
    y = emulate(K, A, b)
    ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
    assert np.isfinite(y).all()
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-5
