"""GPU check of als_chol_rows_kernel (session option kernel = 4) against the fp64 oracle and the tile kernel, on the
cases of tests/test_gpu_parity.py::test_tiled_cholesky_vs_oracle.  Prints one line per case; exit code 1 on a miss."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import wrmf_cases as wc  # noqa: E402
from rsparse_b200 import Session  # noqa: E402
from rsparse_b200 import _lib as L  # noqa: E402

TOL = 1e-5
CTAS = int(os.environ.get("CHOL_CTAS", "0"))
KUT = int(os.environ.get("CHOL_KERNEL", "0"))   # kernel under test: 0 defaults (rank 128: tcgen05 Gram, rank 64: warp per system), 4 FFMA2-Gram row-per-thread
cases = wc.half_iteration_cases()
bad = 0
for name in ("synth_implicit_chol_k64", "synth_implicit_cg_k128", "synth_ragged_implicit_cg_k128", "synth_explicit_cg_k128",
             "synth_ragged_explicit_cg_k64", "synth_long_implicit_cg_k128"):
    c = dict(cases[name])
    X64, Y64 = c["X"].astype(np.float64), c["Y0"].astype(np.float64).copy()
    if c["feedback"] == "implicit":
        G = X64.T @ X64 + c["lam"] * np.eye(X64.shape[1])
        lo = oracle.als_implicit(c["ptr"], c["idx"], c["val"], X64, Y64, G, c["lam"], wc.CHOL, 3, 2)
    else:
        cnt = np.bincount(c["idx"], minlength=X64.shape[0]).astype(np.float64)
        lo = oracle.als_explicit(c["ptr"], c["idx"], c["val"], X64, Y64, cnt, c["lam"], wc.CHOL, 3, c["dynamic_lambda"], 2)
    out = {}
    for kernel in (KUT, 1):
        n_src, k = c["X"].shape
        s = Session(None, (c["ptr"], c["idx"], c["val"]), c["Y0"].shape[0], n_src, k, c["feedback"], wc.CHOL, c["cg_steps"],
                    c["dynamic_lambda"], c["lam"], kernel, 0, CTAS)
        s.set_factors(L.ITEMS, c["X"])
        s.set_factors(L.USERS, c["Y0"])
        try:
            loss = s.half_iteration(L.USERS)
            Y = s.get_factors(L.USERS)
            rel = float(np.linalg.norm(Y.astype(np.float64) - Y64) / np.linalg.norm(Y64))
            out[kernel] = (rel, abs(loss - lo) / abs(lo), bool(np.all(Y[np.diff(c["ptr"]) == 0] == 0)))
        except Exception as e:  # noqa: BLE001
            out[kernel] = ("error", str(e)[:200], False)
        s.close()
    ok = out[KUT][0] != "error" and out[KUT][0] < TOL and out[KUT][1] < TOL and out[KUT][2]
    bad += not ok
    print("%-34s rows-kernel relF %s loss-rel %s zero-rows %s | generic-kernel relF %s  %s" % (
        name, out[KUT][0], out[KUT][1], out[KUT][2], out[1][0], "OK" if ok else "MISS"), flush=True)
print("CHOL_ROWS_%s" % ("OK" if bad == 0 else "FAILED"))
sys.exit(1 if bad else 0)
