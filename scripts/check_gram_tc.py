import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wrmf_cases as wc
from rsparse_b200 import gram, Session, _lib as L
for n in (5000, 100000, 1000003):
    X = np.ascontiguousarray(wc.det_factors(n, 128, 5, 0.1) * (1.0 + np.arange(128, dtype=np.float32)) ** -0.5)
    ref = X.astype(np.float64).T @ X.astype(np.float64) + 0.3 * np.eye(128)
    res = {}
    for mode in ("tc", "ffma"):
        os.environ["B200ALS_GRAM"] = mode
        G = gram(X, 0.3)
        err = np.abs(G - ref) / (np.sqrt(np.outer(np.diag(ref), np.diag(ref))))
        res[mode] = G
        print(n, mode, "max scaled err %.3e" % err.max(), "max rel err on diag %.3e" % (np.abs(np.diag(G) - np.diag(ref)) / np.diag(ref)).max(), "finite", np.isfinite(G).all())
    print("  tc vs ffma max scaled diff %.3e" % (np.abs(res["tc"] - res["ffma"]) / np.sqrt(np.outer(np.diag(ref), np.diag(ref)))).max())
# timing inside a session step
for mode in ("tc", "ffma"):
    os.environ["B200ALS_GRAM"] = mode
    s = Session.synthetic(1000000, 0, 1000000, 1000000, 80, 42, 128, "implicit", L.CONJUGATE_GRADIENT, 3, True, 0.1, 0)
    s.randomize_factors(L.ITEMS, 1, 0.1, 0.5); s.randomize_factors(L.USERS, 2, 0.01, 0)
    for i in range(3):
        loss = s.half_iteration(L.USERS)
    print(mode, "loss", loss, s.last_timing())
    s.close()
