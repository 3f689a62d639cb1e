import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rsparse_b200 import Session, _lib as L
import oracle, wrmf_cases as wc
for (n_user, n_item) in ((200000, 50000), (2000000, 1000000)):
    nnz, k, lam = 80, 128, 0.1
    s = Session.synthetic(n_user, 0, n_user, n_item, nnz, 42, k, "explicit", L.CONJUGATE_GRADIENT, 3, True, lam, 0)
    s.randomize_factors(L.ITEMS, 1234, 0.1, 0.5)
    s.randomize_factors(L.USERS, 5678, 0.01, 0.0)
    X = s.get_factors(L.ITEMS); Y0 = s.get_factors(L.USERS)
    print("inputs finite", np.isfinite(X).all(), np.isfinite(Y0).all())
    for it in range(4):
        loss = s.half_iteration(L.USERS)
        Y = s.get_factors(L.USERS)
        print(n_user, "iter", it, "loss", loss, "Y finite", np.isfinite(Y).all(), "n bad rows", int((~np.isfinite(Y).all(axis=1)).sum()), "absmax", np.nanmax(np.abs(Y)))
    if n_user <= 200000:
        ptr = np.zeros(n_user + 1, np.int32); idx = np.zeros(n_user * nnz, np.int32); v = np.zeros(n_user * nnz, np.float64)
        L.check(L.lib().b200als_synth_csr_host(n_user, n_item, nnz, 42, 1, 0, L.vp(ptr), L.vp(idx), None, L.vp(v)))
        cnt = np.bincount(idx, minlength=n_item).astype(np.float32)
        Yo = Y0.copy()
        for it in range(4):
            lo = oracle.als_explicit(ptr, idx, v, X, Yo, cnt, lam, 1, 3, True, 8)
            print("oracle iter", it, lo, np.isfinite(Yo).all())
    s.close()
