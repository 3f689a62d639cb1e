"""Worker for tests/test_multigpu.py and scripts: run under torchrun with one process per GPU.
Row-sharded user half-iteration (exchange of the solved rows inside libb200als.so: peer-memory pushes over NVLink, or
NCCL broadcasts with B200ALS_EXCHANGE=nccl) checked against the CPU oracle, on every rank's copy of the factors."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle
    import wrmf_cases as wc
    from rsparse_b200 import Session, parallel
    from rsparse_b200 import _lib as L
    rank, world, local_rank = parallel.env_rank_world()
    L.check(L.lib().b200als_set_device(local_rank))
    parallel.init_engine_comm()
    n_user, n_item, nnz, k, lam = 40000 * world, 20000, 80, 128, 0.1   # >= 32768 rows per rank: chunked exchange
    b, e = parallel.shard_range(n_user, rank, world)
    X = np.ascontiguousarray(wc.det_factors(n_item, k, 901, 0.1) * (1.0 + np.arange(k, dtype=np.float32)) ** -0.5)
    Y0 = wc.det_factors(n_user, k, 902)
    results = {}
    for kernel in (3, 2, 1):
        s = Session.synthetic(e - b, b, n_user, n_item, nnz, 42, k, "implicit", L.CONJUGATE_GRADIENT, 3, True, lam, kernel)
        s.set_factors(L.ITEMS, X)
        s.set_factors(L.USERS, Y0)
        loss1 = s.half_iteration(L.USERS)
        loss2 = s.half_iteration(L.USERS)      # second step: works on exchanged factors / accumulated basis
        Yk = s.get_factors(L.USERS)
        # every rank must hold the same full matrix after the exchange
        chk = float(np.abs(Yk.astype(np.float64)).sum())
        assert parallel.max_over_ranks(chk) == -parallel.max_over_ranks(-chk), "ranks disagree on the exchanged factors"
        results[kernel] = (loss1, loss2, Yk, s.last_timing())
        mode = s.exchange_mode()
        want = os.environ.get("B200ALS_EXCHANGE")
        assert mode in ("p2p", "nccl") and (want is None or mode == want), (mode, want)
        s.close()
    # explicit feedback (no Gram all-reduce in front of the exchange)
    se = Session.synthetic(e - b, b, n_user, n_item, nnz, 43, k, "explicit", L.CONJUGATE_GRADIENT, 3, True, lam, 0)
    se.set_factors(L.ITEMS, X)
    se.set_factors(L.USERS, Y0)
    le = se.half_iteration(L.USERS)
    mode = se.exchange_mode()
    Ye = se.get_factors(L.USERS)
    chk = float(np.abs(Ye.astype(np.float64)).sum())
    assert parallel.max_over_ranks(chk) == -parallel.max_over_ranks(-chk), "ranks disagree (explicit)"
    se.close()
    if rank == 0:
        ptr = np.zeros(n_user + 1, np.int32)
        idx = np.zeros(n_user * nnz, np.int32)
        v64 = np.zeros(n_user * nnz, np.float64)
        L.check(L.lib().b200als_synth_csr_host(n_user, n_item, nnz, 42, 0, 0, L.vp(ptr), L.vp(idx), None, L.vp(v64)))
        G = oracle.gram(X, lam)
        Yo = Y0.copy()
        lo1 = oracle.als_implicit(ptr, idx, v64, X, Yo, G, lam, wc.CG, 3, oracle.max_threads())
        lo2 = oracle.als_implicit(ptr, idx, v64, X, Yo, G, lam, wc.CG, 3, oracle.max_threads())
        for kernel, (l1, l2, Y, tm) in results.items():
            rel = np.linalg.norm(Y.astype(np.float64) - Yo) / np.linalg.norm(Yo)
            print("kernel %d world %d: relF %.2e loss %.7f/%.7f oracle %.7f/%.7f timing %s" % (kernel, world, rel, l1, l2, lo1, lo2, tm))
            assert rel < 2e-5 and abs(l1 - lo1) < 1e-5 * lo1 and abs(l2 - lo2) < 1e-5 * lo2
        L.check(L.lib().b200als_synth_csr_host(n_user, n_item, nnz, 43, 1, 0, L.vp(ptr), L.vp(idx), None, L.vp(v64)))
        cnt = np.bincount(idx, minlength=n_item).astype(np.float32)
        Yo = Y0.copy()
        loe = oracle.als_explicit(ptr, idx, v64, X, Yo, cnt, lam, wc.CG, 3, True, oracle.max_threads())
        rel = np.linalg.norm(Ye.astype(np.float64) - Yo) / np.linalg.norm(Yo)
        print("explicit world %d: relF %.2e loss %.7f oracle %.7f exchange=%s (requested %s)" % (
            world, rel, le, loe, mode, os.environ.get("B200ALS_EXCHANGE", "auto")))
        assert rel < 2e-5 and abs(le - loe) < 1e-5 * loe
        print("MULTIGPU_OK world=%d" % world)
    parallel.barrier()
    L.lib().b200als_comm_destroy()


if __name__ == "__main__":
    main()
