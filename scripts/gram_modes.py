#!/usr/bin/env python
"""BASELINE configs[4] asks for "fp32 vs tensor-core bf16 Gram" at rank 256: one user half-iteration of a C5-shaped slice
(default: 500k users x 5M items x 100 nnz, rank 256, CG(3)) with XtX computed (a) on tcgen05 as 3xTF32 (fp32-grade, the
default), (b) on tcgen05 with bf16 operands, (c) by the fp32 FMA kernel.  Prints one JSON line: Gram time per mode, the
relative Frobenius distance of the solved user factors and of the loss between modes, and the distance of each mode from
the fp64 CPU oracle on a sample of rows."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=500_000)
    ap.add_argument("--items", type=int, default=5_000_000)
    ap.add_argument("--nnz", type=int, default=100)
    ap.add_argument("--rank", type=int, default=256)
    ap.add_argument("--oracle-rows", type=int, default=2000)
    a = ap.parse_args()
    import oracle
    from rsparse_b200 import Session
    from rsparse_b200 import _lib as L
    lam = 0.1
    out = {"shape": "%d x %d, %d nnz/row, rank %d, implicit CG(3)" % (a.users, a.items, a.nnz, a.rank), "modes": {}}
    Y, Xh = {}, None
    for name, code in (("tf32x3", 3), ("bf16", 1), ("ffma", 2)):
        s = Session.synthetic(a.users, 0, a.users, a.items, a.nnz, 42, a.rank, "implicit", L.CONJUGATE_GRADIENT, 3, True, lam,
                              0, 0, 0, 0, 0, code)
        s.randomize_factors(L.ITEMS, 1234, 0.1, 0.5)
        for rep in range(3):                      # last repetition is the one kept (warm clocks, allocations done)
            s.randomize_factors(L.USERS, 5678, 0.01, 0.0)
            loss = s.half_iteration(L.USERS)
            tm = s.last_timing()
        Y[name] = s.get_factors(L.USERS)
        if Xh is None:
            Xh = s.get_factors(L.ITEMS)
            s.randomize_factors(L.USERS, 5678, 0.01, 0.0)
            Y0 = s.get_factors(L.USERS)[:a.oracle_rows].copy()
        s.close()
        out["modes"][name] = {"gram_ms": tm["gram_ms"], "prep_ms": tm["prep_ms"], "solve_ms": tm["solve_ms"], "loss": loss}
    rel = lambda p, q: float(np.linalg.norm(p.astype(np.float64) - q.astype(np.float64)) / np.linalg.norm(q.astype(np.float64)))
    out["relF_user_factors"] = {"bf16_vs_tf32x3": rel(Y["bf16"], Y["tf32x3"]), "ffma_vs_tf32x3": rel(Y["ffma"], Y["tf32x3"])}
    out["loss_rel"] = {"bf16_vs_tf32x3": abs(out["modes"]["bf16"]["loss"] - out["modes"]["tf32x3"]["loss"]) / out["modes"]["tf32x3"]["loss"],
                       "ffma_vs_tf32x3": abs(out["modes"]["ffma"]["loss"] - out["modes"]["tf32x3"]["loss"]) / out["modes"]["tf32x3"]["loss"]}
    out["gram_speedup_bf16_over_tf32x3"] = out["modes"]["tf32x3"]["gram_ms"] / out["modes"]["bf16"]["gram_ms"]
    out["gram_speedup_tf32x3_over_ffma"] = out["modes"]["ffma"]["gram_ms"] / out["modes"]["tf32x3"]["gram_ms"]
    # fp64 oracle on the first rows (exact XtX in double)
    m = a.oracle_rows
    ptr, idx, val = oracle.synth_csr(m, a.items, a.nnz, 42)
    X64 = Xh.astype(np.float64)
    G = X64.T @ X64 + lam * np.eye(a.rank)
    Yo = Y0.astype(np.float64)
    oracle.als_implicit(ptr, idx, val, X64, Yo, G, lam, 1, 3, oracle.host_threads())
    out["relF_vs_fp64_oracle_first_rows"] = {k: rel(v[:m], Yo) for k, v in Y.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
