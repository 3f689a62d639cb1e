"""Lane-level emulation (numpy, fp32) of als_chol_rows_kernel's index logic: thread r owns row r, panels of 4 columns,
P3 writes each updated block one block to the left.  Checks the solve against numpy on a random SPD system.  This is a
design check for the CUDA kernel (rsparse_b200/csrc/als_chol_rows.cuh), not a product path."""
import numpy as np


def emulate(K, A, b):
    f = np.float32
    LDT = K + 4
    NB4 = K // 4
    a = np.zeros((K, K), f)            # a[r][i]: register i (scalar view of the float2 pairs) of thread r
    for r in range(K):
        wmax = (r // 32) * 32 + 31
        for c in range(K):
            if 4 * (c // 4) <= wmax:
                a[r, c] = A[r, c]      # Gram result incl. upper entries inside the warp's range
            else:
                a[r, c] = np.nan       # never loaded
    br = b.astype(f).copy()
    Lt = np.full((K, LDT), np.nan, f)
    zz = np.full(K, np.nan, f)
    rs = np.full(K, np.nan, f)
    for p in range(NB4):
        j0 = 4 * p
        D = np.zeros((4, 8), f)
        for r in range(j0, j0 + 4):
            D[r - j0, 0:4] = a[r, 0:4]
            D[r - j0, 4] = br[r]
        d0, d1, d2, d3 = D[0], D[1], D[2], D[3]
        b0, b1, b2, b3 = D[0, 4], D[1, 4], D[2, 4], D[3, 4]
        i0 = f(1) / np.sqrt(d0[0]); L10 = d1[0] * i0; L20 = d2[0] * i0; L30 = d3[0] * i0
        i1 = f(1) / np.sqrt(d1[1] - L10 * L10); L21 = (d2[1] - L20 * L10) * i1; L31 = (d3[1] - L30 * L10) * i1
        i2 = f(1) / np.sqrt(d2[2] - L20 * L20 - L21 * L21); L32 = (d3[2] - L30 * L20 - L31 * L21) * i2
        i3 = f(1) / np.sqrt(d3[3] - L30 * L30 - L31 * L31 - L32 * L32)
        z0 = b0 * i0; z1 = (b1 - L10 * z0) * i1; z2 = (b2 - L20 * z0 - L21 * z1) * i2
        z3 = (b3 - L30 * z0 - L31 * z1 - L32 * z2) * i3
        zz[j0:j0 + 4] = [z0, z1, z2, z3]
        rs[j0:j0 + 4] = [i0, i1, i2, i3]
        ls = {}
        for r in range(K):
            wmax = (r // 32) * 32 + 31
            if wmax < j0:
                continue
            with np.errstate(all="ignore"):
                l0 = a[r, 0] * i0
                l1 = (a[r, 1] - l0 * L10) * i1
                l2 = (a[r, 2] - l0 * L20 - l1 * L21) * i2
                l3 = (a[r, 3] - l0 * L30 - l1 * L31 - l2 * L32) * i3
                for q, l in enumerate((l0, l1, l2, l3)):
                    Lt[j0 + q, r] = l
                br[r] = br[r] - l0 * z0 - l1 * z1 - l2 * z2 - l3 * z3
            ls[r] = (l0, l1, l2, l3)
        # barrier, then P3 (reads the complete panel)
        for r, (l0, l1, l2, l3) in ls.items():
            wmax = (r // 32) * 32 + 31
            for ib in range(NB4 - 1):
                if j0 + 4 + 4 * ib > wmax:
                    break
                c = j0 + 4 + 4 * ib
                with np.errstate(all="ignore"):
                    for e in range(4):
                        a[r, 4 * ib + e] = a[r, 4 * ib + 4 + e] - l0 * Lt[j0, c + e] - l1 * Lt[j0 + 1, c + e] \
                            - l2 * Lt[j0 + 2, c + e] - l3 * Lt[j0 + 3, c + e]
    # blocked back substitution
    for b0_ in range(K - 32, -1, -32):
        if b0_ + 32 < K:
            for i in range(b0_, b0_ + 32):
                part = f(0)
                for l in range(b0_ + 32, K):
                    part += Lt[i, l] * zz[l]
                zz[i] -= part
        zi = zz[b0_:b0_ + 32].copy()
        ri = rs[b0_:b0_ + 32]
        for s in range(31, -1, -1):
            ys = zi[s] * ri[s]
            for lane in range(s):
                zi[lane] = zi[lane] - Lt[b0_ + lane, b0_ + s] * ys
        zz[b0_:b0_ + 32] = zi * ri
    return zz


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for K in (64, 128):
        n = 50 if K == 64 else 80
        X = (rng.standard_normal((n, K)) * 0.1).astype(np.float32)
        w = rng.integers(1, 10, n).astype(np.float32)
        G = (rng.standard_normal((500, K)) * 0.1).astype(np.float32)
        A = (G.T @ G + 0.1 * np.eye(K) + (X.T * w) @ X).astype(np.float32)
        b = (X.T @ (w + 1)).astype(np.float32)
        y = emulate(K, A, b)
        ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        print(K, "relF", np.linalg.norm(y - ref) / np.linalg.norm(ref), "nan" if np.isnan(y).any() else "finite")


