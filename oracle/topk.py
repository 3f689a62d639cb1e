"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of the reference's `top_product`
(/root/reference/src/matrix_top_product.cpp:20-102), for SMALL cases only: the score row in double
(`arma::rowvec yvec = x.row(j) * y`, :54), the walk over items in increasing index with the two exclusion
rules (:63-78), the size-k min-heap of (score, index) pairs with the strict replacement test
`q.top().first < val` (:80-86) and the fill-from-the-end output order (:88-95).
PINNED: tests/test_oracle.py::test_topk_restatement_is_pinned_to_the_reference checks it against the reference's own
source compiled in place (oracle/ref_topk_shim.cpp -> oracle/_ref/libref_topk.so) and the committed outputs of that
binary (tests/golden/topk.npz)."""
import heapq

import numpy as np

NA_INTEGER = -2147483648


def top_product(x, y, k, nr_ptr=None, nr_idx=None, exclude_1based=(), glob_mean=0.0):
    """x: (n_user, rank), y: (n_item, rank) (rows = embeddings).  Returns (idx 1-based int32 (n_user,k) with
    NA_INTEGER, scores float64 (n_user,k) with NaN)."""
    x64, y64 = np.asarray(x, np.float64), np.asarray(y, np.float64)
    n_user, n_item = x64.shape[0], y64.shape[0]
    exclude_set = set(int(e) for e in exclude_1based)
    use_filter = nr_ptr is not None and nr_ptr[-1] > 0     # :33 not_empty_filter_matrix
    res = np.full((n_user, k), NA_INTEGER, np.int32)
    scores = np.full((n_user, k), np.nan, np.float64)
    for j in range(n_user):
        yvec = y64 @ x64[j]
        cols = nr_idx[nr_ptr[j]:nr_ptr[j + 1]] if use_filter else ()
        u = 0
        q = []
        for i in range(n_item):
            val = float(yvec[i])
            skip = False
            if use_filter and len(cols) > 0 and u < len(cols):
                if i == cols[u]:
                    skip = True
                    u += 1
            if (i + 1) in exclude_set:
                skip = True
            if len(q) < k:
                if not skip:
                    heapq.heappush(q, (val, i))
            elif q[0][0] < val and not skip:
                heapq.heapreplace(q, (val, i))
        q_size = len(q)
        for t in range(q_size):
            v, i = heapq.heappop(q)
            res[j, q_size - t - 1] = i + 1
            scores[j, q_size - t - 1] = v
    if glob_mean != 0.0:
        scores = scores + glob_mean
    return res, scores
