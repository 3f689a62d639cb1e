"""Generates tests/golden/topk.npz -- run HERE (authoring container), where /root/reference exists.

Every expected array is produced by the REFERENCE'S OWN SOURCE: /root/reference/src/matrix_top_product.cpp:20-102
compiled unmodified, in place, against oracle/mini_rcpp + oracle/mini_arma (oracle/build_ref.sh ->
oracle/_ref/libref_topk.so) and driven like R/utils.R:31-59 (find_top_product).  Inputs are rebuilt by the tests from
tests/wrmf_cases.py::topk_cases() (integer hashing, no RNG), only outputs are stored.

    python tests/golden/make_golden_topk.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import wrmf_cases as wc  # noqa: E402

assert oracle.ref_topk_available() or os.path.isdir("/root/reference/src"), "needs the reference tree"
out = {}
for name, c in wc.topk_cases().items():
    nr = c["nr"]
    idx, sc = oracle.ref_top_product(c["x"], c["y"], c["k"], None if nr is None else nr.indptr, None if nr is None else nr.indices,
                                     [e + 1 for e in c["exclude"]], c["glob_mean"])
    out[name + "/idx"] = idx
    out[name + "/scores"] = sc
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "topk.npz"), **out)
print("wrote topk.npz:", sorted(out))
