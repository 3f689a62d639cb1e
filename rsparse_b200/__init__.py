"""rsparse_b200 -- Blackwell-native WRMF / ALS engine behind rsparse's matrix-factorization API.

The compute lives in `libb200als.so` (hand-written sm_100a CUDA behind the C ABI of
include/b200als.h).  This package is the host-side mirror of the reference's R layer
(R/model_WRMF.R) used by the tests and the benchmark in an image without R:
    from rsparse_b200 import WRMF
    model = WRMF(rank=128, lambda_=0.1, feedback="implicit", solver="conjugate_gradient", precision="float")
    user_emb = model.fit_transform(x, n_iter=10)
    model.components          # rank x n_item
"""
from . import _lib  # noqa: F401
from .ops import als_explicit, als_implicit, gram, top_product  # noqa: F401
from .wrmf import WRMF, Session  # noqa: F401
