"""Index-map emulation of `issue_tile` in als_cg_tile_kernel (rsparse_b200/csrc/als_cg_tile.cuh): a warp instruction
copies CW consecutive 16-byte chunks of RP gathered rows.  Pure numpy restatement of the loop bounds and offsets: every
(gathered row, 16-byte chunk below the rank) must be copied exactly once, padding chunks (4 c >= k) never, for every
padded width the kernel is instantiated with and every CTA size the launch classes use."""
import numpy as np
import pytest


@pytest.mark.parametrize("kpad,k", [(16, 16), (16, 12), (32, 32), (32, 20), (64, 64), (64, 40), (128, 128), (128, 100),
                                    (256, 256), (256, 200)])
@pytest.mark.parametrize("warps", [2, 4, 8, 16])
@pytest.mark.parametrize("n", [0, 1, 7, 33, 100, 208])
def test_every_chunk_of_every_gathered_row_is_copied_once(kpad, k, warps, n):
    cpr = kpad // 4                      # 16-byte chunks per padded row
    cw = min(cpr, 32)                    # chunks of one row per warp instruction
    rp = 32 // cw                        # rows per warp instruction
    nc = cpr // cw                       # instructions per row
    hits = np.zeros((max(n, 1), cpr), dtype=int)
    for w in range(warps):
        for lane in range(32):
            cl = lane % cw
            j = w * rp + lane // cw
            while j < n:
                for cc in range(nc):
                    if 4 * (cl + cc * cw) < k:
                        hits[j, cl + cc * cw] += 1      # dst = tile + j * KPAD + 4 * (cl + cc * CW)
                j += warps * rp
    want = np.zeros_like(hits)
    if n > 0:
        want[:n, :k // 4] = 1
    assert np.array_equal(hits[:max(n, 1)], want)
