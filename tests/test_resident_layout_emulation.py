"""Index-map emulation of als_cg_resident_kernel's register layout (rsparse_b200/csrc/als_resident.cuh): tile slots
-> half-warps -> register groups -> transposing halving reduction -> owner lanes -> w_j broadcast -> exchange
between the half-warps.  Pure numpy restatement of the index arithmetic (not of the numerics): it checks that every
owner lane ends up with the dot product of ITS slot and that lane L ends up with chunk L of sum_j w_j x_j, for full
and ragged rows.  The numerics are covered by the -m gpu parity tests."""
import numpy as np
import pytest

K, IPW, GROUPS, WARPS = 128, 20, 10, 4


def owner_group(lane):
    b3, b2, b1, b0 = (lane >> 3) & 1, (lane >> 2) & 1, (lane >> 1) & 1, lane & 1
    in3 = 2 * b1 + b0
    in5 = 3 * b2 + in3
    if in3 >= 3 or in5 >= 5:
        return -1, in5
    return 5 * b3 + in5, in5


def group_owner_lane(qn):
    b3, in5 = qn // 5, qn % 5
    b2, in3 = in5 // 3, in5 % 3
    return 8 * b3 + 4 * b2 + 2 * (in3 // 2) + (in3 % 2)


def halver(t, lane_bit, swap_free_levels):
    """t: [32 lanes][N]; returns [32] after the recursion of Halver<N>::run<M, levels>."""
    n = t.shape[1]
    h = (n + 1) // 2
    lanes = np.arange(32)
    o = np.zeros((32, h))
    if swap_free_levels > 0 and n % 2 == 0:
        for v in range(h):
            o[:, v] = t[:, v] + t[lanes ^ lane_bit, v + h]
    else:
        upper = (lanes & lane_bit) != 0
        for v in range(h):
            lo = t[:, v]
            hi = t[:, v + h] if v + h < n else np.zeros(32)
            send = np.where(upper, lo, hi)
            keep = np.where(upper, hi, lo)
            o[:, v] = keep + send[lanes ^ lane_bit]
    if lane_bit == 1:
        return o[:, 0]
    return halver(o, lane_bit // 2, max(swap_free_levels - 1, 0))


@pytest.mark.parametrize("n", [80, 79, 41, 7, 3, 1])
@pytest.mark.parametrize("w", [0, 3])
def test_half_warp_layout(n, w):
    rng = np.random.default_rng(n * 7 + w)
    nw = (n - w + WARPS - 1) // WARPS if n > w else 0
    tile = rng.standard_normal((IPW, K))          # tile slot s = gathered row w + 4 s (garbage beyond nw)
    vec = rng.standard_normal(K)
    lanes = np.arange(32)
    # ---- tile load: xt[lane][2q + c] = 4 floats
    xt = np.zeros((32, IPW, 4))
    seen = np.zeros((IPW, 32), dtype=int)
    for lane in range(32):
        g, l, ob = lane >> 4, lane & 15, ((lane >> 3) & 1) * 5
        nb = [ob, 5 - ob]
        nwh = (nw - g + 1) >> 1
        for q in range(GROUPS):
            ok = (q % 5) < nwh - nb[q // 5]
            for c in range(2):
                slot = 2 * (nb[q // 5] + q % 5) + g
                chunk = (c ^ g) * 16 + l
                assert ok == (slot < nw)
                if ok:
                    xt[lane, 2 * q + c] = tile[slot, 4 * chunk:4 * chunk + 4]
                    seen[slot, chunk] += 1
    assert (seen[:nw] == 1).all() and (seen[nw:] == 0).all()
    # ---- dots
    vchunk = vec.reshape(32, 4)
    t = np.zeros((32, GROUPS))
    for lane in range(32):
        for q in range(GROUPS):
            t[lane, q] = xt[lane, 2 * q] @ vchunk[lane] + xt[lane, 2 * q + 1] @ vchunk[lane ^ 16]
    u = halver(t, 8, 1)
    want = tile @ vec
    wq = np.zeros(32)
    wbuf = np.full(32, np.nan)
    weights = rng.standard_normal(IPW)
    for lane in range(32):
        og, in5 = owner_group(lane)
        if og < 0:
            continue
        slot = 2 * og + (lane >> 4)
        assert group_owner_lane(og) | (lane & 16) == lane
        if slot < nw:
            np.testing.assert_allclose(u[lane], want[slot], rtol=1e-12, atol=1e-12)
        else:
            assert u[lane] == 0.0
        wq[lane] = weights[slot] if slot < nw else 0.0
        wbuf[(lane & 24) + in5] = wq[lane]
    # ---- apply
    acc0 = np.zeros((32, 4))
    acc1 = np.zeros((32, 4))
    for lane in range(32):
        for h in range(2):
            base = (lane & 24) ^ (h * 8)
            for r in range(5):
                q = 5 * h + r
                acc0[lane] += wbuf[base + r] * xt[lane, 2 * q]
                acc1[lane] += wbuf[base + r] * xt[lane, 2 * q + 1]
    acc = acc0 + acc1[lanes ^ 16]
    want_acc = (weights[:nw, None] * tile[:nw]).sum(axis=0).reshape(32, 4)
    np.testing.assert_allclose(acc, want_acc, rtol=1e-12, atol=1e-12)


def test_copy_map_covers_each_slot_once():
    """cp.async flavour of issue_tile: instruction (qn, c) of lane (g, l) writes floats [64 c + 4 l, +4) of slot 2 qn + g."""
    cover = np.zeros((IPW, K), dtype=int)
    for lane in range(32):
        g, l = lane >> 4, lane & 15
        for qn in range(GROUPS):
            for c in range(2):
                cover[2 * qn + g, 64 * c + 4 * l:64 * c + 4 * l + 4] += 1
    assert (cover == 1).all()
