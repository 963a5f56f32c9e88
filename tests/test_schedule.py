"""Work partition of the GEMM-regime kernels, restated in Python (the arithmetic of dense_unit_kernel in
csrc/smx_dense_kernel.cu): every half block of every output block is owned by exactly one warp of exactly one CTA group,
no warp owns more than 4 units or touches more than 3 output blocks, and the load per SM sub-partition (warp % 4) is within
one unit of even."""
import pytest


def unit_partition(nblk, nw=8):
    groups = -(-nblk // (2 * nw))
    gsz = -(-nblk // groups)
    owned = {}
    for g in range(groups):
        g0 = g * gsz
        gcnt = min(gsz, nblk - g0)
        H = 2 * gcnt
        loads = [0] * 4
        for warp in range(nw):
            t0 = warp * (H // nw) + min(warp, H % nw)
            nu = H // nw + (1 if warp < H % nw else 0)
            assert nu <= 4
            blocks = {g0 + ((t0 + u) >> 1) for u in range(nu)}
            assert len(blocks) <= 3 and all(b < nblk for b in blocks)
            jb0 = g0 + (t0 >> 1)
            assert all(0 <= b - jb0 <= 2 for b in blocks)
            for u in range(nu):
                key = (g0 + ((t0 + u) >> 1), (t0 + u) & 1)
                assert key not in owned
                owned[key] = (g, warp)
            loads[warp % 4] += nu
        assert max(loads) - min(loads) <= 2 and max(loads) <= -(-H // 4) + 1
    return owned


@pytest.mark.parametrize("nblk", range(1, 48))
def test_half_block_partition_covers_every_unit_once(nblk):
    owned = unit_partition(nblk)
    assert set(owned) == {(b, h) for b in range(nblk) for h in (0, 1)}


def test_thirteen_blocks_are_balanced_over_the_sub_partitions():
    # the cfg5 shape (d_out = 100): 26 units on 8 warps = 7, 7, 6, 6 per sub-partition (whole blocks: 4 of 13 on one)
    H, nw = 26, 8
    per_warp = [H // nw + (1 if w < H % nw else 0) for w in range(nw)]
    per_sp = [per_warp[w] + per_warp[w + 4] for w in range(4)]
    assert per_warp == [4, 4, 3, 3, 3, 3, 3, 3] and sorted(per_sp) == [6, 6, 7, 7]
