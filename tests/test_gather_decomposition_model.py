"""CPU model of the region gather kernel's channel decomposition (csrc/isb_regions.cu,
region_gather_kernel / isb_region_gather): CTA x warp x unit -> channels, restated in Python.
Every channel of an image is produced by exactly one (CTA, warp, unit), a warp's units start
at increasing channels (so it may stop at the first empty one), and exactly one warp writes the
zero padding behind the last channel -- no GPU needed."""

import pytest

WARPS = 8


def launch_shape(C, HW, B, G_default=4):
    CW = 1
    for cw in (8, 4, 2, 1):
        if cw * HW * 4 <= 4 * 1024:
            CW = cw
            break
    units = (C + WARPS * CW - 1) // (WARPS * CW)
    G = G_default
    while G > 1 and B * ((units + G - 1) // G) < 148 * 8:
        G >>= 1
    G = min(G, units)
    return CW, G, (units + G - 1) // G


@pytest.mark.parametrize("C", [2048, 512, 96, 100, 32, 30, 7])
@pytest.mark.parametrize("HW,B", [(196, 256), (1024, 256), (49, 3), (108, 2), (90, 1)])
def test_every_channel_once(C, HW, B):
    CW, G, grid_x = launch_shape(C, HW, B)
    owner = {}
    padding_writers = 0
    for cta in range(grid_x):
        cta_c0 = cta * (WARPS * CW * G)
        for warp in range(WARPS):
            prev = -1
            for u in range(G):
                c0 = cta_c0 + (u * WARPS + warp) * CW
                assert c0 > prev                       # increasing: break at the first empty unit is safe
                prev = c0
                nch = max(0, min(CW, C - c0))
                if nch <= 0:
                    break
                for c in range(c0, c0 + nch):
                    assert c not in owner
                    owner[c] = (cta, warp, u)
                if c0 + nch >= C:
                    padding_writers += 1
    assert sorted(owner) == list(range(C))
    assert padding_writers == 1
