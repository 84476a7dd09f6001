"""CPU model of the screen's work decomposition (csrc/isb_topk.cuh TopkSched, csrc/isb_search.cu
pick_n_groups / make_search_plan), restated in Python: every (row block, n-tile) pair is covered
exactly once, in single-CTA and in CTA-pair mode, and the wave barrier's expected arrival
counts add up to the number of segments -- no GPU needed."""

import pytest

BM, BN, SMS = 128, 256, 148


def pick_n_groups(m_blocks, n_tiles, grid):
    hi = max(1, min(n_tiles, ((grid + m_blocks - 1) // m_blocks) * 8 + 8))

    def eff(ng):
        segs = m_blocks * ng
        waves = (segs + grid - 1) // grid
        longest = (n_tiles + ng - 1) // ng + 3.0        # kSegOverheadTiles
        return (m_blocks * n_tiles / grid) / (waves * longest)
    best = max(eff(ng) for ng in range(1, hi + 1))
    for ng in range(1, hi + 1):
        if eff(ng) >= best - 0.01:
            return ng
    return 1


def plan(Q, N):
    m_blocks = (Q + BM - 1) // BM
    n_tiles = (N + BN - 1) // BN
    grid = min(m_blocks * n_tiles, SMS)
    pair_m = (m_blocks + 1) // 2
    pair = m_blocks >= 2 and pair_m * n_tiles >= SMS // 2
    if pair:
        grid = (SMS // 2) * 2
        return dict(pair=True, sched_m=pair_m, workers=grid // 2, n_tiles=n_tiles,
                    n_groups=pick_n_groups(pair_m, n_tiles, grid // 2), m_blocks=m_blocks)
    return dict(pair=False, sched_m=m_blocks, workers=grid, n_tiles=n_tiles,
                n_groups=pick_n_groups(m_blocks, n_tiles, grid), m_blocks=m_blocks)


def segment(p, s):
    g, m = divmod(s, p["sched_m"])
    return m, g * p["n_tiles"] // p["n_groups"], (g + 1) * p["n_tiles"] // p["n_groups"]


@pytest.mark.parametrize("Q,N", [(10000, 1000000), (10000, 125000), (10000, 250000), (641, 40001), (385, 25000),
                                 (300, 20000), (1000, 5000), (64, 70000), (1, 1), (129, 19000), (16384, 16384)])
def test_segments_cover_every_tile_once(Q, N):
    p = plan(Q, N)
    assert 1 <= p["n_groups"] <= p["n_tiles"]
    seen = {}
    n_seg = p["sched_m"] * p["n_groups"]
    for s in range(n_seg):
        m, t0, t1 = segment(p, s)
        assert t0 < t1                      # no empty segment (n_groups <= n_tiles)
        for rank in ((0, 1) if p["pair"] else (0,)):
            mb = 2 * m + rank if p["pair"] else m
            for t in range(t0, t1):
                seen[(mb, t)] = seen.get((mb, t), 0) + 1
    real = {(mb, t) for mb in range(p["m_blocks"]) for t in range(p["n_tiles"])}
    assert real <= set(seen) and all(v == 1 for v in seen.values())
    # an odd number of row blocks leaves the last pair's second CTA on rows beyond Q (masked out)
    assert len(seen) - len(real) in (0, p["n_tiles"])
    # wave barrier: worker w takes segments w, w + workers, ...; wave k expects
    # min(workers, n_seg - k * workers) arrivals -- exactly the workers that have a k-th segment
    waves = (n_seg + p["workers"] - 1) // p["workers"]
    for k in range(waves):
        arrivals = sum(1 for w in range(p["workers"]) if w + k * p["workers"] < n_seg)
        assert arrivals == min(p["workers"], n_seg - k * p["workers"])
    assert sum(min(p["workers"], n_seg - k * p["workers"]) for k in range(waves)) == n_seg


def test_headline_plans():
    one = plan(10000, 1000000)
    assert one["pair"] and one["sched_m"] == 40 and one["workers"] == 74 and one["n_groups"] == 11
    shard = plan(10000, 125000)              # the per-GPU shard of the 8-GPU run
    assert shard["pair"] and shard["n_groups"] == 9      # 5 waves of 55-tile segments (measured: 6.08 ms vs 6.12 with 11)
    assert not plan(64, 70000)["pair"]       # a single row block: the single-CTA kernel
    mining = plan(16384, 16384)              # BASELINE configs[2]: one 64-tile segment per row-block pair
    assert mining["pair"] and mining["n_groups"] == 1
