"""CPU model of the reduced candidate lists of the sharded search (search.ShardedIndex.list_length,
csrc/isb_search.cu global_threshold_kernel / rerank_owned_kernel): every shard lists only kl < kc of
its best screen scores, the threshold is the kc-th best of the union, and a FULL list that lies
entirely above the threshold is reported as cut.  Property checked on random and adversarial
placements: whenever no shard reports a cut, every database row whose screen score exceeds the
threshold has been listed (so the global completeness certificate of the merge covers the rows that
were not re-ranked) -- and the placements that break the assumption are reported.  No GPU needed."""

import math

import pytest
import torch

NEG = float("-inf")


def list_length(kc, R):
    mean = kc / float(R)
    return min(kc, int(mean + 6.0 * mean ** 0.5 + 0.999) + 8)


def exchange(screen_shards, kl, kth):
    """screen_shards: list of [Q, n_r] screen scores.  Returns (thr [Q], listed masks, cut [Q] bool)."""
    Q = screen_shards[0].size(0)
    lists = []
    for s in screen_shards:
        n = min(kl, s.size(1))
        v, i = s.topk(n, dim=1)
        full = torch.full((Q, kl), NEG)
        full[:, :n] = v
        lists.append((full, i))
    union = torch.cat([l[0] for l in lists], 1)
    srt = union.sort(dim=1, descending=True).values
    enough = (union > NEG).sum(1) >= kth
    thr = torch.where(enough, srt[:, min(kth, srt.size(1)) - 1], torch.full((Q,), NEG))
    if kth > kl:      # too few entries overall while some list is full: nothing can be concluded
        any_full = torch.stack([l[0][:, kl - 1] > NEG for l in lists]).any(0)
        thr = torch.where(~enough & any_full, torch.full((Q,), float("inf")), thr)
    cut = torch.zeros(Q, dtype=torch.bool)
    for full, _ in lists:
        cut |= (full > thr[:, None]).sum(1) == kl
    return thr, lists, cut


def rows_above_threshold_are_listed(screen_shards, lists, thr, q):
    for s, (full, idx) in zip(screen_shards, lists):
        above = (s[q] > thr[q]).nonzero().flatten().tolist()
        listed = set(idx[q].tolist())
        if any(a not in listed for a in above):
            return False
    return True


@pytest.mark.parametrize("R,n_total,k", [(8, 40000, 100), (4, 9000, 100), (2, 3000, 100), (3, 1000, 60), (8, 520, 100)])
def test_no_cut_reported_means_every_row_above_the_threshold_is_listed(R, n_total, k):
    g = torch.Generator().manual_seed(R * 1000 + k)
    kc = min(k + 28, 128)
    kl = list_length(kc, R)
    Q = 64
    bounds = [(r * n_total // R, (r + 1) * n_total // R) for r in range(R)]
    screen = torch.randn(Q, n_total, generator=g)
    shards = [screen[:, lo:hi] for lo, hi in bounds]
    thr, lists, cut = exchange(shards, kl, kc)
    assert not bool(cut.any())                       # random placement: the reduced lists suffice
    for q in range(Q):
        assert rows_above_threshold_are_listed(shards, lists, thr, q)
    # the threshold is the one full-length lists give
    thr_full, _, cut_full = exchange(shards, kc, kc)
    assert torch.equal(thr, thr_full) and not bool(cut_full.any())


def test_adversarial_placement_is_reported_not_missed():
    # every query's best rows sit on shard 0 (a database sorted by instance)
    g = torch.Generator().manual_seed(3)
    R, n_total, k, Q = 8, 16000, 100, 40
    kc, kl = 128, list_length(128, 8)
    screen = torch.randn(Q, n_total, generator=g)
    screen[:, :300] += 6.0                           # 300 rows of shard 0 dominate every query
    shards = [screen[:, r * 2000:(r + 1) * 2000] for r in range(R)]
    thr, lists, cut = exchange(shards, kl, kc)
    for q in range(Q):
        ok = rows_above_threshold_are_listed(shards, lists, thr, q)
        assert ok or bool(cut[q])                    # a miss is never silent
    assert bool(cut.all())                           # ... and here every query is one
    # full-length lists: no cut can be reported, nothing is missed
    thr2, lists2, cut2 = exchange(shards, kc, kc)
    assert not bool(cut2.any())
    assert all(rows_above_threshold_are_listed(shards, lists2, thr2, q) for q in range(Q))


def test_short_union_with_a_full_list_cannot_be_concluded():
    # one big shard and tiny ones: fewer than kth entries overall although the big shard's list is full
    g = torch.Generator().manual_seed(9)
    kc, kl = 128, 48
    shards = [torch.randn(5, 5000, generator=g)] + [torch.randn(5, 4, generator=g) for _ in range(7)]
    thr, lists, cut = exchange(shards, kl, kc)
    assert bool((thr == float("inf")).all())          # -> nothing re-ranked, the merge cannot certify
    # with full-length lists the union holds >= kc entries and the threshold is finite
    thr2, _, cut2 = exchange(shards, kc, kc)
    assert bool(torch.isfinite(thr2).all()) and not bool(cut2.any())


def test_list_length_covers_the_binomial_tail():
    # the number of global top-kc rows on one of R equally likely shards is Binomial(kc, 1/R): the list
    # length leaves a tail below 1e-9 per (query, shard)
    for R in (2, 3, 4, 8):
        kc = 128
        kl = list_length(kc, R)
        if kl >= kc:
            continue
        p = 1.0 / R
        tail = sum(math.comb(kc, j) * p ** j * (1 - p) ** (kc - j) for j in range(kl, kc + 1))
        assert tail < 1e-9, (R, kl, tail)
