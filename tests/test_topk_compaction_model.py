"""CPU model of the streaming top-k epilogue's row compaction (csrc/isb_topk.cuh,
warp_compact_row): the bit-by-bit threshold select with its early exit, restated in Python
to check the invariants the screen's completeness certificate rests on -- no GPU needed:

  * at least `keep` entries survive and at most keep + slack (exactly keep when slack = 0);
  * the returned threshold T is a lower bound of the keep-th largest key;
  * every dropped entry is <= every kept one (so "everything outside the candidate list
    scores <= the worst candidate" keeps holding).
"""

import random

import pytest


def compact_row(keys, keep, slack):
    """(T, kept) exactly as warp_compact_row computes them (keys: uint32 ints)."""
    cnt = len(keys)
    k0 = keys[0]
    diff = 0
    for k in keys:
        diff |= k ^ k0
    T, c_T, exact = k0, cnt, True
    if diff:
        hb = diff.bit_length() - 1
        T = 0 if hb == 31 else k0 & ~((2 << hb) - 1) & 0xFFFFFFFF
        for b in range(hb, -1, -1):
            trial = T | (1 << b)
            c = sum(1 for k in keys if k >= trial)
            if c >= keep:
                T, c_T = trial, c
                if c <= keep + slack and b > 0:
                    exact = False
                    break
    if not exact and c_T <= keep + slack:
        return T, [k for k in keys if k >= T]
    greater = [k for k in keys if k > T]
    equal = [k for k in keys if k == T]
    return T, greater + equal[:max(0, keep - len(greater))]


def _row(rng, n, kind):
    if kind == "random":
        return [rng.getrandbits(32) for _ in range(n)]
    if kind == "clustered":      # scores of one row share sign and exponent: common leading bits
        base = rng.getrandbits(32) & 0xFFFF0000
        return [base | rng.getrandbits(16) for _ in range(n)]
    if kind == "ties":
        vals = [rng.getrandbits(32) for _ in range(5)]
        return [rng.choice(vals) for _ in range(n)]
    return [rng.getrandbits(32)] * n          # "constant"


@pytest.mark.parametrize("kind", ["random", "clustered", "ties", "constant"])
@pytest.mark.parametrize("keep,slack", [(128, 64), (128, 0), (38, 64), (1, 0), (16, 64)])
def test_compaction_invariants(kind, keep, slack):
    rng = random.Random(hash((kind, keep, slack)) & 0xFFFF)
    for trial in range(60):
        n = rng.randint(keep, 512)
        keys = _row(rng, n, kind)
        T, kept = compact_row(keys, keep, slack)
        kth = sorted(keys, reverse=True)[keep - 1]
        assert T <= kth                                      # a lower bound of the keep-th best
        assert keep <= len(kept) <= max(keep + slack, keep)  # room for the next tile's appends
        if slack == 0:
            assert len(kept) == keep and sorted(kept, reverse=True) == sorted(keys, reverse=True)[:keep]
        dropped = sorted(keys)
        for k in kept:
            dropped.remove(k)
        assert not dropped or max(dropped) <= min(kept)      # nothing better than a survivor is lost
        assert min(kept) >= T
