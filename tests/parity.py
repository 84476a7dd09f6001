"""Shared parity checks (test infrastructure)."""

import torch

import oracle

# north_star: "scores must agree within 1e-5 relative"
SCORE_RTOL = 1e-5
# two fp32 summation orders may order two database rows differently when their
# scores differ by less than fp32 accumulation noise; such a disagreement with
# the fp32 oracle is accepted only if the oracle's own scores of the two rows
# are this close, and the fp64 adjudicator sides with us.
FP32_TIE_ATOL = 2e-6


def check_topk_against_oracle(q, db, k, scores, idx, f64_exact=True):
    """q, db CPU fp32; scores/idx = result of the CUDA path (any device)."""
    scores, idx = scores.cpu(), idx.cpu()
    o_s, o_i = oracle.topk_search(q, db, k)
    a_s, a_i = oracle.topk_search_f64(q, db, k)
    # 1. index-exact against the fp64 adjudicator
    if f64_exact:
        bad = (idx != a_i).any(dim=1).nonzero().flatten()
        if bad.numel():
            # an fp64 tie/near-tie (|d| < 1e-13) is the only excuse
            for r in bad.tolist():
                pos = (idx[r] != a_i[r]).nonzero().flatten()
                mine = (q[r].double() * db[idx[r, pos]].double()).sum(1)
                assert torch.allclose(mine, a_s[r, pos], rtol=0, atol=1e-13), \
                    "row %d: indices differ from the fp64 ranking" % r
    # 2. scores: within 1e-5 relative of the reference fp32 path
    assert torch.allclose(scores, o_s, rtol=SCORE_RTOL, atol=1e-7), \
        "max rel score err %g" % ((scores - o_s).abs() / o_s.abs().clamp_min(1e-6)).max()
    assert torch.allclose(scores.double(), a_s, rtol=2e-7, atol=1e-9)
    # 3. against the fp32 oracle: identical except fp32-noise ties
    mism = (idx != o_i)
    n_mism = int(mism.sum())
    if n_mism:
        rows, cols = mism.nonzero(as_tuple=True)
        # oracle's fp32 score of the row we returned at that position
        s_mine = (q[rows] * db[idx[rows, cols]]).sum(1)
        gap = (s_mine - o_s[rows, cols]).abs()
        assert float(gap.max()) <= FP32_TIE_ATOL, \
            "index mismatch not explained by fp32 noise: gap %g" % gap.max()
    return n_mism
