"""Shared parity checks (test infrastructure)."""

import json
import os

import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# north_star: "scores must agree within 1e-5 relative"
SCORE_RTOL = 1e-5
# two fp32 summation orders may order two database rows differently when their
# scores differ by less than fp32 accumulation noise; such a disagreement with
# the fp32 oracle is accepted only if the oracle's own scores of the two rows
# are this close, and the fp64 adjudicator sides with us.
# fp32 dot noise scales with the score: the tolerated gap is 16 ulp of the score
# (FP32_TIE_RTOL = 16 * 2^-24) + 5e-8.  Measured on the B200 box: largest gap 1.04e-7 on random
# rows (|score| ~ 0.3, 4 ulp) and 4.8e-7 on near-duplicate clusters (|score| ~ 1, 8 ulp),
# profiles/r02_parity_achieved.jsonl.
FP32_TIE_RTOL = 16 * 2.0 ** -24
FP32_TIE_ATOL = 5e-8
# ... and a disagreement may only sit where the oracle's OWN ranking is undecided at that
# tolerance: the oracle's score at that position is within the tie tolerance of one of its
# neighbours in the oracle's sorted list (the oracle is asked for k+1 so the last position has a
# right neighbour).  On data without near-ties that set is (almost) empty, on near-duplicate
# clusters it is large.  On data WITHOUT planted near-duplicates the count is bounded too:
# at most this fraction of the Q*k entries (+2); callers with clustered data pass None.
FP32_TIE_MAX_FRACTION = 2e-3

# region descriptors (unit-norm rows, split-operand fp32-grade projection): per-row L2
# error against the oracle's descriptor, and 1 - cosine in fp64.  The floor is the 16-bit
# (hi + lo bf16) representation of the two operands: 4.3e-6 relative on the projection at
# K = 100352 even with exact accumulation (tools/gemm_precision.py), 3e-6 .. 6.3e-6 measured
# over the suite (profiles/r02_parity_achieved.jsonl); torch's own fp32 matmul on the GPU is
# 5.9e-6 from the fp64 product on the same operands.  Tolerance = 1.6 x the largest measured.
DESC_L2_TOL = 1e-5
DESC_COS_TOL = 1e-10


def record(kind, **values):
    """Append the achieved error of a parity check to gpurun_out/parity_achieved.jsonl
    (only when that directory exists: the GPU box visit) so tolerances can be pinned to
    what is measured."""
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    with open(os.path.join(d, "parity_achieved.jsonl"), "a") as f:
        f.write(json.dumps(dict(kind=kind, test=test, **values)) + "\n")


def check_descriptors(got, want, l2_tol=DESC_L2_TOL, cos_tol=DESC_COS_TOL, unit=True):
    """got/want [B, D] descriptors (model/siamese.py:222: unit-norm rows).  Asserts the
    per-row L2 error and 1 - cosine (fp64) and records what was achieved."""
    g, w = got.detach().cpu().double(), want.detach().cpu().double()
    assert g.shape == w.shape
    if g.numel() == 0:
        return 0.0
    l2 = (g - w).norm(dim=1)
    cos = (g * w).sum(1) / (g.norm(dim=1) * w.norm(dim=1)).clamp_min(1e-300)
    record("descriptor", rows=g.size(0), dim=g.size(1), max_l2=float(l2.max()), max_1mcos=float((1 - cos).max()),
           max_abs=float((g - w).abs().max()))
    assert float(l2.max()) <= l2_tol, "descriptor L2 error %g > %g" % (float(l2.max()), l2_tol)
    assert float((1 - cos).max()) <= cos_tol, "1 - cosine %g > %g" % (float((1 - cos).max()), cos_tol)
    if unit:
        assert float((g.norm(dim=1) - 1).abs().max()) <= 1e-6
    return float(l2.max())


def check_topk_against_oracle(q, db, k, scores, idx, f64_exact=True, oracle_result=None,
                              max_mismatch_fraction=FP32_TIE_MAX_FRACTION):
    """q, db CPU fp32; scores/idx = result of the CUDA path (any device).  Returns the
    number of entries that differ from the fp32 oracle (all of them fp32-noise ties that
    the fp64 ranking decides our way, each at a position the oracle's own ranking leaves undecided)."""
    scores, idx = scores.cpu(), idx.cpu()
    if oracle_result is None and k < db.size(0):
        o_s1, o_i1 = oracle.topk_search(q, db, k + 1)
        o_s, o_i, o_next = o_s1[:, :k], o_i1[:, :k], o_s1[:, k:]
    else:
        o_s, o_i = oracle.topk_search(q, db, k) if oracle_result is None else oracle_result
        o_next = o_s[:, -1:]                       # unknown (k+1)-th: last position counts as undecided
    # 1. index-exact against the fp64 adjudicator
    if f64_exact:
        a_s, a_i = oracle.topk_search_f64(q, db, k)
        bad = (idx != a_i).any(dim=1).nonzero().flatten()
        if bad.numel():
            # an fp64 tie/near-tie (|d| < 1e-13) is the only excuse
            for r in bad.tolist():
                pos = (idx[r] != a_i[r]).nonzero().flatten()
                mine = (q[r].double() * db[idx[r, pos]].double()).sum(1)
                assert torch.allclose(mine, a_s[r, pos], rtol=0, atol=1e-13), \
                    "row %d: indices differ from the fp64 ranking" % r
        assert torch.allclose(scores.double(), a_s, rtol=2e-7, atol=1e-9)
    # 2. scores: within 1e-5 relative of the reference fp32 path
    rel = ((scores - o_s).abs() / o_s.abs().clamp_min(1e-6)).max() if scores.numel() else 0.0
    assert torch.allclose(scores, o_s, rtol=SCORE_RTOL, atol=1e-7), "max rel score err %g" % rel
    # 3. against the fp32 oracle: identical except fp32-noise ties
    mism = (idx != o_i)
    n_mism = int(mism.sum())
    max_gap = 0.0
    if n_mism:
        rows, cols = mism.nonzero(as_tuple=True)
        # oracle's fp32 score of the row we returned at that position
        s_mine = (q[rows] * db[idx[rows, cols]]).sum(1)
        gap = (s_mine - o_s[rows, cols]).abs()
        max_gap = float(gap.max())
        tol = FP32_TIE_ATOL + FP32_TIE_RTOL * o_s[rows, cols].abs()
        assert bool((gap <= tol).all()), "index mismatch not explained by fp32 noise: gap %g" % max_gap
        ext = torch.cat([o_s[:, :1] + 1.0, o_s, o_next], dim=1)
        near = torch.minimum(ext[:, :-2] - o_s, o_s - ext[:, 2:])          # gap to the closer neighbour
        undecided = near <= FP32_TIE_ATOL + FP32_TIE_RTOL * o_s.abs()
        assert bool(undecided[rows, cols].all()), \
            "%d entries differ from the fp32 oracle where its own ranking is decided" % \
            int((~undecided[rows, cols]).sum())
        if max_mismatch_fraction is not None:
            assert n_mism <= max_mismatch_fraction * idx.numel() + 2, \
                "%d of %d entries differ from the fp32 oracle" % (n_mism, idx.numel())
    record("topk", Q=q.size(0), N=db.size(0), D=q.size(1), k=k, mismatches_vs_f32=n_mism, max_tie_gap=max_gap,
           max_rel_score_err=float(rel))
    return n_mism
