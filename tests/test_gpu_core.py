"""GPU parity tests of the core kernels through the C ABI (ctypes -> libisb.so):
row operators, bf16 conversion, the tcgen05 GEMM and the top-k search."""

import pytest
import torch

import oracle
from conftest import load_golden
from parity import check_topk_against_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from instance_search_b200 import ops as o
    return o


def _randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


# ------------------------------------------------------------------ a1 / a2
@pytest.mark.parametrize("shape", [(1, 1), (3, 7), (5, 100), (4, 2048), (6, 100352), (2, 100353),
                                   (300, 512), (0, 16)])
def test_l2norm_rows(ops, shape):
    x = _randn(*shape, seed=1) if shape[0] else torch.zeros(shape)
    want = oracle.normalize_l2(x) if shape[0] else x
    got = ops.l2norm_rows(x.cuda()).cpu()
    # fp32, different summation order than torch's: 1e-6 relative
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-9)


def test_l2norm_eps_inside_sqrt(ops):
    x = torch.zeros(2, 64)
    x[1] = 1e-6
    got = ops.l2norm_rows(x.cuda()).cpu()
    assert torch.equal(got[0], torch.zeros(64))
    assert torch.allclose(got, oracle.normalize_l2(x), rtol=1e-6, atol=0)


def test_shift_rows(ops):
    x, s = _randn(7, 333, seed=2), _randn(333, seed=3)
    assert torch.equal(ops.shift_rows(x.cuda(), s.cuda()).cpu(), oracle.shift(x, s))


# ------------------------------------------------------------------ bf16 terms
def test_to_bf16_terms(ops):
    x = _randn(37, 100, seed=4) * 3
    hi = ops.to_bf16(x.cuda(), 0).cpu()
    assert hi.shape == (37, 104) and hi.dtype == torch.bfloat16
    assert torch.equal(hi[:, :100], x.to(torch.bfloat16))
    assert torch.equal(hi[:, 100:].float(), torch.zeros(37, 4))
    lo = ops.to_bf16(x.cuda(), 1).cpu()
    r1 = x - hi[:, :100].float()
    assert torch.equal(lo[:, :100], r1.to(torch.bfloat16))
    lo2 = ops.to_bf16(x.cuda(), 2).cpu()
    r2 = r1 - lo[:, :100].float()
    assert torch.equal(lo2[:, :100], r2.to(torch.bfloat16))
    # three terms reproduce fp32 to ~2^-24
    rec = hi.float() + lo.float() + lo2.float()
    assert torch.allclose(rec[:, :100], x, rtol=2e-7, atol=1e-30)


# ------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,N,K,splits", [
    (128, 256, 64, 1),      # one tile, one k-block
    (128, 256, 256, 1),     # ring wraps once
    (128, 256, 1024, 1),    # ring wraps many times
    (256, 512, 128, 1),     # 2x2 tiles
    (200, 300, 136, 1),     # ragged M, N, K (TMA zero fill)
    (1, 1, 8, 1),
    (1000, 2000, 2048, 1),
    (1536, 512, 6272, 4),   # split-K
    (130, 258, 6400, 7),    # ragged + split-K
    (4096, 4096, 512, 1),   # many tiles per CTA: TMEM double buffering
])
def test_gemm_nt(ops, M, N, K, splits):
    a = _randn(M, K, seed=5).cuda()
    b = _randn(N, K, seed=6).cuda()
    bias = _randn(N, seed=7).cuda()
    a16, b16 = ops.to_bf16(a), ops.to_bf16(b)
    got = ops.gemm_nt(a16, b16, bias=bias, splits=splits, k=K)
    # checker: the same bf16 values multiplied in fp64 on the GPU
    want = (a16[:, :K].double() @ b16[:, :K].double().t() + bias.double())
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item()
    # fp32 accumulation of K exact products
    assert err <= 2e-6 * scale * max(1.0, (K / 64) ** 0.5), (err, scale)


def test_gemm_nt_fp32_grade_by_k_concat(ops):
    # [A_hi|A_lo|A_hi] . [B_hi|B_hi|B_lo]^T ~ fp32-grade product (include/isb.h)
    M, N, K = 256, 512, 1024
    a, b = _randn(M, K, seed=8).cuda(), _randn(N, K, seed=9).cuda()
    a_hi, a_lo = ops.to_bf16(a, 0), ops.to_bf16(a, 1)
    b_hi, b_lo = ops.to_bf16(b, 0), ops.to_bf16(b, 1)
    A = torch.cat([a_hi, a_lo, a_hi], 1).contiguous()
    B = torch.cat([b_hi, b_hi, b_lo], 1).contiguous()
    got = ops.gemm_nt(A, B)
    want = a.double() @ b.double().t()
    rel = (got.double() - want).abs().max().item() / want.abs().max().item()
    assert rel < 2e-5, rel
    one = ops.gemm_nt(a_hi, b_hi)
    rel1 = (one.double() - want).abs().max().item() / want.abs().max().item()
    assert rel1 > 20 * rel  # the split really buys precision


@pytest.mark.parametrize("M,N,K,splits", [(256, 512, 1024, 1), (130, 300, 200, 1), (64, 256, 6400, 5)])
def test_gemm_nt_split_operands(ops, M, N, K, splits):
    # a_hi.b_hi + a_lo.b_hi + a_hi.b_lo in one TMEM accumulator == the K-concatenated form
    a, b = _randn(M, K, seed=18).cuda(), _randn(N, K, seed=19).cuda()
    bias = _randn(N, seed=20).cuda()
    a_hi, a_lo = ops.to_bf16(a, 0), ops.to_bf16(a, 1)
    b_hi, b_lo = ops.to_bf16(b, 0), ops.to_bf16(b, 1)
    got = ops.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, bias=bias, splits=splits, k=K)
    want = a.double() @ b.double().t() + bias.double()
    rel = (got.double() - want).abs().max().item() / want.abs().max().item()
    assert rel < 2e-5, rel
    if splits == 1:
        cat = ops.gemm_nt(torch.cat([a_hi, a_lo, a_hi], 1).contiguous(),
                          torch.cat([b_hi, b_hi, b_lo], 1).contiguous(), bias=bias)
        # same products, different order of the k-blocks (the split kernel interleaves the three
        # terms of every k-range): equal up to fp32 accumulation order
        # (the tensor core accumulates 192 k-steps in fp32: ~1e-4 absolute on sums of ~100)
        assert (got - cat).abs().max().item() <= 1e-5 * cat.abs().max().item()


# ------------------------------------------------------------------ top-k search
def _search(ops, q, db, k, margin=None, stats=None, exact=True):
    qd, dbd = q.cuda(), db.cuda()
    s, i = ops.topk_search(qd, dbd, ops.to_bf16(dbd), k, margin, stats=stats, exact=exact)
    torch.cuda.synchronize()
    return s, i


def test_topk_search_golden(ops):
    g = load_golden("search_tiny")
    s, i = _search(ops, g["q"], g["db"], 10)
    assert torch.equal(i.cpu(), g["idx"])
    assert torch.allclose(s.cpu(), g["scores"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("Q,N,D,k", [
    (1, 1, 8, 1),            # degenerate
    (3, 5, 8, 5),            # k == N
    (40, 1500, 64, 10),
    (130, 777, 72, 100),     # ragged everything, D not a multiple of 64
    (300, 20000, 256, 100),
    (64, 70000, 128, 100),   # many n-tiles per segment
    (1000, 5000, 64, 20),    # several m-blocks
    (641, 40001, 200, 100),  # CTA-pair kernel: 6 row blocks (3 pairs, the last one ragged), ragged N and D
    (385, 25000, 96, 10),    # CTA-pair kernel: odd number of row blocks (the last pair's second CTA is idle)
])
def test_topk_search_random(ops, Q, N, D, k):
    q = oracle.normalize_l2(_randn(Q, D, seed=10))
    db = oracle.normalize_l2(_randn(N, D, seed=11))
    s, i = _search(ops, q, db, k)
    check_topk_against_oracle(q, db, k, s, i)
    assert (i >= 0).all() and (i < N).all()
    # size-independent properties: sorted, no duplicates
    assert (s[:, :-1] >= s[:, 1:]).all()
    assert all(len(set(r)) == k for r in i.cpu().tolist()[:50])


def test_topk_search_pair_and_single_cta_kernels_agree(ops):
    # the 256 x 256 CTA-pair screen and the single-CTA 128 x 256 screen feed the same exact
    # re-rank: identical indices and scores (option screen_pair=0 selects the single-CTA kernel)
    from instance_search_b200 import _lib
    Q, N, D, k = 700, 30000, 128, 100
    q = oracle.normalize_l2(_randn(Q, D, seed=31))
    db = oracle.normalize_l2(_randn(N, D, seed=32))
    s2, i2 = _search(ops, q, db, k)
    with _lib.options(screen_pair=0):
        s1, i1 = _search(ops, q, db, k)
    assert _lib.get_option("screen_pair") is None      # restored
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    check_topk_against_oracle(q, db, k, s2, i2)


def test_topk_search_with_seeded_thresholds(ops):
    # option screen_seed=1: every row's threshold starts at the kc-th best score over the first 2048
    # database rows (a lower bound of the kc-th best overall) instead of -inf: same result
    Q, N, D, k = 130, 33000, 72, 100
    q = oracle.normalize_l2(_randn(Q, D, seed=41))
    db = oracle.normalize_l2(_randn(N, D, seed=42))
    from instance_search_b200 import _lib
    s0, i0 = _search(ops, q, db, k)
    with _lib.options(screen_seed=1):
        s1, i1 = _search(ops, q, db, k)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    check_topk_against_oracle(q, db, k, s1, i1)


def test_topk_search_clustered_and_planted(ops):
    # database with near-duplicate clusters + planted exact matches of each query
    Q, N, D, k = 200, 36000, 128, 50
    centers = _randn(300, D, seed=12)
    lab = torch.randint(0, 300, (N,), generator=torch.Generator().manual_seed(13))
    db = oracle.normalize_l2(centers[lab] + 0.05 * _randn(N, D, seed=14))
    q = oracle.normalize_l2(centers[:Q] + 0.05 * _randn(Q, D, seed=15))
    planted = torch.arange(Q) * 7 + 3
    db[planted] = q
    stats = {}
    s, i = _search(ops, q, db, k, stats=stats)
    assert torch.equal(i[:, 0].cpu(), planted)          # round trip: a row finds itself
    assert torch.allclose(s[:, 0].cpu(), torch.ones(Q), atol=1e-6)
    # near-duplicate clusters: hundreds of fp32-undecided neighbours by construction, so only the
    # per-position rule applies (each disagreement sits where the oracle's own ranking is undecided)
    check_topk_against_oracle(q, db, k, s, i, max_mismatch_fraction=None)
    # ~100 near-duplicates per cluster, score spacing far below bf16 noise: the
    # certificate must have sent (most of) these rows to the fp32-grade re-screen
    assert stats["rows"] == Q and stats["resolved_fp32_grade"] > Q // 2
    assert stats["resolved_exhaustive"] == 0


def test_topk_search_certificate_is_silent_on_ordinary_data(ops):
    Q, N, D, k = 256, 50000, 256, 100
    q = oracle.normalize_l2(_randn(Q, D, seed=21))
    db = oracle.normalize_l2(_randn(N, D, seed=22))
    stats = {}
    s, i = _search(ops, q, db, k, stats=stats)
    check_topk_against_oracle(q, db, k, s, i)
    assert stats == {"rows": Q, "resolved_fp32_grade": 0, "resolved_exhaustive": 0}
    s2, i2 = _search(ops, q, db, k, exact=False)      # the sync-free variant: same answer here
    assert torch.equal(i2, i) and torch.equal(s2, s)


def test_topk_search_duplicated_database_rows(ops):
    # 300 exact copies of one row: more ties at the k-th score than any margin holds
    # -> fp32-grade re-screen cannot certify either -> exhaustive pass; ties -> lower index
    Q, N, D, k = 20, 4000, 64, 40
    q = oracle.normalize_l2(_randn(Q, D, seed=23))
    db = oracle.normalize_l2(_randn(N, D, seed=24))
    dup = torch.randperm(N, generator=torch.Generator().manual_seed(25))[:300]
    db[dup] = q[3]
    stats = {}
    s, i = _search(ops, q, db, k, stats=stats)
    want = dup.sort().values[:k]
    assert torch.equal(i[3].cpu(), want)
    assert torch.allclose(s[3].cpu(), torch.ones(k), atol=1e-6)
    assert stats["resolved_exhaustive"] >= 1
    a_s, a_i = oracle.topk_search_f64(q, db, k)
    assert torch.equal(i.cpu(), a_i)


def test_topk_exhaustive_matches_screened_search(ops):
    # the last-line kernel alone, on every row, against the normal path
    from instance_search_b200 import _lib
    Q, N, D, k = 33, 7001, 72, 25
    q = oracle.normalize_l2(_randn(Q, D, seed=26)).cuda()
    db = oracle.normalize_l2(_randn(N, D, seed=27)).cuda()
    s, i = ops.topk_search(q, db, ops.to_bf16(db), k)
    L = _lib.lib()
    rows = torch.arange(Q, dtype=torch.int32, device="cuda")
    s2, i2 = torch.empty_like(s), torch.empty_like(i)
    nb = L.isb_topk_exhaustive_workspace_bytes(Q, N, k)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    _lib.check(L.isb_topk_exhaustive(q.data_ptr(), db.data_ptr(), N, D, k, 0, rows.data_ptr(), Q,
                                     s2.data_ptr(), i2.data_ptr(), ws.data_ptr(), nb, None), "exhaustive")
    torch.cuda.synchronize()
    assert torch.equal(i2, i) and torch.equal(s2, s)


def test_topk_search_idx_offset_and_merge(ops):
    # row-sharded database: per-shard search + merge == single search  
    Q, N, D, k = 150, 9000, 64, 30
    q = oracle.normalize_l2(_randn(Q, D, seed=16))
    db = oracle.normalize_l2(_randn(N, D, seed=17))
    full_s, full_i = _search(ops, q, db, k)
    R = 4
    bounds = [0, 2000, 2500, 7000, 9000]  # uneven shards
    cs, ci = [], []
    for r in range(R):
        shard = db[bounds[r]:bounds[r + 1]].cuda()
        s, i = ops.topk_search(q.cuda(), shard, ops.to_bf16(shard), k, idx_offset=bounds[r])
        cs.append(s), ci.append(i)
    ms, mi = ops.topk_merge(torch.stack(cs), torch.stack(ci))
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)


def test_topk_small_margin_is_not_certified_by_a_one_sample_sigma(ops):
    # k = 1, margin = 0: sigma comes from ONE candidate and can be ~0; the expected-bf16-noise
    # floor keeps such a row from being certified on a gap far below the screen noise.  The
    # result must still be the fp64 arg-max for every row (rows the floor rejects are resolved).
    Q, N, D = 300, 40000, 256
    q = oracle.normalize_l2(_randn(Q, D, seed=51))
    db = oracle.normalize_l2(_randn(N, D, seed=52))
    stats = {}
    s, i = _search(ops, q, db, 1, margin=0, stats=stats)
    a_s, a_i = oracle.topk_search_f64(q, db, 1)
    assert torch.equal(i.cpu(), a_i)
    assert stats["resolved_fp32_grade"] >= 1      # top-2 gaps below 8 sigma exist among 300 rows


def test_topk_search_errors(ops):
    from instance_search_b200 import IsbError
    q = torch.zeros(4, 16).cuda()
    db = torch.zeros(10, 16).cuda()
    with pytest.raises(IsbError):
        ops.topk_search(q, db, ops.to_bf16(db), 11)          # k > N
    with pytest.raises(IsbError):
        ops.topk_search(q.cpu(), db, ops.to_bf16(db), 2)     # CPU tensor: no fallback
    with pytest.raises(IsbError):
        ops.topk_search(q[:, :12].contiguous(), db[:, :12].contiguous(),
                        ops.to_bf16(db[:, :12].contiguous()), 2)  # D % 8 != 0


def test_deferred_pipelined_searches_equal_the_plain_search(ops):
    # defer=True: no host sync inside; the re-rank runs on a side stream under the next batch's
    # screen, two workspaces alternate.  Four different batches queued back to back, resolved
    # afterwards, must equal one-at-a-time searches (planted duplicates make one batch use the
    # certificate's second line at resolve time).
    from instance_search_b200.search import DescriptorIndex
    N, D, k = 30000, 128, 50
    db = oracle.normalize_l2(_randn(N, D, seed=61))
    qs = [oracle.normalize_l2(_randn(300 + 17 * j, D, seed=62 + j)) for j in range(4)]
    db[100:400] = qs[2][5]                      # batch 2, row 5: 300 exact duplicates -> exhaustive line
    index = DescriptorIndex(db.cuda())
    want = [index.search(q.cuda(), k) for q in qs]
    torch.cuda.synchronize()
    queued = [index.search(q.cuda(), k, defer=True) for q in qs]
    n_bad = [t.resolve() for _, _, t in queued]
    torch.cuda.synchronize()
    assert n_bad[2] >= 1 and n_bad[0] == 0
    for (s0, i0), (s1, i1, _) in zip(want, queued):
        assert torch.equal(i0, i1) and torch.equal(s0, s1)


@pytest.mark.parametrize("M,N,K,splits", [(256, 512, 1024, 1), (300, 700, 6400, 3), (1536, 2048, 12544, 4),
                                          (129, 256, 200, 1)])
def test_gemm_pair_and_single_cta_kernels_agree(ops, M, N, K, splits):
    # isb_gemm_nt[_split] runs the CTA-pair kernel (cta_group::2, M = 256) when there are >= 2 row
    # blocks; option gemm_pair=0 keeps the single-CTA kernel: same products, same chunked fp32
    # accumulation order -> identical results
    from instance_search_b200 import _lib
    a, b = _randn(M, K, seed=71).cuda(), _randn(N, K, seed=72).cuda()
    a_hi, a_lo, b_hi, b_lo = ops.to_bf16(a, 0), ops.to_bf16(a, 1), ops.to_bf16(b, 0), ops.to_bf16(b, 1)
    y2 = ops.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, splits=splits, k=K)
    p2 = ops.gemm_nt(a_hi, b_hi, splits=splits, k=K)
    with _lib.options(gemm_pair=0):
        y1 = ops.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, splits=splits, k=K)
        p1 = ops.gemm_nt(a_hi, b_hi, splits=splits, k=K)
    assert torch.equal(y1, y2) and torch.equal(p1, p2)
    want = a.double() @ b.double().t()
    assert (y2.double() - want).abs().max().item() < 2e-5 * want.abs().max().item()


def test_gemm_long_contraction_accumulates_in_chunks(ops):
    # K = 100352 (the whitening projection): the tensor core's fp32 accumulator drifts on long chains
    # (3.6e-4 relative on one 300k-step chain); chunked accumulation keeps the split-operand product at
    # the 16-bit representation floor whatever the number of split-K partitions
    M, N, K = 128, 256, 100352
    g = torch.Generator().manual_seed(73)
    a = torch.relu(torch.randn(M, K, generator=g)).cuda()
    a = a / a.norm(dim=1, keepdim=True)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    a_hi, a_lo, b_hi, b_lo = ops.to_bf16(a, 0), ops.to_bf16(a, 1), ops.to_bf16(b, 0), ops.to_bf16(b, 1)
    want = a.double() @ b.double().t()
    for splits in (1, 9, 37):
        y = ops.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, splits=splits, k=K)
        rel = ((y.double() - want).norm(dim=1) / want.norm(dim=1)).max().item()
        assert rel < 1e-5, (splits, rel)      # 6.6e-6 measured; one 300k-step chain gave 3.6e-4
