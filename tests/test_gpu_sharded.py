"""GPU parity tests of the row-sharded search with candidate exchange
(include/isb.h: isb_topk_candidates -> isb_topk_global_threshold ->
isb_topk_rerank_owned -> isb_topk_merge_certified), SURVEY.md section 8e.

The R shards live on ONE device here and are stepped in lock-step, the two
all-gathers being torch.stack; the stages, their order and their arguments are
those of ShardedIndex.search (tests/test_sharded_gloo.py covers that plumbing
with real processes).  Results must equal the single-index search and the
oracle, whatever the number of shards."""

import pytest
import torch

import oracle
from parity import check_topk_against_oracle

pytestmark = pytest.mark.gpu


def _rows(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return oracle.normalize_l2(torch.randn(n, d, generator=g))


def _lockstep_search(q, db, k, R, kc=None, kl=None):
    """What R ranks compute, on one device.  Returns (scores, idx, n_uncertified, stats).
    kl: entries listed per shard (None: kc, full lists; ShardedIndex.list_length reduces it)."""
    from instance_search_b200 import ops
    from instance_search_b200.search import DescriptorIndex, shard_bounds
    dev = torch.device("cuda:0")
    qd = q.to(dev)
    kk = min(k, db.size(0))
    if kc is None:
        kc = min(kk + ops.DEFAULT_MARGIN, ops.MAX_CANDIDATES)
    shards = [DescriptorIndex(db[lo:hi].to(dev), lo) for lo, hi in shard_bounds(db.size(0), R)]
    kl = kc if kl is None else kl
    cands = [sh.candidates(qd, min(kk, kl), kl) for sh in shards]
    all_screen = torch.stack([c[0] for c in cands])                 # all-gather 1
    thr = ops.topk_global_threshold(all_screen, kc)
    packed_all = torch.stack([sh.rerank_owned(qd, kk, c[0], c[1], thr) for sh, c in zip(shards, cands)])  # all-gather 2
    offs = torch.tensor([lo for lo, _ in shard_bounds(db.size(0), R)], dtype=torch.int64, device=dev)
    s, i, unc_rows, n_unc = ops.topk_merge_certified(packed_all, offs, thr, kk)
    gstat = packed_all[:, :, 2 * kk:].contiguous().view(torch.float32)
    torch.cuda.synchronize()
    scored = gstat[:, :, 1].sum(0)       # candidates re-ranked per query over all shards
    return s, i, int(n_unc.item()), {"scored": scored.cpu(), "thr": thr.cpu(), "unc_rows": unc_rows.cpu(),
                                     "all_screen": all_screen.cpu()}


@pytest.mark.parametrize("R,Q,N,D,k", [
    (2, 33, 5000, 128, 10),
    (3, 70, 20011, 256, 100),     # uneven shards, k + margin = 128
    (8, 17, 4096, 64, 25),
    (4, 5, 300, 32, 100),         # shards (75 rows) shorter than k: padded candidate lists
    (1, 20, 3000, 128, 10),       # R = 1 degenerates to the single-GPU result
])
def test_candidate_exchange_equals_oracle(R, Q, N, D, k):
    q, db = _rows(Q, D, 11), _rows(N, D, 12)
    s, i, n_unc, st = _lockstep_search(q, db, k, R)
    assert n_unc == 0
    check_topk_against_oracle(q, db, k, s, i)
    # the point of the exchange: k + margin exact scores per query IN TOTAL (ties at the
    # threshold may add a few), not per shard
    kc = min(k + 28, 128)
    assert int(st["scored"].min()) >= min(kc, N)
    assert int(st["scored"].max()) <= min(kc, N) + 4


@pytest.mark.parametrize("R,Q,N,D,k", [(8, 60, 40000, 128, 100), (4, 33, 20011, 256, 100), (2, 20, 9000, 64, 100),
                                       (3, 12, 2000, 64, 100)])
def test_reduced_candidate_lists_equal_full_lists(R, Q, N, D, k):
    # every shard lists mean + 6 sigma + 8 candidates instead of k + margin (ShardedIndex.list_length):
    # same result, same number of exact re-ranks in total, and no row needs the second line when
    # the rows are placed at random
    from instance_search_b200.search import ShardedIndex
    q, db = _rows(Q, D, 31), _rows(N, D, 32)
    kc = min(k + 28, 128)
    probe = ShardedIndex.__new__(ShardedIndex)
    probe.world_size, probe.reduce_lists = R, True
    kl = probe.list_length(kc)
    assert kl < kc and R * kl >= kc
    s0, i0, n0, st0 = _lockstep_search(q, db, k, R)
    s1, i1, n1, st1 = _lockstep_search(q, db, k, R, kl=kl)
    assert n0 == 0 and n1 == 0
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    assert torch.equal(st0["thr"], st1["thr"]) and torch.equal(st0["scored"], st1["scored"])
    check_topk_against_oracle(q, db, k, s1, i1)


def test_reduced_list_cut_above_the_threshold_is_reported():
    # a database sorted by content: the queries' neighbours all sit on shard 0, whose reduced list
    # (48 of the 128 wanted) is cut far above the global threshold -> every such row must be
    # reported as uncertified (never answered from the incomplete lists); rows whose neighbours are
    # spread out stay certified
    R, N, D, k = 8, 16000, 64, 100
    g = torch.Generator().manual_seed(5)
    centre = torch.randn(1, D, generator=g)
    db = oracle.normalize_l2(torch.randn(N, D, generator=g))
    db[:600] = oracle.normalize_l2(centre + 0.2 * torch.randn(600, D, generator=g))   # one tight cluster on shard 0
    q = torch.cat([oracle.normalize_l2(centre + 0.2 * torch.randn(10, D, generator=g)),      # near the cluster
                   oracle.normalize_l2(torch.randn(10, D, generator=g))])                    # anywhere
    s, i, n_unc, st = _lockstep_search(q, db, k, R, kl=48)
    bad = set(st["unc_rows"][:n_unc].tolist())
    assert set(range(10)) <= bad                       # cut lists detected
    ok = [r for r in range(20) if r not in bad]
    assert len(ok) >= 8
    o_s, o_i = oracle.topk_search(q, db, k)
    assert torch.equal(i.cpu()[ok], o_i[ok])
    # with full lists no list is cut: whatever the certificate still rejects there is the dense
    # neighbourhood itself (600 near-duplicates within bf16 noise), not the exchange
    s2, i2, n2, st2 = _lockstep_search(q, db, k, R)
    ok2 = [r for r in range(20) if r not in set(st2["unc_rows"][:n2].tolist())]
    assert set(range(10, 20)) <= set(ok2)
    assert torch.equal(i2.cpu()[ok2], o_i[ok2])


def test_candidate_exchange_matches_single_index():
    from instance_search_b200.search import DescriptorIndex
    q, db = _rows(64, 512, 21), _rows(30000, 512, 22)
    one_s, one_i = DescriptorIndex(db.cuda()).search(q.cuda(), 100)
    for R in (2, 5):
        s, i, n_unc, _ = _lockstep_search(q, db, 100, R)
        assert n_unc == 0
        assert torch.equal(i.cpu(), one_i.cpu())
        assert torch.equal(s.cpu(), one_s.cpu())


def test_global_threshold_kernel_is_the_kc_th_best():
    from instance_search_b200 import ops
    g = torch.Generator().manual_seed(5)
    R, Q, kc = 5, 40, 37
    x = torch.randn(R, Q, kc, generator=g)
    x[:, 3, :] = float("-inf")                    # a row with no candidates at all
    x[1:, 4, :] = float("-inf")                   # exactly kc candidates
    x[:, 5, 10:] = float("-inf")                  # 50 candidates, some shards short
    x[:, 6, :] = 0.25                             # all tied
    x[2, 7, :] = x[3, 7, :]                       # duplicated values across shards
    thr = ops.topk_global_threshold(x.cuda()).cpu()
    # a rank beyond the list length (reduced lists): the 20th best of 3 x 9 entries
    x9 = torch.randn(3, 11, 9)
    want = x9.permute(1, 0, 2).reshape(11, 27).sort(dim=1, descending=True).values[:, 19]
    assert torch.equal(ops.topk_global_threshold(x9.cuda(), 20).cpu(), want)
    # ... fewer than that exist while a list is full: nothing can be concluded -> +inf
    x9[:, 4, 3:] = float("-inf")
    x9[1, 4] = torch.arange(9.0)
    assert ops.topk_global_threshold(x9.cuda(), 20).cpu()[4] == float("inf")
    x9[1, 4, 8] = float("-inf")
    assert ops.topk_global_threshold(x9.cuda(), 20).cpu()[4] == float("-inf")
    flat = x.permute(1, 0, 2).reshape(Q, R * kc)
    want = flat.sort(dim=1, descending=True).values[:, kc - 1]
    assert torch.equal(thr, want)
    assert thr[3] == float("-inf") and thr[4] == x[0, 4].min()
    y = x.clone()
    y[:, 8, 5:] = float("-inf")                   # 25 < kc candidates -> nothing dropped -> -inf
    assert ops.topk_global_threshold(y.cuda()).cpu()[8] == float("-inf")


def test_duplicate_dense_rows_are_flagged_and_resolvable():
    """300 exact duplicates of the best match: more rows tie with the k-th score than the
    margin holds, the global certificate must reject those queries (ShardedIndex.search then
    answers them with every shard's own certified search + isb_topk_merge)."""
    from instance_search_b200 import ops
    from instance_search_b200.search import DescriptorIndex, shard_bounds
    q, db = _rows(12, 128, 31), _rows(6000, 128, 32)
    db[1000:1300] = q[2]          # 300 duplicates, spread over shards 0 and 1 of 4
    k, R = 100, 4
    s, i, n_unc, st = _lockstep_search(q, db, k, R)
    assert n_unc >= 1 and 2 in st["unc_rows"][:n_unc].tolist()
    # the resolve path of ShardedIndex.search, in lock-step
    rows = st["unc_rows"][:n_unc].long().sort().values
    dev = torch.device("cuda:0")
    ls, li = [], []
    for lo, hi in shard_bounds(db.size(0), R):
        a, b = DescriptorIndex(db[lo:hi].to(dev), lo).search(q[rows].to(dev), k)
        ls.append(a), li.append(b)
    rs, ri = ops.topk_merge(torch.stack(ls), torch.stack(li))
    s.index_copy_(0, rows.to(dev), rs)
    i.index_copy_(0, rows.to(dev), ri)
    want_s, want_i = oracle.topk_search_f64(q, db, k)
    # duplicates tie exactly: the k best are the k lowest-indexed duplicates
    assert torch.equal(i.cpu()[2], torch.arange(1000, 1100))
    others = [r for r in range(12) if r != 2]
    assert torch.equal(i.cpu()[others], want_i[others])
    assert torch.allclose(s.cpu().double(), want_s, rtol=2e-7, atol=1e-9)


def _merge_reference(cs, ci, k):
    """contract of isb_topk_merge: best first, ties -> lower index, index < 0 last / dropped"""
    R, Q, kk = cs.shape
    s = cs.permute(1, 0, 2).reshape(Q, R * kk).clone()
    i = ci.permute(1, 0, 2).reshape(Q, R * kk)
    s[i < 0] = float("-inf")
    key = i.clone()
    key[i < 0] = torch.iinfo(torch.int64).max
    o1 = key.argsort(dim=1, stable=True)
    s1, i1 = s.gather(1, o1), i.gather(1, o1)
    o2 = s1.argsort(dim=1, descending=True, stable=True)
    s2, i2 = s1.gather(1, o2)[:, :k], i1.gather(1, o2)[:, :k]
    s2 = torch.where(i2 < 0, torch.full_like(s2, float("-inf")), s2)
    return s2, i2


@pytest.mark.parametrize("R,Q,k,levels,n_invalid", [
    (2, 50, 10, 0, 0),        # continuous scores
    (8, 37, 100, 0, 3),       # the 8-GPU headline shape, a few padded entries
    (8, 33, 100, 7, 0),       # 7 distinct score values: hundreds of ties at the cut -> index search
    (4, 9, 128, 3, 40),       # k = 128, heavy ties, many invalid
    (3, 5, 100, 0, 295),      # fewer valid entries (5) than k
    (2, 6, 300, 5, 10),       # k > 128: the block-wide bitonic kernel
])
def test_topk_merge_kernels(R, Q, k, levels, n_invalid):
    from instance_search_b200 import ops
    g = torch.Generator().manual_seed(R * 1000 + k)
    cs = torch.randn(R, Q, k, generator=g)
    if levels:
        cs = (cs * 1.5).round().clamp(-(levels // 2), levels // 2) / 4
    ci = torch.stack([torch.randperm(1 << 20, generator=g)[:R * k].reshape(R, k) for _ in range(Q)], 1)
    ci[0, 0, 0] = (1 << 40) + 5           # a global index beyond 32 bits
    for q in range(Q):
        bad = torch.randperm(R * k, generator=g)[:n_invalid]
        ci[bad // k, q, bad % k] = -1
        cs[bad // k, q, bad % k] = float("-inf")
    s, i = ops.topk_merge(cs.cuda(), ci.cuda())
    want_s, want_i = _merge_reference(cs, ci, k)
    assert torch.equal(i.cpu(), want_i)
    assert torch.equal(s.cpu(), want_s)


def test_persisted_database_loads_into_an_index(tmp_path):
    """store.py: descriptors written once, every shard loads its rows straight to the GPU."""
    from instance_search_b200 import ops, store
    from instance_search_b200.search import DescriptorIndex, shard_bounds
    q, db = _rows(40, 96, 51), _rows(9001, 96, 52)
    path = str(tmp_path / "db.isbd")
    store.write_descriptors(path, db.cuda(), chunk_rows=1000)       # from device memory, in chunks
    index = store.load_index(path)
    assert len(index) == 9001 and torch.equal(index.db_f32.cpu(), db)
    s, i = index.search(q.cuda(), 20)
    check_topk_against_oracle(q, db, 20, s, i)
    f = store.DescriptorFile(path)
    for lo, hi in shard_bounds(9001, 4):
        assert torch.equal(f.load_rows(lo, hi, "cuda:0", chunk_rows=700).cpu(), db[lo:hi])
