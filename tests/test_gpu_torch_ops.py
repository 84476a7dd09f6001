"""GPU: every torch custom op (torch.ops.isb.*) called through the dispatcher and compared with
the oracle (or with the ctypes wrapper it stands for, which the other GPU tests pin to the oracle)."""

import pytest
import torch

import oracle
from parity import check_descriptors, check_topk_against_oracle

pytestmark = pytest.mark.gpu


def _randn(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_row_ops_and_gemm_through_the_dispatcher():
    import instance_search_b200  # noqa: F401  (registers the ops)
    o = torch.ops.isb
    x, s, g = _randn(9, 333, seed=1), _randn(333, seed=2), _randn(9, 333, seed=3)
    xd = x.cuda()
    assert torch.allclose(o.l2norm_rows(xd, 1e-10).cpu(), oracle.normalize_l2(x), rtol=1e-6, atol=1e-9)
    assert torch.equal(o.shift_rows(xd, s.cuda()).cpu(), oracle.shift(x, s))
    assert torch.allclose(o.col_sums(g.cuda()).cpu(), g.sum(0), rtol=1e-5, atol=1e-5)
    xr = x.clone().requires_grad_(True)
    oracle.normalize_l2(xr).backward(g)
    got = o.l2norm_rows_backward(xd, g.cuda(), 1e-10).cpu()
    assert torch.allclose(got, xr.grad, rtol=1e-5, atol=2e-6 * xr.grad.abs().max().item())
    a, b = _randn(70, 136, seed=4).cuda(), _randn(90, 136, seed=5).cuda()
    a_hi, a_lo, b_hi, b_lo = o.f32_to_bf16(a, 0, 0), o.f32_to_bf16(a, 1, 0), o.f32_to_bf16(b, 0, 0), o.f32_to_bf16(b, 1, 0)
    want = a.double() @ b.double().t()
    assert (o.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, 1).double() - want).abs().max() < 2e-5 * want.abs().max()
    assert (o.gemm_nt(a_hi, b_hi, 1).double() - a_hi[:, :136].double() @ b_hi[:, :136].double().t()).abs().max() \
        < 1e-5 * want.abs().max()
    E = oracle.normalize_l2(_randn(300, 72, seed=6))
    # (16-bit operands: 4e-6 / sqrt(D) rms per entry -> a few 1e-6 at the maximum over 90k entries at D = 72)
    assert torch.allclose(o.all_pairs_similarities(E.cuda(), 3).cpu(), E @ E.t(), rtol=1e-5, atol=5e-6)
    with pytest.raises(NotImplementedError):
        o.l2norm_rows(x, 1e-10)                       # CPU tensor: no kernel, no fallback


def test_search_ops_single_and_sharded_through_the_dispatcher():
    import instance_search_b200  # noqa: F401
    o = torch.ops.isb
    Q, N, D, k = 200, 9000, 64, 30
    q, db = oracle.normalize_l2(_randn(Q, D, seed=7)), oracle.normalize_l2(_randn(N, D, seed=8))
    qd, dbd = q.cuda(), db.cuda()
    s, i = o.topk_search(qd, dbd, o.f32_to_bf16(dbd, 0, 0), k, -1, 0)
    check_topk_against_oracle(q, db, k, s, i)
    # the sharded pipeline, stage by stage, over three uneven shards == the single search
    bounds = [0, 2500, 7000, 9000]
    kc = k + 28
    shards = [dbd[bounds[r]:bounds[r + 1]].contiguous() for r in range(3)]
    cand = [o.topk_candidates(qd, o.f32_to_bf16(sh, 0, 0), k, kc) for sh in shards]
    thr = o.topk_global_threshold(torch.stack([c[0] for c in cand]))
    packed = torch.stack([o.topk_rerank_owned(qd, sh, k, c[0], c[1], thr) for sh, c in zip(shards, cand)])
    offs = torch.tensor(bounds[:3], dtype=torch.int64, device="cuda")
    ms, mi, unc_rows, n_unc = o.topk_merge_certified(packed, offs, thr, k)
    assert int(n_unc) == 0 and torch.equal(mi, i) and torch.equal(ms, s)
    # plain merge of per-shard results
    per = [o.topk_search(qd, sh, o.f32_to_bf16(sh, 0, 0), k, -1, bounds[r]) for r, sh in enumerate(shards)]
    ms2, mi2 = o.topk_merge(torch.stack([p[0] for p in per]), torch.stack([p[1] for p in per]))
    assert torch.equal(mi2, i) and torch.equal(ms2, s)


def test_region_ops_through_the_dispatcher():
    import instance_search_b200  # noqa: F401
    o = torch.ops.isb
    g = torch.Generator().manual_seed(9)
    B, C, H, W, ncls, D, k = 5, 64, 14, 13, 20, 32, 6
    Kin = C * 49
    x = torch.relu(torch.randn(B, C, H, W, generator=g))
    cls_w, cls_b = torch.randn(ncls, C, generator=g) / C ** 0.5, 0.01 * torch.randn(ncls, generator=g)
    shift, lin_b = 0.01 * torch.randn(Kin, generator=g), 0.01 * torch.randn(D, generator=g)
    lin_w = torch.randn(D, Kin, generator=g) / Kin ** 0.5
    od, oc, oi, on = oracle.region_descriptor_forward(x, cls_w, cls_b, shift, lin_w, lin_b, k, (7, 7))
    xd, cw, cb, sh, lw, lb = (t.cuda() for t in (x, cls_w, cls_b, shift, lin_w, lin_b))
    cw_hi, cw_lo = o.f32_to_bf16(cw, 0, 0), o.f32_to_bf16(cw, 1, 0)
    lw_hi, lw_lo = o.f32_to_bf16(lw, 0, 0), o.f32_to_bf16(lw, 1, 0)
    amax = float(cls_w.abs().max())
    # the composite op
    d, c, i, n = o.region_descriptors(xd, cw, cw_hi, cw_lo, cb, sh, lw_hi, lw_lo, lb, 7, 7, k, amax)
    assert torch.equal(i.cpu(), oi) and torch.equal(n.cpu().long(), on)
    assert torch.allclose(c.cpu(), oc, rtol=1e-5, atol=2e-6)
    check_descriptors(d, od)
    # the same from its stages: select (k + 2 windows) -> gather -> exact logits -> project -> finalize
    ke = k + 2
    idx_e, nsel_e, _, norm_e, approx_e, runner, n1 = o.region_select(xd, cw, cw_hi, cw_lo, cb, 7, 7, ke, 10, False)
    U_hi, U_lo, win_mean = o.region_gather(xd, idx_e, nsel_e, norm_e, sh, 7, 7, k, 3)
    idx, norm, nsel, cls_out, changed, n_changed, n2 = o.region_logits(win_mean, cw, cb, k, nsel_e, idx_e, norm_e,
                                                                       approx_e, runner, amax)
    assert int(n1[0]) == 0 and int(n2[0]) == 0 and int(n_changed) == 0
    assert torch.equal(idx.cpu(), oi) and torch.allclose(cls_out.cpu(), oc, rtol=1e-5, atol=2e-6)
    y = o.gemm_nt_split(U_hi, U_lo, lw_hi, lw_lo, 4)
    check_descriptors(o.descriptor_finalize(y, lb, nsel, 1e-10), od)
    # DescriptorNet head
    x7 = torch.relu(torch.randn(3, C, 7, 7, generator=g))
    want = oracle.descriptor_forward(x7, shift, lin_w, lin_b)
    check_descriptors(o.global_descriptors(x7.cuda(), sh, lw_hi, lw_lo, lb), want)


def test_mining_metrics_dba_loss_ops_through_the_dispatcher():
    import instance_search_b200  # noqa: F401
    o = torch.ops.isb
    g = torch.Generator().manual_seed(10)
    N, D, per = 1024, 64, 8
    lab = torch.arange(N) // per
    E = oracle.normalize_l2(torch.randn(N // per, D, generator=g)[lab] + 0.5 * torch.randn(N, D, generator=g))
    anchors = torch.arange(0, 400)
    positives = (anchors // per) * per + (anchors % per + 1) % per
    S = E @ E.t()
    for semi in (False, True):
        neg, nsim, psim = o.select_negatives(E.cuda(), lab.int().cuda(), anchors.cuda(), positives.cuda(), semi, 1)
        want = oracle.select_negatives(S, lab, list(zip(anchors.tolist(), positives.tolist())), semi)
        assert int((neg.cpu() != want).sum()) <= 1
    sim = torch.randn(12, 500, generator=g)
    v, j = o.row_kth_largest(sim.cuda(), 2)
    assert torch.equal(v.cpu(), sim.kthvalue(499, dim=1).values)
    cols = torch.randint(0, 500, (12, 4), generator=g).int()
    ranks = o.row_ranks(sim.cuda(), cols.cuda()).cpu()
    order = sim.argsort(dim=1, descending=True, stable=True)
    pos = torch.empty_like(order)
    pos.scatter_(1, order, torch.arange(500).expand(12, 500))
    assert torch.equal(ranks.long(), pos.gather(1, cols.long()))
    ref_set = [(None, "L%d" % l, "r%d" % n) for n, l in enumerate(lab[:200].tolist())]
    out = o.instance_avg(E[:200].cuda(), lab[:200].int().cuda(), -1)
    assert torch.allclose(out.cpu(), oracle.instance_avg(E[:200], ref_set, -1), rtol=1e-5, atol=1e-6)
    a, p, n = (oracle.normalize_l2(_randn(20, 32, seed=s)) for s in (11, 12, 13))
    loss, clamp = o.triplet_loss_forward(a.cuda(), p.cuda(), n.cuda(), 0.2, True, True)
    assert torch.allclose(loss.cpu(), oracle.triplet_loss(a, p, n, 0.2, True, True), rtol=1e-5, atol=1e-7)
    ga, gp, gn = o.triplet_loss_backward(a.cuda(), p.cuda(), n.cuda(), clamp, torch.ones(1, device="cuda"), True, True)
    ar = a.clone().requires_grad_(True)
    oracle.triplet_loss(ar, p, n, 0.2, True, True).backward()
    assert torch.allclose(ga.cpu(), ar.grad, rtol=1e-5, atol=1e-7)
