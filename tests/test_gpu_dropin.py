"""GPU parity of the drop-in operator surface (model/custom_modules.py,
model/siamese.py, utils/metrics.py, utils/train_siamese.py, test/instance_avg.py
mirrors) against the oracle and the reference-minted goldens."""

import math

import pytest
import torch
import torch.nn as nn

import oracle
from conftest import load_golden
from parity import check_descriptors
from test_host_cpu import ToyNet

pytestmark = pytest.mark.gpu


def _randn(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------ custom modules
@pytest.mark.parametrize("shape", [(4, 33), (3, 2048), (2, 100352)])
def test_normalize_l2_forward_backward(shape):
    from instance_search_b200.model.custom_modules import NormalizeL2
    x = _randn(*shape, seed=1)
    g = _randn(*shape, seed=2)
    xd = x.cuda().requires_grad_(True)
    y = NormalizeL2()(xd)
    y.backward(g.cuda())
    xr = x.clone().requires_grad_(True)
    yr = oracle.normalize_l2(xr)
    yr.backward(g)
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-6, atol=1e-9)
    # reference backward (model/custom_modules.py:59-67) == autograd of the forward
    scale = xr.grad.abs().max().item()
    assert torch.allclose(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=2e-6 * scale)


def test_shift_forward_backward():
    from instance_search_b200.model.custom_modules import Shift
    m = Shift(77).cuda()
    assert list(m.state_dict().keys()) == ["param"]
    m.param.data.copy_(_randn(77, seed=3))
    x = _randn(9, 77, seed=4).cuda().requires_grad_(True)
    g = _randn(9, 77, seed=5).cuda()
    y = m(x)
    y.backward(g)
    assert torch.equal(y.detach().cpu(), oracle.shift(x.detach().cpu(), m.param.detach().cpu()))
    assert torch.equal(x.grad, g)
    assert torch.allclose(m.param.grad.cpu(), g.cpu().sum(0), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("normalized,size_average", [(True, True), (True, False), (False, True)])
def test_triplet_loss(normalized, size_average):
    from instance_search_b200.model.custom_modules import TripletLoss
    B, D, margin = 37, 96, 0.2
    a, p, n = (oracle.normalize_l2(_randn(B, D, seed=s)) for s in (6, 7, 8))
    p = oracle.normalize_l2(a + 0.3 * p)        # close positives ...
    n[::2] = oracle.normalize_l2(a[::2] + 0.2 * n[::2])   # ... and hard negatives on every other row
    crit = TripletLoss(margin, size_average=size_average, normalized=normalized)
    ad, pd, nd = (t.cuda().requires_grad_(True) for t in (a, p, n))
    loss = crit(ad, pd, nd)
    (loss * 1.5).backward()
    want = oracle.triplet_loss(a, p, n, margin, size_average, normalized)
    assert loss.shape == (1,) and torch.allclose(loss.cpu(), want, rtol=1e-5, atol=1e-7)
    ar, pr, nr = (t.clone().requires_grad_(True) for t in (a, p, n))
    (oracle.triplet_loss(ar, pr, nr, margin, size_average, normalized) * 1.5).backward()
    for got, ref in ((ad, ar), (pd, pr), (nd, nr)):
        assert torch.allclose(got.grad.cpu(), ref.grad, rtol=1e-5, atol=1e-7)
    assert 0 < int((ad.grad.abs().sum(1) == 0).sum()) < B   # some rows clamped, some not


# ------------------------------------------------------------------ metrics
def _sets(ref_lab, test_lab):
    return ([(None, "L%d" % l, "t%d" % i) for i, l in enumerate(test_lab)],
            [(None, "L%d" % l, "r%d" % i) for i, l in enumerate(ref_lab)])


@pytest.mark.parametrize("kth", [1, 2, 3])
def test_metrics_golden(kth):
    from instance_search_b200.utils import metrics
    g = load_golden("metrics_tiny")
    test_set, ref_set = _sets(g["ref_lab"].tolist(), g["test_lab"].tolist())
    sim = g["sim"].cuda()
    p, c, t, ms, ml = metrics.precision1(sim, test_set, ref_set, kth)
    assert [p, c, t] == g["prec_kth%d" % kth].tolist()
    assert ms.shape == (sim.size(0), 1) and torch.equal(ms.cpu(), g["max_sim_kth%d" % kth])
    assert [int(s[1:]) for s in ml] == g["max_label_kth%d" % kth].tolist()
    assert metrics.mean_avg_precision(sim, test_set, ref_set, kth) == g["map_kth%d" % kth]   # bit-exact
    for i, a in enumerate(g["ap_kth%d" % kth].tolist()):
        o = metrics.avg_precision(sim, i, test_set, ref_set, kth)
        assert (o is None and math.isnan(a)) or o == a


def test_metrics_random_vs_oracle():
    from instance_search_b200.utils import metrics
    gen = torch.Generator().manual_seed(9)
    Q, N = 60, 3000
    sim = torch.randn(Q, N, generator=gen)
    ref_lab = torch.randint(0, 40, (N,), generator=gen).tolist()
    test_lab = torch.randint(0, 45, (Q,), generator=gen).tolist()   # some queries without positives
    test_set, ref_set = _sets(ref_lab, test_lab)
    for kth in (1, 2):
        want = oracle.precision1(sim, test_set, ref_set, kth)
        got = metrics.precision1(sim.cuda(), test_set, ref_set, kth)
        assert got[:3] == want[:3] and got[4] == want[4] and torch.equal(got[3].cpu(), want[3])
        assert metrics.mean_avg_precision(sim.cuda(), test_set, ref_set, kth) == \
            oracle.mean_avg_precision(sim, test_set, ref_set, kth)


def test_row_ranks_ties_and_unused_slots():
    from instance_search_b200 import ops
    sim = torch.tensor([[0.5, 0.9, 0.5, 0.1, 0.9], [1.0, 2.0, 3.0, 4.0, 5.0]]).cuda()
    cols = torch.tensor([[2, 0, 4, -1], [0, 4, -1, -1]], dtype=torch.int32).cuda()
    r = ops.row_ranks(sim, cols).cpu().tolist()
    assert r == [[3, 2, 1, -1], [4, 0, -1, -1]]       # ties -> lower column first
    v, i = ops.row_kth_largest(sim, 2)
    assert i.tolist() == [4, 3] and v.tolist() == [pytest.approx(0.9), 4.0]


# ------------------------------------------------------------------ DBA
def test_instance_avg_golden():
    from instance_search_b200.test.instance_avg import instance_avg
    g = load_golden("instance_avg_tiny")
    ref_set = [(None, "L%d" % l, "r%d" % i) for i, l in enumerate(g["ref_lab"].tolist())]
    for k in (-1, 0, 2, 100):
        out, ds = instance_avg(0, g["emb"].cuda(), ref_set, None, k)
        assert ds is ref_set
        assert torch.allclose(out.cpu(), g["out_k%d" % k], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------ nets
def _toy_region_net(seed, k=6, D=16, C=8, ncls=5):
    from instance_search_b200.model.siamese import RegionDescriptorNet
    torch.manual_seed(seed)
    net = RegionDescriptorNet(ToyNet(C, ncls), k, D, (7, 7))
    net.feature_reduc1[1].param.data.normal_(0, 0.01)
    return net.cuda().eval()


def _oracle_head(net, fmap):
    conv, shift, lin = net.classifier[0], net.feature_reduc1[1], net.feature_reduc1[2]
    return oracle.region_descriptor_forward(
        fmap.cpu(), conv.weight.detach().cpu().view(conv.out_channels, -1), conv.bias.detach().cpu(),
        shift.param.detach().cpu(), lin.weight.detach().cpu(), lin.bias.detach().cpu(), net.k,
        net.feature_size2d)


def test_region_descriptor_net_eval_matches_oracle_and_reloads():
    net = _toy_region_net(0)
    x = _randn(5, 3, 40, 48, seed=11).cuda()       # trunk stride 2 -> 20 x 24 map, 14 x 18 windows
    with torch.no_grad():
        fmap = net.features(x)
        desc = net(x)                               # eval: descriptor only (model/siamese.py:231)
        d2, cls_out = net.forward_single(x)
    od, oc, _, _ = _oracle_head(net, fmap)
    assert desc.shape == (5, 16) and cls_out.shape == (5, 5, 6)
    assert torch.equal(desc, d2)
    check_descriptors(desc, od)
    assert torch.allclose(cls_out.cpu(), oc, rtol=1e-5, atol=2e-6)
    # parameters change in place (optimizer step / load_state_dict) -> cached operands rebuilt
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    other = _toy_region_net(1)
    other.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(other(x), desc)
        net.feature_reduc1[2].weight.mul_(-1.0)
        net.feature_reduc1[2].bias.mul_(-1.0)
        assert torch.allclose(net(x), -desc, atol=1e-6)


def _train_mode_grads(net, x, g_desc, g_cls, composed):
    """Gradients of (desc, cls_out) w.r.t. the head parameters and the input images, through the
    fused autograd head or through the reference composition (_forward_single_composed per image)."""
    for p in net.parameters():
        p.grad = None
    xi = x.clone().requires_grad_(True)
    fmap = net.features(xi)
    if composed:
        outs = [net._forward_single_composed(fmap[b:b + 1]) for b in range(fmap.size(0))]
        desc, cls_out = torch.cat([d for d, _ in outs], 0), torch.cat([c for _, c in outs], 0)
    else:
        desc, cls_out = net.forward_single(xi)
    ((desc * g_desc).sum() + (cls_out * g_cls).sum()).backward()
    conv, shift, lin = net.classifier[0], net.feature_reduc1[1], net.feature_reduc1[2]
    grads = {"x": xi.grad, "cls_w": conv.weight.grad, "cls_b": conv.bias.grad, "shift": shift.param.grad,
             "lin_w": lin.weight.grad, "lin_b": lin.bias.grad}
    return desc.detach(), cls_out.detach(), {k: v.detach().clone() for k, v in grads.items()}


@pytest.mark.parametrize("B,h,w,k", [(3, 32, 32, 6), (2, 36, 44, 6), (4, 16, 18, 12)])
def test_region_descriptor_net_train_mode_fused_backward_matches_composed_path(B, h, w, k):
    # training (model/siamese.py:225-229): the fused head under autograd (regions_autograd) against the
    # reference composition per image on torch's conv / linear autograd -- outputs and every gradient
    net = _toy_region_net(2, k=k)
    for p in list(net.classifier.parameters()) + list(net.feature_reduc1.parameters()):
        p.requires_grad = True
    x = _randn(B, 3, h, w, seed=12).cuda()
    torch.backends.cudnn.allow_tf32 = False          # the composed path's conv / linear in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    from instance_search_b200.model.nn_utils import set_net_train
    set_net_train(net, True)                              # train mode, BatchNorm kept frozen (as the reference)
    out = net(x, x)                                       # train: tuple of (desc, cls_out) per input
    assert isinstance(out, tuple) and len(out) == 2 and out[0][0].requires_grad
    g_desc = _randn(B, 16, seed=13).cuda()
    g_cls = _randn(B, 5, k, seed=14).cuda()
    d0, c0, want = _train_mode_grads(net, x, g_desc, g_cls, composed=True)
    d1, c1, got = _train_mode_grads(net, x, g_desc, g_cls, composed=False)
    check_descriptors(d1, d0)
    assert torch.allclose(c1, c0, rtol=1e-4, atol=1e-5)
    for name in want:
        scale = want[name].abs().max().item()
        err = (got[name] - want[name]).abs().max().item()
        assert err <= 1e-5 * scale + 1e-9, (name, err, scale)
    # eval-mode forward of the same net is the same descriptor
    net.eval()
    with torch.no_grad():
        check_descriptors(net(x), d1, l2_tol=1e-6)


def test_descriptor_net_eval_matches_oracle():
    from instance_search_b200.model.siamese import DescriptorNet
    torch.manual_seed(3)
    net = DescriptorNet(ToyNet(8, 5), 12, (7, 7)).cuda().eval()
    net.feature_reduc1[1].param.data.normal_(0, 0.01)
    x = _randn(4, 3, 14, 14, seed=13).cuda()       # -> 7 x 7 map
    with torch.no_grad():
        d = net(x)
        fmap = net.features(x)
    shift, lin = net.feature_reduc1[1], net.feature_reduc1[2]
    want = oracle.descriptor_forward(fmap.cpu(), shift.param.detach().cpu(), lin.weight.detach().cpu(),
                                     lin.bias.detach().cpu())
    check_descriptors(d, want)


# ------------------------------------------------------------------ harness functions
class _P(object):
    cuda_device, feature_dim, embeddings_cuda_size, train_bn = 0, 16, 2 ** 30, False


def test_get_embeddings_similarities_and_descriptor_test():
    from instance_search_b200.train.siamese_regions import get_embeddings, NegativeSelector
    from instance_search_b200.utils import train_siamese as ts
    net = _toy_region_net(4)
    gen = torch.Generator().manual_seed(14)
    ds = [(torch.randn(3, 32, 32, generator=gen), "L%d" % (i % 5), "im%d" % i) for i in range(23)]
    ds[7] = (torch.randn(3, 40, 32, generator=gen), "L2", "odd")          # a different size mid-list
    emb = get_embeddings(net, ds, 0, 16, batch_size=4)
    one = torch.cat([get_embeddings(net, [t], 0, 16) for t in ds])         # the reference's one-by-one loop
    assert torch.allclose(emb, one, atol=1e-6)
    S, dev = ts.get_similarities(_P, get_embeddings, net, ds)
    assert dev == 0 and net.training and not net.features[1].training      # back in train mode, BN frozen
    assert torch.allclose(S.cpu(), emb.cpu() @ emb.cpu().t(), rtol=1e-5, atol=2e-6)
    net.eval()
    res = ts.test_descriptor_net(_P, get_embeddings, net, ds[:9], ds[9:], kth=1)
    sim = emb[:9].cpu() @ emb[9:].cpu().t()
    p1 = oracle.precision1(sim, ds[:9], ds[9:], 1)
    assert res[1] == p1[1] and res[2] == 9
    assert abs(res[6] - oracle.mean_avg_precision(sim, ds[:9], ds[9:], 1)) < 1e-12
    assert abs(res[5] - float(p1[3].sum())) < 1e-4
    # negative selection for a few couples == oracle on the materialised matrix
    sel = NegativeSelector(emb, ds)
    couples = [(0, 5), (1, 6), (2, 12), (3, 8)]
    lab = torch.tensor([i % 5 for i in range(23)])
    lab[7] = 2
    for semi, epoch in ((True, 0), (False, 5)):
        got = sel.select(couples, epoch, 2)
        want = oracle.select_negatives(S.cpu(), lab, couples, semi).tolist()
        assert [(-1 if g is None else g) for g in got] == want


def test_siamese_regions_test_evaluate_matches_the_oracle(capsys):
    # the evaluation section of the reference's test/siamese_regions_test.py (main, :72-88):
    # embeddings of both sets, P@1 and mAP, then the same after database-side augmentation
    from instance_search_b200.train.siamese_regions import get_embeddings
    from instance_search_b200.test.siamese_regions_test import evaluate
    net = _toy_region_net(4)
    gen = torch.Generator().manual_seed(21)
    ds = [(torch.randn(3, 32, 32, generator=gen), "L%d" % (i % 4), "im%d" % i) for i in range(26)]
    test_set, ref_set = ds[:8], ds[8:]
    res = evaluate(net, test_set, ref_set, device=0, dba=-1)
    assert not net.training
    t_emb = get_embeddings(net, test_set, 0, 16).cpu()
    r_emb = get_embeddings(net, ref_set, 0, 16).cpu()
    sim = t_emb @ r_emb.t()
    p1 = oracle.precision1(sim, test_set, ref_set, 1)
    prec1, c, t, mAP = res["plain"]
    assert (c, t) == (p1[1], 8) and abs(prec1 - p1[0]) < 1e-12
    assert abs(mAP - oracle.mean_avg_precision(sim, test_set, ref_set, 1)) < 1e-9
    d_emb = oracle.instance_avg(r_emb, ref_set, -1)
    sim_d = t_emb @ d_emb.t()
    p1d = oracle.precision1(sim_d, test_set, ref_set, 1)
    prec1, c, t, mAP = res["dba"]
    assert (c, t) == (p1d[1], 8)
    assert abs(mAP - oracle.mean_avg_precision(sim_d, test_set, ref_set, 1)) < 1e-6
    out = capsys.readouterr().out
    assert "Descriptor (TEST): " in out and "Descriptor (TEST DBA k=-1): " in out


def test_get_similarities_streams_to_the_host_when_the_placement_rule_says_so():
    # utils/train_siamese.py:30-43: a matrix above P.embeddings_cuda_size lives on the host; the
    # product still runs on the GPU, in row blocks, and the N x N matrix never exists in HBM
    from instance_search_b200.train.siamese_regions import get_embeddings
    from instance_search_b200.utils import train_siamese as ts

    class Small(_P):
        embeddings_cuda_size = 1024        # 23 x 23 x 4 > 1024 -> device -1
    net = _toy_region_net(4)
    gen = torch.Generator().manual_seed(15)
    ds = [(torch.randn(3, 32, 32, generator=gen), "L%d" % (i % 5), "im%d" % i) for i in range(23)]
    S, dev = ts.get_similarities(Small, get_embeddings, net, ds)
    assert dev == -1 and not S.is_cuda and S.shape == (23, 23)
    net.eval()
    emb = get_embeddings(net, ds, 0, 16).cpu()
    assert torch.allclose(S, emb @ emb.t(), rtol=1e-5, atol=2e-6)
    blocks = ts._similarities_to_host(emb.cuda(), block_rows=7)         # ragged last block
    assert torch.allclose(blocks, S, rtol=0, atol=1e-6)


def test_classif_regions_embeddings_match_the_oracle():
    # the classification track's sub-window embedding (train/classif_regions.py:107-132) on the
    # fused head: class scores at the window of highest maximal activation, L2-normalised
    from instance_search_b200.model.siamese import TuneClassifSub
    from instance_search_b200.train.classif_regions import get_embeddings
    torch.manual_seed(6)
    net = TuneClassifSub(ToyNet(8, 5), 7, (7, 7)).cuda().eval()          # last FC resized to 7 classes
    gen = torch.Generator().manual_seed(16)
    ds = [(torch.randn(3, 36, 40, generator=gen), "L%d" % (i % 3), "im%d" % i) for i in range(11)]
    emb = get_embeddings(net, ds, 0, 7, batch_size=4)
    assert emb.shape == (11, 7)
    conv = net.classifier[0]
    for i, (im, _, _) in enumerate(ds):
        with torch.no_grad():
            fmap = net.features(im.unsqueeze(0).cuda()).cpu()
        want, _ = oracle.classif_regions_embedding(fmap, conv.weight.detach().cpu().view(7, -1),
                                                   conv.bias.detach().cpu(), (7, 7))
        assert torch.allclose(emb[i].cpu(), want[0], rtol=1e-5, atol=2e-6)
