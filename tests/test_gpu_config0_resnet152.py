"""BASELINE.json configs[0] on the GPU path: torchvision ResNet-152 trunk (PyTorch) ->
TuneClassifSub -> drop-in RegionDescriptorNet (fused CUDA head) -> top-10 search, against the
reference's CPU path (the oracle head, model/siamese.py:185-223 per image, on the SAME trunk
output; sim = mm, first 10 of the descending sort -- test/siamese_regions_test.py:72-79 with the
net of train/siamese_regions.py:157-168).

Sizes are reduced from the config's 256 queries x 1000 database images to what the CPU path
finishes in about a minute (every image costs the oracle one 822 MB GEMV per selected window).
The net is random-init (no network for the pretrained weights); its BatchNorm
statistics are calibrated on one batch -- an uncalibrated random ResNet-152 maps every image to
the same direction (activations ~1e8, all cosines 0.9999+), which leaves nothing to rank.
Images are smooth random fields (upsampled 7 x 7 noise) so that different images give different
features; queries are noisy copies of database images."""

import pytest
import torch
import torch.nn.functional as F

import oracle
from parity import check_descriptors, DESC_L2_TOL, record

pytestmark = pytest.mark.gpu

MEAN = torch.tensor([0.36, 0.30, 0.28]).view(1, 3, 1, 1)    # mean_std.ipynb cell 4
STD = torch.tensor([0.21, 0.20, 0.20]).view(1, 3, 1, 1)


class _P(object):
    """The fields get_siamese_net reads (train/siamese_regions_p.py:21-113, train/global_p.py)."""
    cnn_model, num_classes, feature_size2d, untrained_blocks = "resnet152", 464, (7, 7), -1
    regions_k, feature_dim, cuda_device = 6, 2048, 0
    classif_model, preload_net = None, None


def _images(n, h, w, gen):
    base = torch.rand(n, 3, 7, 7, generator=gen)
    return F.interpolate(base, size=(h, w), mode="bilinear", align_corners=False)


@pytest.fixture(scope="module")
def net():
    import torchvision
    from instance_search_b200.train.siamese_regions import get_siamese_net
    torch.manual_seed(0)
    base = torchvision.models.resnet152(weights=None)
    net = get_siamese_net(_P, pretrained=False, base_net=base)
    net.feature_reduc1[1].param.data.normal_(0, 0.01)       # a trained Shift is not zero
    # calibrate the BatchNorm statistics of the random trunk on one batch
    for m in net.features.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = None
    net.features.train()
    cal = (_images(16, 224, 224, torch.Generator().manual_seed(99)) - MEAN) / STD
    with torch.no_grad():
        net.features(cal.cuda())
    net.eval()
    net.invalidate_head_cache()
    return net


def _oracle_descriptors(net, fmaps):
    conv, shift, lin = net.classifier[0], net.feature_reduc1[1], net.feature_reduc1[2]
    return oracle.region_descriptor_forward(
        fmaps, conv.weight.detach().cpu().view(conv.out_channels, -1), conv.bias.detach().cpu(),
        shift.param.detach().cpu(), lin.weight.detach().cpu(), lin.bias.detach().cpu(), net.k, net.feature_size2d)


@pytest.mark.parametrize("h,w,n_query,n_db", [
    (224, 224, 32, 128),     # the config's image size: 7 x 7 map, one window
    (288, 352, 6, 20),       # a larger input: 9 x 11 map, 3 x 5 windows, k = 6 of them selected
])
def test_config0_resnet152_region_descriptors_and_top10(net, h, w, n_query, n_db):
    from instance_search_b200.train.siamese_regions import get_embeddings
    from instance_search_b200.search import DescriptorIndex
    gen = torch.Generator().manual_seed(h)
    db_img = _images(n_db, h, w, gen)
    q_img = (db_img[:n_query] + 0.05 * torch.randn(n_query, 3, h, w, generator=gen)).clamp(0, 1)
    norm = lambda t: (t - MEAN) / STD                                           # noqa: E731
    ref_set = [(im, "L%d" % i, "r%d" % i) for i, im in enumerate(norm(db_img))]
    test_set = [(im, "L%d" % i, "t%d" % i) for i, im in enumerate(norm(q_img))]

    # ---- GPU path: the drop-in entry points
    ref_emb = get_embeddings(net, ref_set, 0, 2048, batch_size=16)              # train/siamese_regions.py:26
    test_emb = get_embeddings(net, test_set, 0, 2048, batch_size=16)
    s, i = DescriptorIndex(ref_emb).search(test_emb, 10)

    # ---- CPU reference path on the same trunk output
    with torch.no_grad():
        fm_ref = torch.cat([net.features(torch.stack([t[0] for t in ref_set[a:a + 16]]).cuda()).cpu()
                            for a in range(0, n_db, 16)])
        fm_test = torch.cat([net.features(torch.stack([t[0] for t in test_set[a:a + 16]]).cuda()).cpu()
                             for a in range(0, n_query, 16)])
    assert fm_ref.shape[1:] == (2048, h // 32, w // 32)
    o_ref, _, oi_ref, _ = _oracle_descriptors(net, fm_ref)
    o_test, _, oi_test, _ = _oracle_descriptors(net, fm_test)
    check_descriptors(ref_emb, o_ref)
    check_descriptors(test_emb, o_test)
    # the windows the fused head selected == the reference's (forward_single returns them via cls_out only;
    # compare through the functional head)
    from instance_search_b200 import regions
    _, _, gi, gn = regions.region_descriptors(fm_test.cuda(), net._head(), net.k, net.feature_size2d)
    assert torch.equal(gi.cpu(), oi_test)

    sim = oracle.similarity(o_test, o_ref)                                      # test/siamese_regions_test.py:76
    o_s, o_i = sim.sort(dim=1, descending=True)
    o_s, o_i = o_s[:, :10], o_i[:, :10]
    # exact top-10: identical indices; a swap is tolerated only between two database images whose
    # reference scores differ by less than the descriptor parity (2 x DESC_L2_TOL = 2e-5)
    mism = (i.cpu() != o_i)
    if bool(mism.any()):
        rows, cols = mism.nonzero(as_tuple=True)
        gap = (sim[rows, i.cpu()[rows, cols]] - o_s[rows, cols]).abs()
        assert float(gap.max()) <= 2 * DESC_L2_TOL, "top-10 differs beyond descriptor parity: %g" % gap.max()
    assert int(mism.sum()) <= 2
    # scores: dots of two descriptors each within DESC_L2_TOL of the oracle's
    assert torch.allclose(s.cpu(), o_s, rtol=1e-5, atol=2 * DESC_L2_TOL)
    # retrieval sanity: precision@1 of the two paths is the same number
    from instance_search_b200.utils import metrics
    gsim = torch.mm(test_emb, ref_emb.t())
    assert metrics.precision1(gsim, test_set, ref_set)[1] == oracle.precision1(sim, test_set, ref_set)[1]
    record("config0", h=h, w=w, mismatches=int(mism.sum()), p_at_1=oracle.precision1(sim, test_set, ref_set)[0])
