"""CPU: host-side logic of the drop-in modules (no kernels run)."""

import math

import pytest
import torch
import torch.nn as nn

import oracle
from conftest import load_golden


class ToyNet(nn.Module):
    """A pre-split net in the form model/nn_utils.py:57-58 accepts."""

    def __init__(self, C=8, ncls=5):
        super(ToyNet, self).__init__()
        self.features = nn.Sequential(nn.Conv2d(3, C, 3, stride=2, padding=1), nn.BatchNorm2d(C), nn.ReLU())
        self.feature_reduc = nn.Sequential(nn.AdaptiveAvgPool2d(1))
        self.classifier = nn.Sequential(nn.Linear(C, ncls))


def test_region_descriptor_net_construction_and_state_dict_keys():
    from instance_search_b200.model.siamese import RegionDescriptorNet, DescriptorNet
    net = RegionDescriptorNet(ToyNet(), k=6, feature_dim=16, feature_size2d=(7, 7))
    keys = set(net.state_dict().keys())
    # the keys a reference checkpoint carries outside the trunk
    for k in ["classifier.0.weight", "classifier.0.bias", "feature_reduc1.1.param",
              "feature_reduc1.2.weight", "feature_reduc1.2.bias"]:
        assert k in keys
    assert net.classifier[0].weight.shape == (5, 8, 1, 1)          # convolutionalised FC
    assert net.feature_reduc1[2].weight.shape == (16, 8 * 49)       # in_features = C * fh * fw
    assert isinstance(net.feature_reduc[0], nn.AvgPool2d) and net.feature_reduc[0].stride == 1
    assert net.k == 6 and net.feature_size == 16 and net.feature_size2d == (7, 7)
    assert float(net.feature_reduc1[1].param.abs().sum()) == 0.0    # Shift starts at zero
    assert not any(p.requires_grad for p in net.features.parameters())   # untrained=-1 freezes
    d = DescriptorNet(ToyNet(), feature_dim=0, feature_size2d=(7, 7))
    assert d.feature_size == 5                                        # falls back to classifier size
    assert "feature_reduc1.1.param" in d.state_dict()


def test_convolutionalize_copies_weights_and_checks_sizes():
    from instance_search_b200.model.nn_utils import convolutionalize
    fc = nn.Linear(8 * 4, 3)
    conv = convolutionalize(fc, (2, 2))
    x = torch.randn(5, 8, 2, 2)
    assert torch.allclose(conv(x).view(5, 3), fc(x.view(5, -1)), atol=1e-6)
    with pytest.raises(ValueError):
        convolutionalize(nn.Linear(10, 3), (2, 2))


def test_set_net_train_keeps_batchnorm_frozen():
    from instance_search_b200.model.nn_utils import set_net_train
    from instance_search_b200.model.siamese import RegionDescriptorNet
    net = RegionDescriptorNet(ToyNet(), 6, 16, (7, 7))
    set_net_train(net, True)
    assert net.training and not net.features[1].training
    set_net_train(net, True, bn_train=True)
    assert net.features[1].training
    set_net_train(net, False)
    assert not net.training


def test_modules_refuse_cpu_tensors():
    from instance_search_b200 import IsbError
    from instance_search_b200.model.custom_modules import NormalizeL2, Shift, TripletLoss
    with pytest.raises(IsbError):
        NormalizeL2()(torch.randn(2, 8))
    with pytest.raises(IsbError):
        Shift(8)(torch.randn(2, 8))
    with pytest.raises(IsbError):
        TripletLoss(0.1)(torch.randn(2, 8), torch.randn(2, 8), torch.randn(2, 8))


def _ranks_cpu(sim_row, cols):
    s = sim_row
    return sorted(int(((s > s[c]) | ((s == s[c]) & (torch.arange(s.numel()) < c))).sum()) for c in cols)


@pytest.mark.parametrize("kth", [1, 2, 3])
def test_ap_from_positive_ranks_is_bit_identical_to_the_full_walk(kth):
    from instance_search_b200.utils.metrics import _ap_from_ranks
    g = load_golden("metrics_tiny")
    ref_lab, test_lab = g["ref_lab"].tolist(), g["test_lab"].tolist()
    ref_set = [(None, "L%d" % l, "") for l in ref_lab]
    test_set = [(None, "L%d" % l, "") for l in test_lab]
    for i in range(g["sim"].size(0)):
        cols = [j for j, l in enumerate(ref_lab) if l == test_lab[i]]
        mine = _ap_from_ranks(_ranks_cpu(g["sim"][i], cols), len(cols), kth)
        want = oracle.avg_precision(g["sim"], i, test_set, ref_set, kth)
        assert mine == want or (mine is None and want is None)
        a = g["ap_kth%d" % kth][i].item()
        assert (mine is None and math.isnan(a)) or mine == a     # and to the reference's own output


def test_get_embeddings_batches_equal_sized_images_only():
    from instance_search_b200.train import siamese_regions as sr
    calls = []

    class Net(object):
        def __call__(self, x):
            calls.append(tuple(x.shape))
            return x.flatten(1)[:, :4].cuda() if torch.cuda.is_available() else x.flatten(1)[:, :4]
    if not torch.cuda.is_available():
        pytest.skip("output buffer lives on the GPU")
    ds = [(torch.randn(3, 4, 4), "a", ""), (torch.randn(3, 4, 4), "a", ""), (torch.randn(3, 5, 4), "b", ""),
          (torch.randn(3, 4, 4), "b", "")]
    out = sr.get_embeddings(Net(), ds, -1, 4, batch_size=8)
    assert calls == [(2, 3, 4, 4), (1, 3, 5, 4), (1, 3, 4, 4)] and out.shape == (4, 4) and not out.is_cuda


def test_label_ids_and_placement_rule():
    from instance_search_b200 import mining
    ids, labels = mining.label_ids([(None, "x", ""), (None, "y", ""), (None, "x", "")])
    assert ids.tolist() == [0, 1, 0] and labels == ["x", "y"]

    class P(object):
        cuda_device, feature_dim, embeddings_cuda_size = 0, 0, 2 ** 30

    class Net(object):
        feature_size = 512
    assert mining.embeddings_device_dim(P, Net, 1000) == (0, 512)         # feature_dim <= 0 -> net's
    assert mining.embeddings_device_dim(P, Net, 2 ** 20) == (-1, 512)     # N*D*4 > 2^30 -> host
    assert oracle.embeddings_device_dim(0, 0, 2 ** 30, 512, 2 ** 20) == (-1, 512)


def test_uncertified_image_lists_are_merged():
    # the two certificates of the region head report [count, image, image, ...] each; the host
    # redoes the union of the listed images (regions.region_descriptors)
    from instance_search_b200.regions import uncertified_images
    B = 6
    n_unc = torch.zeros(2, 1 + B, dtype=torch.int32)
    assert uncertified_images(n_unc) == []
    n_unc[0, :3] = torch.tensor([2, 4, 1])       # screen certificate: images 4 and 1
    n_unc[1, :2] = torch.tensor([1, 4])          # selection certificate: image 4 again
    n_unc[1, 2:] = 5                             # stale entries beyond the count are ignored
    assert uncertified_images(n_unc) == [1, 4]
