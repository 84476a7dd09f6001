"""CPU: the oracle reproduces the fixtures minted from the reference itself
(oracle/make_goldens.py).  Bit-exact: same torch build, one thread."""

import math

import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("tag", ["7x7", "8x8", "9x12", "14x14"])
def test_region_descriptor_matches_reference(tag):
    g = load_golden("region_tiny")
    fs = tuple(int(v) for v in g["fsize"])
    d, c, idx, nsel = oracle.region_descriptor_forward(
        g["x_" + tag], g["cls_w"], g["cls_b"], g["shift"], g["lin_w"], g["lin_b"],
        g["k"], fs)
    assert torch.equal(d, g["desc_" + tag])
    assert torch.equal(c, g["cls_out_" + tag])
    assert torch.equal(idx, g["idx_" + tag])
    # invariants the reference code implies
    assert torch.allclose(d.norm(dim=1), torch.ones(d.size(0)), atol=1e-6)
    hw = (g["x_" + tag].size(2) - fs[0] + 1) * (g["x_" + tag].size(3) - fs[1] + 1)
    assert int(nsel[0]) == min(hw, g["k"])
    assert c[:, :, int(nsel[0]):].abs().sum() == 0


def test_region_resnet18_head():
    g = load_golden("region_resnet18_head")
    d, c, idx, _ = oracle.region_descriptor_forward(
        g["x"], g["cls_w"], g["cls_b"], g["shift"], g["lin_w"], g["lin_b"], g["k"],
        tuple(int(v) for v in g["fsize"]))
    assert torch.equal(d, g["desc"]) and torch.equal(c, g["cls_out"])
    assert torch.equal(idx, g["idx"])


def test_descriptor_head():
    g = load_golden("descriptor_tiny")
    assert torch.equal(oracle.descriptor_forward(g["x"], g["shift"], g["lin_w"], g["lin_b"]),
                       g["desc"])


def _sets(g):
    ref_set = [(None, "L%d" % l, "r%d" % i) for i, l in enumerate(g["ref_lab"].tolist())]
    test_set = [(None, "L%d" % l, "t%d" % i) for i, l in enumerate(g["test_lab"].tolist())]
    return test_set, ref_set


@pytest.mark.parametrize("kth", [1, 2, 3])
def test_metrics(kth):
    g = load_golden("metrics_tiny")
    test_set, ref_set = _sets(g)
    p, c, t, ms, ml = oracle.precision1(g["sim"], test_set, ref_set, kth)
    assert [p, c, t] == g["prec_kth%d" % kth].tolist()
    assert torch.equal(ms, g["max_sim_kth%d" % kth]) and ms.shape == (g["sim"].size(0), 1)
    assert [int(s[1:]) for s in ml] == g["max_label_kth%d" % kth].tolist()
    assert oracle.mean_avg_precision(g["sim"], test_set, ref_set, kth) == g["map_kth%d" % kth]
    for i, a in enumerate(g["ap_kth%d" % kth].tolist()):
        o = oracle.avg_precision(g["sim"], i, test_set, ref_set, kth)
        assert (o is None and math.isnan(a)) or o == a


def test_instance_avg():
    g = load_golden("instance_avg_tiny")
    ref_set = [(None, "L%d" % l, "r%d" % i) for i, l in enumerate(g["ref_lab"].tolist())]
    for k in (-1, 0, 2, 100):
        assert torch.equal(oracle.instance_avg(g["emb"].clone(), ref_set, k), g["out_k%d" % k])


def test_search_and_mining_fixtures():
    g = load_golden("search_tiny")
    s, i = oracle.topk_search(g["q"], g["db"], 10)
    assert torch.equal(i, g["idx"]) and torch.equal(s, g["scores"])
    s64, i64 = oracle.topk_search_f64(g["q"], g["db"], 10)
    assert torch.equal(i64, g["idx"])
    assert torch.allclose(s64.float(), s, rtol=1e-5, atol=1e-7)
    m = load_golden("mining_tiny")
    S = oracle.mining.all_pairs_similarities(m["emb"])
    assert torch.equal(S, m["sim"])
    couples = [tuple(c) for c in m["couples"].tolist()]
    assert torch.equal(oracle.select_negatives(S, m["lab"], couples, False), m["neg_hard"])
    assert torch.equal(oracle.select_negatives(S, m["lab"], couples, True), m["neg_semi"])
    assert int(m["neg_semi"][-1]) == -1  # the all-excluded couple falls back


def test_normalize_eps_inside_sqrt():
    # model/custom_modules.py:54 adds eps to the squared norm
    z = oracle.normalize_l2(torch.zeros(2, 5))
    assert torch.equal(z, torch.zeros(2, 5))
    x = torch.full((1, 4), 1e-6)
    want = x / torch.sqrt((x * x).sum(1, keepdim=True) + 1e-10)
    assert torch.equal(oracle.normalize_l2(x), want)
    assert float(oracle.normalize_l2(x).norm()) < 0.3  # far from unit norm: eps dominates
