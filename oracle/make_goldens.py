#!/usr/bin/env python
"""Pin the oracle against the reference itself and mint tests/golden/*.npz.

Run in the BUILD container only (needs /root/reference, which does not exist
on the GPU box):

    python oracle/make_goldens.py

What it does
  1. puts /root/reference/{.,model,utils,train,test} on sys.path (the reference
     is Python-2 style: implicit relative imports, model/siamese.py:6-7,
     utils/__init__.py:1-7);
  2. replaces the two legacy autograd wrappers torch >= 1.5 refuses to execute
     (``NormalizeL2`` / ``Shift`` modules, model/custom_modules.py:28-39,70-76)
     by modules with the same math and torch-0.1 ``keepdim`` semantics -- the
     ONLY modification; RegionDescriptorNet / DescriptorNet / metrics /
     instance_avg / get_lab_indicators run unmodified from the reference tree
     (instance_avg additionally sees its uint8 masks through a Tensor subclass
     whose ``1 - mask`` is a bool complement, torch-0.1 ByteTensor semantics);
  3. runs them on seeded inputs, asserts ``oracle.*`` reproduces every output
     BIT-FOR-BIT, and stores inputs + outputs as fixtures.
"""

import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


# ----------------------------------------------------------------------------
# the negative-selection block of train/siamese_regions.py, run from its source
# ----------------------------------------------------------------------------
def reference_negative_selection():
    """train/siamese_regions.py cannot be imported (its ``P`` reads a missing data file at
    import), so the lines of ``create_batch`` that choose the negative -- from
    ``ind_exl = lab_indicators[lab]`` to ``_, k = sims.max(0)`` (:106-126) -- are taken from the
    reference's source file AS TEXT and executed here.  The one line left out is the last of the
    else branch, ``im3 = train_set[k[0]][0]`` (:127): it indexes a 0-dim tensor, which torch >= 0.4
    refuses, and it only looks the chosen index up in the data set.  Returns
    f(similarities, lab_indicators, lab, i1, i2, epoch, train_epoch_switch) -> index or -1."""
    import textwrap
    src = open(os.path.join(REF, "train", "siamese_regions.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.strip() == "ind_exl = lab_indicators[lab]")
    last = next(i for i, l in enumerate(src) if i > first and l.strip() == "_, k = sims.max(0)")
    assert src[last + 1].strip() == "im3 = train_set[k[0]][0]", "reference source changed"
    block = textwrap.dedent("\n".join(src[first:last + 1]))
    code = compile(block, "train/siamese_regions.py:%d-%d" % (first + 1, last + 1), "exec")

    class _P(object):
        pass

    def run(similarities, lab_indicators, lab, i1, i2, epoch, train_epoch_switch):
        P = _P()
        P.train_epoch_switch = train_epoch_switch
        ns = {"similarities": similarities, "lab_indicators": lab_indicators, "lab": lab, "i1": i1, "i2": i2,
              "epoch": epoch, "P": P, "print": lambda *a, **k: None}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")     # uint8 mask indexing is deprecated, not changed
            exec(code, ns)
        return int(ns["k"]) if "k" in ns else -1
    return run


# ----------------------------------------------------------------------------
# import the reference with the two-class shim
# ----------------------------------------------------------------------------
def import_reference():
    for sub in ("", "model", "utils", "train", "test"):
        sys.path.insert(0, os.path.join(REF, sub))
    import custom_modules as ref_cm  # /root/reference/model/custom_modules.py

    class NormalizeL2(nn.Module):
        # same math as NormalizeL2Fun.forward (model/custom_modules.py:52-57)
        def forward(self, input):
            norm2 = input.pow(2).sum(1, keepdim=True).add_(1e-10)
            norm = norm2.pow(0.5)
            return input / norm.expand_as(input)

    class Shift(nn.Module):
        # same math/state as Shift + ShiftFun.forward (model/custom_modules.py:16-39)
        def __init__(self, n_features):
            super().__init__()
            self.param = nn.Parameter(torch.zeros(n_features))

        def forward(self, input):
            return input + self.param.view(1, -1).expand_as(input)

    ref_cm.NormalizeL2 = NormalizeL2
    ref_cm.Shift = Shift
    import siamese as ref_siamese  # /root/reference/model/siamese.py (unmodified)
    return ref_cm, ref_siamese


class TinyNet(nn.Module):
    """A net exposing the (features, feature_reduc, classifier) triple that
    extract_layers (model/nn_utils.py:56-58) accepts."""

    def __init__(self, cin, c, ncls, fsize):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(cin, c, 1), nn.ReLU())
        self.feature_reduc = nn.Sequential(nn.AvgPool2d(fsize))
        self.classifier = nn.Sequential(nn.Linear(c, ncls))


class Keepdim01(object):
    """torch-0.1 view of a similarity matrix: max(1)/kthvalue(k,1) keep the dim
    (what utils/metrics.py:11-17 was written against)."""

    def __init__(self, t):
        self.t = t

    def size(self, *a):
        return self.t.size(*a)

    def max(self, dim):
        return self.t.max(dim, keepdim=True)

    def kthvalue(self, k, dim):
        return self.t.kthvalue(k, dim, keepdim=True)

    def __getitem__(self, i):
        return self.t[i]


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def eq(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), "%s: oracle != reference (max |d| %g)" % (
        what, (a - b).abs().max().item())


def main():
    os.makedirs(GOLD, exist_ok=True)
    warnings.simplefilter("ignore")
    torch.set_num_threads(1)  # one summation order for the fixtures
    ref_cm, ref_siamese = import_reference()

    # ---------------------------------------------------------------- regions
    # RegionDescriptorNet (model/siamese.py:133-231) on a tiny trunk; maps of
    # 7x7 (1 window), 8x8 (4 windows < k), 9x12 (18 windows), 14x14 (64).
    C, NCLS, D, K, FS = 32, 5, 16, 6, (7, 7)
    g = torch.Generator().manual_seed(1234)
    net = ref_siamese.RegionDescriptorNet(TinyNet(8, C, NCLS, FS), K, D, FS)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.05))
    sd = net.state_dict()
    assert set(k for k in sd if not k.startswith("features")) == {
        "classifier.0.weight", "classifier.0.bias", "feature_reduc1.1.param",
        "feature_reduc1.2.weight", "feature_reduc1.2.bias"}, sd.keys()
    cls_w = sd["classifier.0.weight"].view(NCLS, C).clone()
    cls_b = sd["classifier.0.bias"].clone()
    shift = sd["feature_reduc1.1.param"].clone()
    lin_w = sd["feature_reduc1.2.weight"].clone()
    lin_b = sd["feature_reduc1.2.bias"].clone()
    arrs = dict(cls_w=cls_w, cls_b=cls_b, shift=shift, lin_w=lin_w, lin_b=lin_b,
                k=K, fsize=np.array(FS))
    for tag, (h, w) in {"7x7": (7, 7), "8x8": (8, 8), "9x12": (9, 12),
                        "14x14": (14, 14)}.items():
        xs, descs, clss, idxs = [], [], [], []
        for n in range(3):
            img = torch.randn(1, 8, h, w, generator=g)
            with torch.no_grad():
                net.train()   # train mode returns (desc, cls_out) per input (:225-229)
                d_tr, c_tr = net.forward_single(img)
                net.eval()    # eval mode returns the descriptor only (:231)
                d_ev = net(img)
                feat = net.features(img)
            eq(d_tr, d_ev, "train/eval desc")
            od, oc, oi = oracle.region_descriptor_forward_single(
                feat, cls_w, cls_b, shift, lin_w, lin_b, K, FS)
            eq(od, d_ev, "region desc " + tag)
            eq(oc, c_tr, "region cls_out " + tag)
            xs.append(feat), descs.append(d_ev), clss.append(c_tr)
            pad = torch.full((K,), -1, dtype=torch.int64)
            pad[:oi.numel()] = oi
            idxs.append(pad)
        arrs["x_" + tag] = torch.cat(xs)
        arrs["desc_" + tag] = torch.cat(descs)
        arrs["cls_out_" + tag] = torch.cat(clss)
        arrs["idx_" + tag] = torch.stack(idxs)
    save("region_tiny", **arrs)

    # batch semantics: the batched oracle == per-image reference calls
    xb = arrs["x_9x12"]
    bd, bc, bi, bn = oracle.region_descriptor_forward(xb, cls_w, cls_b, shift, lin_w,
                                                      lin_b, K, FS)
    eq(bd, arrs["desc_9x12"], "batched desc")
    eq(bc, arrs["cls_out_9x12"], "batched cls_out")

    # ----------------------------------------------------------- global head
    # DescriptorNet (model/siamese.py:92-130)
    g = torch.Generator().manual_seed(1235)
    dnet = ref_siamese.DescriptorNet(TinyNet(8, C, NCLS, FS), D, FS)
    with torch.no_grad():
        for p in dnet.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.05))
    dnet.eval()
    img = torch.randn(5, 8, 7, 7, generator=g)
    with torch.no_grad():
        feat = dnet.features(img)
        d_ref = dnet(img)
    dsd = dnet.state_dict()
    od = oracle.descriptor_forward(feat, dsd["feature_reduc1.1.param"],
                                   dsd["feature_reduc1.2.weight"],
                                   dsd["feature_reduc1.2.bias"])
    eq(od, d_ref, "DescriptorNet")
    save("descriptor_tiny", x=feat, shift=dsd["feature_reduc1.1.param"],
         lin_w=dsd["feature_reduc1.2.weight"], lin_b=dsd["feature_reduc1.2.bias"],
         desc=d_ref)

    # ---------------------------------------- ResNet-shaped head (C=512, 7x7)
    # resnet18 trunk through TuneClassifSub exactly as get_siamese_net does
    # (train/siamese_regions.py:157-168) -- checks extract_layers /
    # convolutionalize / get_feature_size glue with a real torchvision net.
    import torchvision.models as tvm
    torch.manual_seed(7)
    cnet = ref_siamese.TuneClassifSub(tvm.resnet18(weights=None), 11, FS)
    rnet = ref_siamese.RegionDescriptorNet(cnet, 6, 4, FS)
    g = torch.Generator().manual_seed(1236)
    with torch.no_grad():
        rnet.feature_reduc1[1].param.copy_(torch.randn(25088, generator=g) * 0.01)
    rnet.eval()
    img = torch.randn(2, 3, 320, 288, generator=g)
    rsd = rnet.state_dict()
    with torch.no_grad():
        feat = rnet.features(img)                       # [2, 512, 10, 9]
        d_ref = torch.cat([rnet(img[i:i + 1]) for i in range(2)])
        rnet.train()
        # BN in train mode would change the trunk; the head is what we pin
        rnet.features.eval()
        c_ref = torch.cat([rnet.forward_single(img[i:i + 1])[1] for i in range(2)])
    od, oc, oi, on = oracle.region_descriptor_forward(
        feat, rsd["classifier.0.weight"].view(11, 512), rsd["classifier.0.bias"],
        rsd["feature_reduc1.1.param"], rsd["feature_reduc1.2.weight"],
        rsd["feature_reduc1.2.bias"], 6, FS)
    eq(od, d_ref, "resnet18 region desc")
    eq(oc, c_ref, "resnet18 region cls_out")
    save("region_resnet18_head", x=feat,
         cls_w=rsd["classifier.0.weight"].view(11, 512), cls_b=rsd["classifier.0.bias"],
         shift=rsd["feature_reduc1.1.param"],
         lin_w=rsd["feature_reduc1.2.weight"],
         lin_b=rsd["feature_reduc1.2.bias"], desc=d_ref, cls_out=c_ref, idx=oi, k=6,
         fsize=np.array(FS))

    # ---------------------------------------------------------------- metrics
    import metrics as ref_metrics  # /root/reference/utils/metrics.py
    g = torch.Generator().manual_seed(1237)
    Q, N, NL = 12, 60, 7
    ref_lab = torch.randint(0, NL, (N,), generator=g)
    test_lab = torch.randint(0, NL + 1, (Q,), generator=g)  # label NL has no positives
    sim = torch.randn(Q, N, generator=g)
    ref_set = [(None, "L%d" % l, "r%d" % i) for i, l in enumerate(ref_lab.tolist())]
    test_set = [(None, "L%d" % l, "t%d" % i) for i, l in enumerate(test_lab.tolist())]
    out = {}
    for kth in (1, 2, 3):
        r = ref_metrics.precision1(Keepdim01(sim), test_set, ref_set, kth)
        o = oracle.precision1(sim, test_set, ref_set, kth)
        assert r[:3] == o[:3] and r[4] == o[4], "precision1"
        eq(r[3], o[3], "precision1 max_sim")
        rm = ref_metrics.mean_avg_precision(sim, test_set, ref_set, kth)
        om = oracle.mean_avg_precision(sim, test_set, ref_set, kth)
        assert rm == om, ("mAP", rm, om)
        aps = [ref_metrics.avg_precision(sim, i, test_set, ref_set, kth) for i in range(Q)]
        assert aps == [oracle.avg_precision(sim, i, test_set, ref_set, kth) for i in range(Q)]
        out["prec_kth%d" % kth] = np.array([r[0], r[1], r[2]])
        out["max_sim_kth%d" % kth] = r[3]
        out["max_label_kth%d" % kth] = np.array([int(s[1:]) for s in r[4]])
        out["map_kth%d" % kth] = rm
        out["ap_kth%d" % kth] = np.array([np.nan if a is None else a for a in aps])
    save("metrics_tiny", sim=sim, ref_lab=ref_lab, test_lab=test_lab, **out)

    # ----------------------------------------- label indicators + device rule
    import train_siamese as ref_ts  # /root/reference/utils/train_siamese.py
    ri = ref_ts.get_lab_indicators(ref_set, -1)
    oi_ = oracle.get_lab_indicators(ref_set)
    assert ri.keys() == oi_.keys()
    for lab in ri:
        eq(ri[lab], oi_[lab], "lab_indicator")

    class P(object):
        cuda_device, feature_dim, embeddings_cuda_size = 0, 2048, 2 ** 30

    class N_(object):
        feature_size = 2048
    for n, sm in ((16384, True), (16385, True), (131072, False), (131073, False)):
        r = ref_ts.embeddings_device_dim(P, N_, n, sim_matrix=sm)
        o = oracle.embeddings_device_dim(0, 2048, 2 ** 30, 2048, n, sim_matrix=sm)
        assert r == o, (n, sm, r, o)

    # ---------------------------------------------------------- instance_avg
    import instance_avg as ref_ia  # /root/reference/test/instance_avg.py

    class ByteMask01(torch.Tensor):
        # torch-0.1 ByteTensor masks: ``1 - mask`` is the complement and is a
        # valid index (test/instance_avg.py:26); torch 2.x wants bool there.
        def __rsub__(self, other):
            assert other == 1
            return self.as_subclass(torch.Tensor).eq(0)

    def lab_ind_01(dataset, device):
        d = ref_ts.get_lab_indicators(dataset, device)
        return {k: v.as_subclass(ByteMask01) for k, v in d.items()}
    ref_ia.get_lab_indicators = lab_ind_01
    g = torch.Generator().manual_seed(1238)
    E = oracle.normalize_l2(torch.randn(N, 32, generator=g))
    iout = {}
    for kk in (-1, 0, 2, 100):
        r, _ = ref_ia.instance_avg(-1, E.clone(), ref_set, None, kk)
        o = oracle.instance_avg(E.clone(), ref_set, kk)
        eq(r, o, "instance_avg k=%d" % kk)
        iout["out_k%d" % kk] = r
    save("instance_avg_tiny", emb=E, ref_lab=ref_lab, **iout)

    # --------------------------------------- search + mining (oracle-minted)
    # No reference function to call (the reference inlines torch.mm + sort);
    # these fixtures freeze the oracle's own answers for regression.
    g = torch.Generator().manual_seed(1239)
    q = oracle.normalize_l2(torch.randn(40, 64, generator=g))
    db = oracle.normalize_l2(torch.randn(1500, 64, generator=g))
    s, i = oracle.topk_search(q, db, 10)
    s64, i64 = oracle.topk_search_f64(q, db, 10)
    assert torch.equal(i, i64), "fixture has an fp32-ambiguous ranking; reseed"
    save("search_tiny", q=q, db=db, scores=s, idx=i)

    g = torch.Generator().manual_seed(1240)
    NM, DM = 96, 64
    lab = torch.arange(NM) // 6
    centers = torch.randn(16, DM, generator=g)
    E = centers[lab] + 0.7 * torch.randn(NM, DM, generator=g)
    E[5] = -E[0]          # couple (0, 5): the positive is the LEAST similar item,
    E = oracle.normalize_l2(E)  # so semi-hard mining excludes everything (-> -1)
    S = oracle.mining.all_pairs_similarities(E)
    couples = [(a, b) for a in range(NM) for b in range(NM)
               if a != b and lab[a] == lab[b]][::5]
    couples.append((0, 5))
    hard = oracle.select_negatives(S, lab, couples, semi_hard=False)
    semi = oracle.select_negatives(S, lab, couples, semi_hard=True)
    assert (semi == -1).any() and (semi >= 0).any()
    # the reference's own selection lines (executed from its source, see
    # reference_negative_selection) on every couple, both modes: pins oracle.select_negative
    ref_select = reference_negative_selection()
    ds = [(None, "L%d" % int(v), "im%d" % n) for n, v in enumerate(lab)]
    # torch-0.1 ByteTensor masks index and fill like today's bool masks (masked_fill_ refuses
    # uint8 since torch 1.2): the indicators go in as bool, the reference lines run unchanged
    ind = {lab_: m.bool() for lab_, m in oracle.mining.get_lab_indicators(ds).items()}
    for mode, want in ((False, hard), (True, semi)):
        epoch = 0 if mode else 5           # P.train_epoch_switch = 2 (train/siamese_regions_p.py)
        got = torch.tensor([ref_select(S, ind, "L%d" % int(lab[a]), a, b, epoch, 2) for a, b in couples])
        assert torch.equal(got, want), "oracle.select_negative differs from train/siamese_regions.py:106-126"
    save("mining_tiny", emb=E, lab=lab, couples=np.array(couples), sim=S,
         neg_hard=hard, neg_semi=semi)
    print("oracle == reference on every pinned function; goldens written")


if __name__ == "__main__":
    main()
