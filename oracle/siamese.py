"""Oracle (test infrastructure): the descriptor heads of the reference nets.

Functional restatement of /root/reference/model/siamese.py on the trunk's
OUTPUT feature map (the ResNet trunk stays in PyTorch on both sides, so it is
not part of the path).  Weights are passed explicitly with the reference's
state-dict meaning:

    cls_w  [ncls, C]      classifier.0.weight (the convolutionalised FC,
                          model/nn_utils.py:26-39, viewed as [ncls, C, 1, 1])
    cls_b  [ncls]         classifier.0.bias
    shift  [C*fh*fw]      feature_reduc1.1.param
    lin_w  [D, C*fh*fw]   feature_reduc1.2.weight
    lin_b  [D]            feature_reduc1.2.bias
"""

import torch
import torch.nn.functional as F

from .custom_modules import normalize_l2, shift as shift_fn


def region_descriptor_forward_single(x, cls_w, cls_b, shift, lin_w, lin_b, k,
                                     feature_size2d):
    """One image. x: [1, C, H, W] trunk feature map.

    reference: model/siamese.py:185-223 (RegionDescriptorNet.forward_single,
    everything after ``x = self.features(x)``), followed literally.
    Returns (desc [1, D], cls_out [1, ncls, k], flat_idx [k'] int64).
    """
    assert x.size(0) == 1, "the reference path is batch 1 (model/siamese.py:184)"
    fh, fw = feature_size2d
    ncls, C = cls_w.shape
    D = lin_w.size(0)
    # :187 feature_reduc = AvgPool2d(feature_size2d, stride=1)   (:164-166)
    c = F.avg_pool2d(x, (fh, fw), stride=1)
    # :188 classifier = 1x1 conv copy of the FC                   (:167-175)
    c = F.conv2d(c, cls_w.view(ncls, C, 1, 1), cls_b)
    # :191-194
    c_maxv, _ = c.max(1)
    c_maxv = c_maxv.reshape(-1)
    kk = min(c_maxv.size(0), k)
    _, flat_idx = c_maxv.topk(kk)

    # :199-203
    def feature_idx(fi):
        cls_idx = fi // c.size(3), fi % c.size(3)
        return (cls_idx[0], cls_idx[0] + fh, cls_idx[1], cls_idx[1] + fw)
    top_idx = [feature_idx(int(i)) for i in flat_idx]
    # :205-208
    acc = torch.zeros(c.size(0), D, dtype=x.dtype)
    cls_out = torch.zeros(c.size(0), c.size(1), k, dtype=x.dtype)
    # :214-220
    i = 0
    for x1, x2, y1, y2 in top_idx:
        cls_out[:, :, i] = c[:, :, x1, y1]
        i += 1
        region = x[:, :, x1:x2, y1:y2].contiguous().view(x.size(0), -1)
        region = normalize_l2(region)               # feature_reduc1.0 (:178)
        region = shift_fn(region, shift)            # feature_reduc1.1 (:179)
        region = F.linear(region, lin_w, lin_b)     # feature_reduc1.2 (:180)
        acc = acc + region
    # :222
    out = normalize_l2(acc)
    return out, cls_out, flat_idx


def region_descriptor_forward(x, cls_w, cls_b, shift, lin_w, lin_b, k,
                              feature_size2d):
    """Batch = loop of forward_single over images, as get_embeddings does.

    reference: train/siamese_regions.py:31-38 (one image at a time).
    Returns (desc [B, D], cls_out [B, ncls, k], idx [B, k] int64 padded -1,
    nsel [B]).
    """
    B = x.size(0)
    descs, clss = [], []
    idx = torch.full((B, k), -1, dtype=torch.int64)
    nsel = torch.zeros(B, dtype=torch.int64)
    for b in range(B):
        d, c, fi = region_descriptor_forward_single(
            x[b:b + 1], cls_w, cls_b, shift, lin_w, lin_b, k, feature_size2d)
        descs.append(d)
        clss.append(c)
        idx[b, :fi.numel()] = fi
        nsel[b] = fi.numel()
    return torch.cat(descs, 0), torch.cat(clss, 0), idx, nsel


def descriptor_forward(x, shift, lin_w, lin_b):
    """Global descriptor head. x: [B, C, fh, fw] trunk feature map.

    reference: model/siamese.py:117-122 (DescriptorNet.forward_single after
    the trunk): flatten -> NormalizeL2 -> Shift -> Linear -> NormalizeL2.
    """
    x = x.reshape(x.size(0), -1)
    x = normalize_l2(x)
    x = shift_fn(x, shift)
    x = F.linear(x, lin_w, lin_b)
    return normalize_l2(x)


def classif_regions_embedding(x, cls_w, cls_b, feature_size2d):
    """Embedding of the classification track's sub-window net (TuneClassifSub): the class scores
    at the window with the highest maximal activation, L2-normalised.  x: [1, C, H, W] trunk
    feature map of ONE image.

    reference: train/classif_regions.py:107-132 (``get_embeddings``' per-image body after
    ``net(x)[0]``, i.e. after TuneClassifSub.forward_single, model/siamese.py:82-86), with the
    torch-0.1 semantics its indexing relies on: ``max(dim)`` keeps the reduced dimension.
    Not pinned against the reference run in the container (the body cannot execute on torch >= 1
    because of those semantics); pinned by construction to the same avg-pool / 1x1-conv arithmetic
    as region_descriptor_forward_single, which is.
    Returns (embedding [1, ncls], flat window index).
    """
    assert x.size(0) == 1
    fh, fw = feature_size2d
    ncls, C = cls_w.shape
    out = F.conv2d(F.avg_pool2d(x, (fh, fw), stride=1), cls_w.view(ncls, C, 1, 1), cls_b)   # net(x)[0]
    max_pred, _ = out.max(1, keepdim=True)              # :119  [1, 1, H', W']
    max_pred1, max_i1 = max_pred.max(2, keepdim=True)   # :120  best row of every column
    _, max_i2 = max_pred1.max(3, keepdim=True)          # :121  best column
    i2 = int(max_i2.view(-1)[0])
    i1 = int(max_i1.view(-1)[i2])
    vec = out[:, :, i1, i2]                              # :126
    return normalize_l2(vec), i1 * out.size(3) + i2     # :127
