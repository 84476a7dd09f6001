"""Oracle (test infrastructure): forward passes of the reference's custom ops.

Follows /root/reference/model/custom_modules.py; torch-0.1 ``sum(1)`` kept the
reduced dimension, which ``keepdim=True`` restates.
"""

import torch


def normalize_l2(x, eps=1e-10):
    """Row L2 normalisation, eps INSIDE the square root.

    reference: model/custom_modules.py:52-57 (NormalizeL2Fun.forward)
        norm2 = input.pow(2).sum(1).add_(eps); norm = norm2.pow(0.5)
        output = input / norm.expand_as(input)
    """
    norm2 = x.pow(2).sum(1, keepdim=True).add_(eps)
    norm = norm2.pow(0.5)
    return x / norm.expand_as(x)


def shift(x, param):
    """y = x + param (broadcast over rows).

    reference: model/custom_modules.py:16-18 (ShiftFun.forward)
    """
    return x + param.view(1, -1).expand_as(x)


def triplet_loss(anchor, pos, neg, margin, size_average=True, normalized=True):
    """Forward of the triplet loss.

    reference: model/custom_modules.py:153-171 (TripletLossFun.forward)
    """
    if normalized:
        loss = (anchor * neg).sum(1)
        loss = loss - (anchor * pos).sum(1)
        loss = loss + margin
    else:
        sqdiff_pos = (anchor - pos).pow(2)
        sqdiff_neg = (anchor - neg).pow(2)
        loss = sqdiff_pos.sum(1) - sqdiff_neg.sum(1)
        loss = (loss + margin * 2) / 2
    loss = loss.clone()
    loss[loss.le(0)] = 0
    loss = loss.sum(0).view(1)
    if size_average:
        loss = loss / anchor.size(0)
    return loss
