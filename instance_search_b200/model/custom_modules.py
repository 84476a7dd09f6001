"""Drop-in for the reference's ``model/custom_modules.py`` on the B200 path.

Same class names, constructor arguments, parameter names (``Shift.param``) and
forward/backward math as the reference; the arithmetic runs in libisb.so
(CUDA tensors only -- a CPU tensor raises, there is no fallback).

    NormalizeL2 / NormalizeL2Fun   model/custom_modules.py:46-76
    Shift / ShiftFun               model/custom_modules.py:11-39
    TripletLoss / TripletLossFun   model/custom_modules.py:140-215

The reference's Functions are legacy instance-style autograd Functions (torch
0.1); these are the static-method form modern torch requires, so
``NormalizeL2Fun.apply(x)`` replaces ``NormalizeL2Fun()(x)``.  MetricLoss is
not on the path (no script of the reference uses it) and is not provided.
"""

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.nn.parameter import Parameter

from .. import ops


def _as_rows(x):
    if x.dim() != 2:
        raise ops.IsbError("expected a [rows, features] tensor, got %s" % (tuple(x.shape),))
    return x


class ShiftFun(Function):
    """y = x + param (broadcast over rows). reference: model/custom_modules.py:11-25"""

    @staticmethod
    def forward(ctx, input, param):
        ctx.save_for_backward(input, param)
        return ops.shift_rows(_as_rows(input), param)

    @staticmethod
    def backward(ctx, grad_output):
        # :20-25  grad_input = grad_output.clone(); grad_param = grad_output^T . 1
        grad_output = grad_output.contiguous()
        return grad_output.clone(), ops.col_sums(grad_output)


class Shift(nn.Module):
    """reference: model/custom_modules.py:28-39 (parameter name ``param``, zero init)"""

    def __init__(self, n_features):
        super(Shift, self).__init__()
        self.param = Parameter(torch.Tensor(n_features))
        self.reset_parameters()

    def reset_parameters(self):
        self.param.data.fill_(0)

    def forward(self, input):
        return ShiftFun.apply(input, self.param)


class NormalizeL2Fun(Function):
    """Row L2 normalisation, eps INSIDE the sqrt. reference: model/custom_modules.py:46-67"""

    @staticmethod
    def forward(ctx, input, eps=1e-10):
        ctx.save_for_backward(input)
        ctx.eps = eps
        return ops.l2norm_rows(_as_rows(input), eps)

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return ops.l2norm_rows_backward(input, grad_output.contiguous(), ctx.eps), None


class NormalizeL2(nn.Module):
    """reference: model/custom_modules.py:70-76"""

    def __init__(self):
        super(NormalizeL2, self).__init__()

    def forward(self, input):
        return NormalizeL2Fun.apply(input)


class TripletLossFun(Function):
    """reference: model/custom_modules.py:140-203"""

    @staticmethod
    def forward(ctx, anchor, pos, neg, margin, size_average=True, normalized=True):
        loss, clamp = ops.triplet_loss_forward(anchor, pos, neg, margin, size_average, normalized)
        ctx.save_for_backward(anchor, pos, neg, clamp)
        ctx.size_average, ctx.normalized = size_average, normalized
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        anchor, pos, neg, clamp = ctx.saved_tensors
        ga, gp, gn = ops.triplet_loss_backward(anchor, pos, neg, clamp, grad_output, ctx.size_average,
                                               ctx.normalized)
        return ga, gp, gn, None, None, None


class TripletLoss(nn.Module):
    """reference: model/custom_modules.py:206-215"""

    def __init__(self, margin, size_average=True, normalized=True):
        super(TripletLoss, self).__init__()
        self.size_average = size_average
        self.margin = margin
        self.normalized = normalized

    def forward(self, anchor, pos, neg):
        return TripletLossFun.apply(anchor, pos, neg, self.margin, self.size_average, self.normalized)
