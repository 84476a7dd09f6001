"""Drop-in for the descriptor nets of the reference's ``model/siamese.py``.

    TuneClassif            model/siamese.py:10-54   (construction + plain forward only)
    TuneClassifSub         model/siamese.py:57-89   (what get_siamese_net wraps, train/siamese_regions.py:161)
    DescriptorNet          model/siamese.py:92-130
    RegionDescriptorNet    model/siamese.py:133-231

Same constructor arguments, attributes (``k``, ``feature_size``,
``feature_size2d``, ``features``, ``feature_reduc``, ``classifier``,
``feature_reduc1``, ``feature_reduc2``), state-dict keys
(``classifier.0.{weight,bias}``, ``feature_reduc1.1.param``,
``feature_reduc1.2.{weight,bias}``) and forward contract, so the reference's
checkpoints load and its callers run unchanged.  The trunk (``features``) stays
in PyTorch.  Everything after it runs, without autograd, as the fused CUDA head
of ``instance_search_b200.regions`` for the WHOLE batch at once -- the result
equals the reference's batch-1 forward (:184) looped over the images.  When a
gradient is required (training), the same fused head runs under autograd
(``regions_autograd.RegionHeadFunction``: split-operand tcgen05 GEMMs for the weight /
operand gradients, a gather kernel for the crop gradients), still batched.  Nets whose
classifier is not AvgPool + one 1x1 conv (AlexNet's two-FC classifier) are composed per
image as the reference composes them (:185-223) from this package's NormalizeL2 / Shift
and torch's conv / linear autograd (``_forward_single_composed``, also the checker of the
fused backward in the tests).
"""

import torch
import torch.nn as nn

from .. import regions, regions_autograd
from .custom_modules import NormalizeL2, Shift
from .nn_utils import convolutionalize, extract_layers, get_feature_size, set_untrained_blocks


class _HeadCache(object):
    """Tensor-core-ready copies of the head parameters, rebuilt when any of them
    changes.  In-place updates bump ``_version`` and re-assignment changes data_ptr, but
    writes through ``.data`` (``p.data.copy_()``, the reference's way of loading weights,
    and Shift.reset_parameters) do neither: the owning module therefore also calls
    ``invalidate()`` from ``train()``, ``_load_from_state_dict`` and ``_apply`` -- every
    train -> eval transition rebuilds the copies."""

    def __init__(self):
        self.key, self.hw = None, None

    def invalidate(self):
        self.key, self.hw = None, None

    def get(self, params, terms, build):
        key = (terms,) + tuple((None if p is None else (p.data_ptr(), p._version)) for p in params)
        if key != self.key:
            self.hw = build()
            self.key = key
        return self.hw


class _CachedHeadMixin(object):
    """Invalidation hooks of the cached head operands (see _HeadCache)."""

    def invalidate_head_cache(self):
        """Call after writing head parameters through ``.data`` while in eval mode."""
        cache = self.__dict__.get("_cache")
        if cache is not None:
            cache.invalidate()

    def train(self, mode=True):
        self.invalidate_head_cache()
        return super(_CachedHeadMixin, self).train(mode)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_head_cache()
        return super(_CachedHeadMixin, self)._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_head_cache()
        return super(_CachedHeadMixin, self).load_state_dict(*args, **kwargs)


class TuneClassif(nn.Module):
    """Classifier fine-tuned from a pretrained net: trunk / spatial reduction / classifier
    with the last FC resized to ``num_classes``.  reference: model/siamese.py:10-54.
    Only what the region-descriptor path needs of it -- it is the net that
    ``get_siamese_net`` hands to RegionDescriptorNet (train/siamese_regions.py:161-164) --
    runs in PyTorch; the classification track that trains it is out of scope."""

    def __init__(self, net, num_classes, untrained=-1, reduc=True):
        super(TuneClassif, self).__init__()
        self.features, self.feature_reduc, self.classifier = extract_layers(net)
        set_untrained_blocks([self.features, self.classifier], untrained)
        names = list(self.classifier._modules.keys())
        last = self.classifier._modules[names[-1]]
        if not isinstance(last, nn.Linear) or last.out_features != num_classes:      # :27-31
            self.classifier._modules[names[-1]] = nn.Linear(last.in_features, num_classes)
        self.feature_size = num_classes
        if not reduc:                                                                # :34-46
            factor = 1
            for m in self.feature_reduc:
                ks = m.kernel_size
                factor *= ks[0] * ks[1] if isinstance(ks, (tuple, list)) else ks * ks
            first = self.classifier._modules[names[0]]
            self.classifier._modules[names[0]] = nn.Linear(first.in_features * factor, first.out_features)
            self.feature_reduc = nn.Sequential()

    def forward(self, x):
        x = self.feature_reduc(self.features(x))
        return self.classifier(x.view(x.size(0), -1))


class TuneClassifSub(TuneClassif):
    """TuneClassif whose FC layers are convolutionalised so every sub-window of the map is
    classified.  reference: model/siamese.py:57-89"""

    def __init__(self, net, num_classes, feature_size2d, untrained=-1):
        super(TuneClassifSub, self).__init__(net, num_classes, untrained, reduc=True)
        reduc_count = sum(1 for _ in self.feature_reduc)
        if reduc_count > 0:
            self.feature_reduc = nn.Sequential(nn.AvgPool2d(tuple(feature_size2d), stride=1))
        count = 0
        for name, module in self.classifier._modules.items():
            if isinstance(module, nn.Linear):
                size2d = (1, 1) if (reduc_count > 0 or count > 0) else tuple(feature_size2d)
                self.classifier._modules[name] = convolutionalize(module, size2d)
                count += 1

    def forward_single(self, x):
        return self.classifier(self.feature_reduc(self.features(x)))

    def forward(self, *scales):
        return [self.forward_single(x) for x in scales]


class DescriptorNet(_CachedHeadMixin, nn.Module):
    """Global descriptor: trunk -> flatten -> L2 -> Shift -> Linear -> L2.
    reference: model/siamese.py:92-130"""

    projection_terms = 3   # 3: fp32-grade split-operand projection; 1: plain bf16

    def __init__(self, net, feature_dim, feature_size2d, untrained=-1):
        super(DescriptorNet, self).__init__()
        self.features, _, classifier = extract_layers(net)
        set_untrained_blocks([self.features], untrained)
        factor = feature_size2d[0] * feature_size2d[1]
        in_features = get_feature_size(self.features, factor)
        if feature_dim <= 0:
            self.feature_size = get_feature_size(classifier)
        else:
            self.feature_size = feature_dim
        self.feature_reduc1 = nn.Sequential(
            NormalizeL2(),
            Shift(in_features),
            nn.Linear(in_features, self.feature_size)
        )
        self.feature_reduc2 = NormalizeL2()
        self._cache = _HeadCache()

    def _head(self):
        shift, lin = self.feature_reduc1[1], self.feature_reduc1[2]
        return self._cache.get(
            (shift.param, lin.weight, lin.bias), self.projection_terms,
            lambda: regions.HeadWeights(None, None, shift.param.data, lin.weight.data,
                                        None if lin.bias is None else lin.bias.data,
                                        terms=self.projection_terms))

    def forward_single(self, x):
        x = self.features(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.feature_reduc1.parameters())):
            x = x.reshape(x.size(0), -1)                 # :119-121
            x = self.feature_reduc1(x)
            return self.feature_reduc2(x)
        return regions.global_descriptors(x.detach(), self._head())

    def forward(self, x1, x2=None, x3=None):
        if self.training and x3 is not None:
            return self.forward_single(x1), self.forward_single(x2), self.forward_single(x3)
        elif self.training:
            return self.forward_single(x1), self.forward_single(x2)
        else:
            return self.forward_single(x1)


class RegionDescriptorNet(_CachedHeadMixin, nn.Module):
    """Top-k region descriptor. reference: model/siamese.py:133-231"""

    projection_terms = 3

    def __init__(self, net, k, feature_dim, feature_size2d, untrained=-1):
        super(RegionDescriptorNet, self).__init__()
        self.k = k
        self.feature_size2d = tuple(feature_size2d)
        self.features, self.feature_reduc, self.classifier = extract_layers(net)
        factor = feature_size2d[0] * feature_size2d[1]
        in_features = get_feature_size(self.features, factor)
        if feature_dim <= 0:
            self.feature_size = get_feature_size(self.classifier)
        else:
            self.feature_size = feature_dim
        reduc_count = sum(1 for _ in self.feature_reduc)
        if reduc_count > 0:
            # ResNet-like: sliding mean over feature_size2d windows, stride 1 (:161-166)
            self.feature_reduc = nn.Sequential(nn.AvgPool2d(self.feature_size2d, stride=1))
        count = 0
        for name, module in self.classifier._modules.items():   # :167-175
            if isinstance(module, nn.Linear):
                size2d = self.feature_size2d
                if reduc_count > 0 or count > 0:
                    size2d = (1, 1)
                self.classifier._modules[name] = convolutionalize(module, size2d)
                count += 1
        set_untrained_blocks([self.features, self.classifier], untrained)
        self.feature_reduc1 = nn.Sequential(
            NormalizeL2(),
            Shift(in_features),
            nn.Linear(in_features, self.feature_size)
        )
        self.feature_reduc2 = NormalizeL2()
        self._cache = _HeadCache()

    # ---- fused CUDA head -------------------------------------------------------
    def _fusable(self):
        """AvgPool(window, stride 1) + ONE 1x1 conv: the ResNet case of the reference."""
        mods = list(self.classifier)
        return (len(list(self.feature_reduc)) == 1 and isinstance(self.feature_reduc[0], nn.AvgPool2d) and
                len(mods) == 1 and isinstance(mods[0], nn.Conv2d) and tuple(mods[0].kernel_size) == (1, 1) and
                mods[0].bias is not None)

    def _head(self):
        conv = self.classifier[0]
        shift, lin = self.feature_reduc1[1], self.feature_reduc1[2]
        return self._cache.get(
            (conv.weight, conv.bias, shift.param, lin.weight, lin.bias), self.projection_terms,
            lambda: regions.HeadWeights(conv.weight.data, conv.bias.data, shift.param.data, lin.weight.data,
                                        None if lin.bias is None else lin.bias.data,
                                        terms=self.projection_terms))

    def _needs_grad(self, x):
        return torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.classifier.parameters()) or
            any(p.requires_grad for p in self.feature_reduc1.parameters()))

    # ---- reference composition, one image (:185-223), autograd-capable ---------
    def _forward_single_composed(self, x):
        c = self.feature_reduc(x)
        c = self.classifier(c)
        c_maxv, _ = c.max(1)
        c_maxv = c_maxv.reshape(-1)
        k = min(c_maxv.size(0), self.k)
        _, flat_idx = c_maxv.topk(k)
        fh, fw = self.feature_size2d
        acc = x.new_zeros(c.size(0), self.feature_size)
        cls_cols = []
        for fi in flat_idx.tolist():
            r, col = fi // c.size(3), fi % c.size(3)
            cls_cols.append(c[:, :, r, col])
            region = x[:, :, r:r + fh, col:col + fw].contiguous().view(x.size(0), -1)
            acc = acc + self.feature_reduc1(region)
        cls_out = torch.stack(cls_cols, 2)
        if k < self.k:   # zero-padded to k slots (:207-208)
            cls_out = torch.cat([cls_out, cls_out.new_zeros(c.size(0), c.size(1), self.k - k)], 2)
        return self.feature_reduc2(acc), cls_out

    def forward_single(self, x, want_cls_out=True):
        """x: [B, 3, h, w] images -> (desc [B, D], cls_out [B, ncls, k]); every image is
        treated as the reference's batch-1 call (model/siamese.py:184).  want_cls_out=False
        (what the eval-mode forward needs, :231) skips the logits output: cls_out is None."""
        x = self.features(x)
        if not self._fusable():
            outs = [self._forward_single_composed(x[b:b + 1]) for b in range(x.size(0))]
            return torch.cat([d for d, _ in outs], 0), torch.cat([c for _, c in outs], 0)
        if self._needs_grad(x):
            # training: the same fused head under autograd (regions_autograd), whole batch at once
            return regions_autograd.region_head(x, self.classifier[0], self.feature_reduc1[1],
                                                self.feature_reduc1[2], self._head(), self.k, self.feature_size2d)
        desc, cls_out, _, _ = regions.region_descriptors(x.detach(), self._head(), self.k, self.feature_size2d,
                                                          want_cls_out=want_cls_out)
        return desc, cls_out

    def forward(self, x1, x2=None, x3=None):
        if self.training and x3 is not None:
            return self.forward_single(x1), self.forward_single(x2), self.forward_single(x3)
        elif self.training:
            return self.forward_single(x1), self.forward_single(x2)
        else:
            return self.forward_single(x1, want_cls_out=False)[0]
