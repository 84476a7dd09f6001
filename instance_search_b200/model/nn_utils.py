"""Construction-time helpers the descriptor nets need (host logic, no kernels).

Behavioural mirror of the reference's ``model/nn_utils.py`` for the functions
``model/siamese.py`` calls: splitting a torchvision net into trunk / spatial
reduction / classifier, turning the classifier's FC into a 1x1 convolution,
inferring the feature size, freezing blocks, and the train/eval switch that
keeps BatchNorm frozen.  torchvision is imported lazily and only for its ResNet
block types; a net that already carries ``features`` / ``feature_reduc`` /
``classifier`` attributes needs no torchvision at all.
"""

import torch.nn as nn


def _resnet_types():
    try:
        import torchvision.models as models
        return models.ResNet, models.resnet.Bottleneck, models.resnet.BasicBlock
    except Exception:  # torchvision absent: only pre-split nets are supported
        return (), (), ()


def set_untrained_blocks(containers, n):
    """First n parameterised modules frozen; n < 0 freezes everything.
    reference: model/nn_utils.py:6-23"""
    for container in containers:
        for m in container:
            for p in m.parameters():
                p.requires_grad = n >= 0
    count = 0
    for seq in containers:
        for m in seq:
            if count >= n:
                break
            params = list(m.parameters())
            if not params:
                continue
            for p in params:
                p.requires_grad = False
            count += 1


def convolutionalize(fc, in_size2d):
    """nn.Linear -> nn.Conv2d with the same weights. reference: model/nn_utils.py:26-39"""
    area = in_size2d[0] * in_size2d[1]
    if fc.in_features % area != 0:
        raise ValueError('FC in_feature size {0} is not divisible by in_size2d {1}'.format(
            fc.in_features, in_size2d))
    in_channels = fc.in_features // area
    conv = nn.Conv2d(in_channels, fc.out_features, in_size2d, bias=fc.bias is not None)
    conv.weight.data.copy_(fc.weight.data.view(fc.out_features, in_channels, *in_size2d))
    if fc.bias is not None:
        conv.bias.data.copy_(fc.bias.data)
    return conv


def get_feature_size(seq, factor=1, default=-1):
    """Output size of the last size-defining module. reference: model/nn_utils.py:42-53"""
    _, Bottleneck, BasicBlock = _resnet_types()
    feature_size = default
    for module in seq:
        if Bottleneck and isinstance(module, Bottleneck):
            feature_size = module.conv3.out_channels * factor
        if BasicBlock and isinstance(module, BasicBlock):
            feature_size = module.conv2.out_channels * factor
        if isinstance(module, nn.Conv2d):
            feature_size = module.out_channels * factor
        if isinstance(module, nn.Linear):
            feature_size = module.out_features
    return feature_size


def extract_layers(net):
    """(features, feature_reduc, classifier) of a net. reference: model/nn_utils.py:56-71"""
    if hasattr(net, 'features') and hasattr(net, 'feature_reduc') and hasattr(net, 'classifier'):
        return net.features, net.feature_reduc, net.classifier
    ResNet, _, _ = _resnet_types()
    if ResNet and isinstance(net, ResNet):
        layers = [net.conv1, net.bn1, net.relu, net.maxpool]
        for stage in (net.layer1, net.layer2, net.layer3, net.layer4):
            layers.extend(stage)
        return nn.Sequential(*layers), nn.Sequential(net.avgpool), nn.Sequential(net.fc)
    return net.features, nn.Sequential(), net.classifier


def set_batch_norm_train(seq, train):
    """reference: model/nn_utils.py:135-155 (every BatchNorm2d under seq)"""
    for m in seq.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.train(mode=train)


def set_net_train(net, train, bn_train=False):
    """reference: model/nn_utils.py:160-163"""
    net.train(mode=train)
    if train and not bn_train:
        set_batch_norm_train(net.features, False)
