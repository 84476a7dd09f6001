"""Hot-path functions of the reference's ``train/siamese_regions.py``:

    get_embeddings              train/siamese_regions.py:26-41
    negative selection          train/siamese_regions.py:106-129 (create_batch)
    get_siamese_net             train/siamese_regions.py:157-168

The reference embeds ONE image per forward (``fold_batches(..., 1)``, an H2D copy
and a Python iteration per image).  Here consecutive images of identical size
are stacked from pinned host memory and go through the trunk and the fused CUDA
head as a batch; the descriptors equal the per-image ones (the head treats every
image as the reference's batch-1 call).
"""

import torch

from .. import mining


def get_embeddings(net, dataset, device, out_size, batch_size=32, transform=None, rank=0, world_size=1,
                   group=None):
    """Descriptors [len(dataset), out_size] of a reference-style data set (list of
    ``(image tensor [3, h, w], label, name)``), stored on ``device`` (>= 0: current
    CUDA device, < 0: host -- utils/general.py:94-98).
    reference: train/siamese_regions.py:26-41 (net must be in eval mode).

    world_size > 1: collective over a torch.distributed group, one process
    per GPU with a replica of the net -- the images are split contiguously across the ranks
    (independent units, no exchange on the data path), every rank embeds its slice, and one
    all-gather of the [n / R, D] blocks gives every rank the full matrix."""
    if world_size > 1:
        from ..sharding import all_gather_rows, shard_bounds
        lo, hi = shard_bounds(len(dataset), world_size)[rank]
        local = get_embeddings(net, dataset[lo:hi], 0, out_size, batch_size, transform)
        full = all_gather_rows(local, len(dataset), rank, world_size, group)
        return full if device >= 0 else full.cpu()
    n = len(dataset)
    out = torch.empty((n, out_size), dtype=torch.float32, device="cuda")
    i = 0
    with torch.no_grad():
        while i < n:
            first = dataset[i][0] if transform is None else transform(dataset[i][0])
            batch = [first]
            j = i + 1
            while j < n and len(batch) < batch_size:
                im = dataset[j][0] if transform is None else transform(dataset[j][0])
                if im.shape != first.shape:
                    break
                batch.append(im)
                j += 1
            x = torch.stack(batch)
            if not x.is_cuda:
                x = x.pin_memory().cuda(non_blocking=True)
            desc = net(x)
            if isinstance(desc, tuple):          # a net left in train mode returns (desc, cls_out)
                desc = desc[0]
            out[i:j] = desc
            i = j
    return out if device >= 0 else out.cpu()


class NegativeSelector(object):
    """Batched form of the mining block of create_batch (train/siamese_regions.py:106-129).

    Built once per epoch from the embeddings ``get_similarities`` would have
    multiplied (utils/train_siamese.py:52-53); ``select`` answers, for any number
    of positive couples at once, what the reference computes per training sample
    from a row of the N x N matrix -- which is never materialised here.
    """

    def __init__(self, embeddings, dataset, rank=0, world_size=1, group=None):
        ids, self.labels = mining.label_ids(dataset)
        # world_size > 1: the couples of every select() are split across the ranks (ShardedMiner)
        self.index = mining.ShardedMiner(embeddings.cuda(), ids, rank, world_size, group)

    def select(self, couples, epoch, train_epoch_switch):
        """couples: sequence of (i1, i2).  Returns a list with, per couple, the index
        of the chosen negative or None when every item is excluded (the reference
        then falls back to choose_rand_neg, utils/dataset.py:57-61)."""
        if len(couples) == 0:
            return []
        a = torch.tensor([c[0] for c in couples], dtype=torch.int64)
        b = torch.tensor([c[1] for c in couples], dtype=torch.int64)
        neg, _, _ = self.index.select_negatives(a, b, epoch < train_epoch_switch)   # :107
        return [None if v < 0 else v for v in neg.tolist()]


def get_siamese_net(P, pretrained=True, base_net=None):
    """The net of the region-siamese track: torchvision trunk -> TuneClassifSub ->
    RegionDescriptorNet, optional checkpoints, moved to ``P.cuda_device``.
    reference: train/siamese_regions.py:157-168.  ``P`` is passed explicitly (the
    reference reads a module-global); ``pretrained=False`` / ``base_net`` exist because the
    torchvision weights need a download the reference takes for granted."""
    from ..model.siamese import RegionDescriptorNet, TuneClassifSub
    if base_net is None:
        import torchvision.models as models
        model = models.resnet152 if P.cnn_model.lower() == 'resnet152' else models.alexnet
        base_net = model(weights="DEFAULT" if pretrained else None)
    class_net = TuneClassifSub(base_net, P.num_classes, P.feature_size2d, untrained=P.untrained_blocks)
    if getattr(P, "classif_model", None):
        class_net.load_state_dict(torch.load(P.classif_model, map_location="cpu"))
    net = RegionDescriptorNet(class_net, P.regions_k, P.feature_dim, P.feature_size2d,
                              untrained=P.untrained_blocks)
    if getattr(P, "preload_net", None):
        net.load_state_dict(torch.load(P.preload_net, map_location="cpu"))
    return net.cuda() if P.cuda_device >= 0 else net
