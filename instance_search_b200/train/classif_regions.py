"""Hot-path function of the reference's ``train/classif_regions.py``: ``get_embeddings`` for the
classification track's sub-window net (``TuneClassifSub``) -- the class scores at the window with
the highest maximal activation, L2-normalised (train/classif_regions.py:107-132).  The rest of
that script (classification fine-tuning) is out of scope; this is the sibling
embedding head SURVEY.md 8f (rank 4) lists next to the batched ``get_embeddings``."""

import torch

from .. import regions


def _head(net):
    conv = net.classifier[0]
    key = (conv.weight.data_ptr(), conv.weight._version, conv.bias.data_ptr(), conv.bias._version)
    cached = getattr(net, "_isb_classif_head", None)
    if cached is None or cached[0] != key:
        hw = object.__new__(regions.HeadWeights)
        w = conv.weight.data.reshape(conv.out_channels, -1).contiguous().float()
        hw.terms, hw.cls_w, hw.cls_b = 1, w, conv.bias.data.contiguous().float()
        hw.cls_w_hi, hw.cls_w_lo = regions.ops.to_bf16(w, 0), regions.ops.to_bf16(w, 1)
        hw.cls_w_absmax = float(w.abs().max())
        cached = (key, hw)
        net._isb_classif_head = cached
    return cached[1]


def get_embeddings(net, dataset, device, out_size, batch_size=32, transform=None, feature_size2d=None):
    """Embeddings [len(dataset), ncls] of a TuneClassifSub net in eval mode over a reference-style data
    set; ``device`` >= 0: current CUDA device, < 0: host.  reference: train/classif_regions.py:107-132
    (one image per forward there; here equal-sized images are stacked and go through the fused head).
    The net must be the ResNet form: ``feature_reduc`` = AvgPool(window, stride 1), ``classifier`` =
    one 1x1 conv; feature_size2d defaults to the pooling window."""
    pool = net.feature_reduc[0]
    fsize = tuple(feature_size2d) if feature_size2d is not None else tuple(
        pool.kernel_size if isinstance(pool.kernel_size, (tuple, list)) else (pool.kernel_size,) * 2)
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
            emb, _ = regions.classif_embeddings(net.features(x), _head(net), fsize)
            out[i:j] = emb
            i = j
    return out if device >= 0 else out.cpu()
