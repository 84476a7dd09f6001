"""Drop-in for the hot-path functions of the reference's ``utils/train_siamese.py``.

    get_lab_indicators      utils/train_siamese.py:14-25
    embeddings_device_dim   utils/train_siamese.py:30-43
    get_similarities        utils/train_siamese.py:48-55
    test_descriptor_net     utils/train_siamese.py:61-82

``get_embeddings`` is passed in by the caller exactly as in the reference
(``train/siamese_regions.py:26``; this package's batched version lives in
``instance_search_b200.train.siamese_regions``).  Device convention of the
reference (utils/general.py:94-98): ``device >= 0`` = current CUDA device,
``< 0`` = host.  All arithmetic runs on the GPU; the placement rule only decides
where the RESULT lives, as it does in the reference.
"""

import torch

from .. import mining, ops
from ..model.nn_utils import set_net_train
from . import metrics

get_lab_indicators = mining.get_lab_indicators
embeddings_device_dim = mining.embeddings_device_dim


def _place(t, device):
    return t.cuda() if device >= 0 else t.cpu()


def get_similarities(P, get_embeddings, net, dataset):
    """All-pairs similarities of the data set's descriptors, net left in train mode.
    reference: utils/train_siamese.py:48-55.  Returns (similarities [N, N], device)."""
    set_net_train(net, False)
    n = len(dataset)
    d, o = embeddings_device_dim(P, net, n, sim_matrix=True)
    embeddings = get_embeddings(net, dataset, d, o)
    if d >= 0:
        similarities = mining.all_pairs_similarities(embeddings.cuda())            # :53 torch.mm(E, E.t())
    else:
        # the reference keeps a matrix above P.embeddings_cuda_size on the host (:41-42): compute it
        # in row blocks on the GPU and stream them out, the N x N matrix never exists in HBM
        similarities = _similarities_to_host(embeddings.cuda())
    set_net_train(net, True, bn_train=P.train_bn)
    return similarities, d


def _similarities_to_host(emb, block_rows=4096):
    """S = E . E^T as a pinned host tensor, one [block_rows, N] slab at a time (split-operand
    tcgen05 GEMM per slab, D2H copy of slab i overlapping the product of slab i + 1)."""
    n = emb.size(0)
    hi, lo = ops.to_bf16(emb, 0), ops.to_bf16(emb, 1)
    out = torch.empty((n, n), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    for s in range(0, n, block_rows):
        e = min(s + block_rows, n)
        slab = ops.gemm_nt_split(hi[s:e], lo[s:e], hi, lo)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            out[s:e].copy_(slab, non_blocking=True)
            slab.record_stream(copy_stream)
    copy_stream.synchronize()
    return out


def test_descriptor_net(P, get_embeddings, net, test_set, test_ref_set, kth=1):
    """P@1 / mAP / similarity statistics of a descriptor net.
    reference: utils/train_siamese.py:61-82 (same return tuple)."""
    d, o = embeddings_device_dim(P, net, max(len(test_set), len(test_ref_set)))
    test_embeddings = get_embeddings(net, test_set, d, o).cuda()
    ref_embeddings = get_embeddings(net, test_ref_set, d, o).cuda()
    # :70  sim = torch.mm(test, ref.t())  -- fp32-grade split-operand tcgen05 GEMM
    hi_t, lo_t = ops.to_bf16(test_embeddings, 0), ops.to_bf16(test_embeddings, 1)
    hi_r, lo_r = ops.to_bf16(ref_embeddings, 0), ops.to_bf16(ref_embeddings, 1)
    sim = ops.gemm_nt_split(hi_t, lo_t, hi_r, lo_r)
    prec1, correct, total, max_sim, max_label = metrics.precision1(sim, test_set, test_ref_set, kth)
    mAP = metrics.mean_avg_precision(sim, test_set, test_ref_set, kth)
    # :74-76  similarity mass on positive / negative pairs
    ids = {}
    ref_ids = torch.tensor([ids.setdefault(lab, len(ids)) for _, lab, _ in test_ref_set], device=sim.device)
    test_ids = torch.tensor([ids.get(lab, -1) for _, lab, _ in test_set], device=sim.device)
    same = test_ids.view(-1, 1) == ref_ids.view(1, -1)
    sum_pos = float((sim * same).sum(dtype=torch.float64))
    sum_neg = float(sim.sum(dtype=torch.float64)) - sum_pos
    sum_max = float(max_sim.sum(dtype=torch.float64))
    lab_dict = dict([(lab, {}) for _, lab, _ in test_set])       # :77-81
    for j, (_, lab, _) in enumerate(test_set):
        dct = lab_dict[lab]
        lab = max_label[j]
        dct.setdefault(lab, dct.get(lab, 0) + 1)
    return prec1, correct, total, sum_pos, sum_neg, sum_max, mAP, lab_dict
