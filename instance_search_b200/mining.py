"""All-pairs similarities and (semi-)hard negative selection on the GPU.

Host-side mirror of utils/train_siamese.py:14-55 and the mining block of
train/siamese_regions.py:106-129 (same code train/siamese_descriptor.py:108-131).
"""

import torch

from . import _lib, ops
from ._lib import IsbError
from .sharding import all_gather_rows, shard_bounds

SCREEN_EPS_3TERM = 2e-5   # absolute error bound of the [hi|lo|hi].[hi|hi|lo] screen, unit rows
# plain bf16 screen: expected rms error of a score on dense rows, both operands rounded to 8
# significant bits: 2.34e-3 |a||b| / sqrt(D), calibrated on unit Gaussian rows at D = 128 and 2048
# (include/isb.h, isb_select_negatives)
SIGMA_BF16 = 2.34e-3


def label_ids(dataset):
    """Dense int ids of the labels of a reference-style dataset (list of
    (tensor, label, name) triples), in first-appearance order, + the label list."""
    seen, ids = {}, []
    for _, lab, _ in dataset:
        ids.append(seen.setdefault(lab, len(seen)))
    return torch.tensor(ids, dtype=torch.int32), list(seen.keys())


def get_lab_indicators(dataset, device):
    """{label: uint8 mask [N]} -- reference: utils/train_siamese.py:14-25.
    device >= 0 -> current CUDA device, < 0 -> CPU (utils/general.py:94-98)."""
    ids, labels = label_ids(dataset)
    if device >= 0:
        ids = ids.cuda()
    return {lab: (ids == i).to(torch.uint8) for i, lab in enumerate(labels)}


def embeddings_device_dim(P, net, n, sim_matrix=False):
    """Placement rule of the reference -- utils/train_siamese.py:30-43."""
    device = P.cuda_device
    out_size = P.feature_dim
    if hasattr(net, 'feature_size') and out_size <= 0:
        out_size = net.feature_size
    if n * out_size * 4 > P.embeddings_cuda_size:
        device = -1
    if sim_matrix and n * n * 4 > P.embeddings_cuda_size:
        device = -1
    return device, out_size


def all_pairs_similarities(emb, terms=3):
    """S = E . E^T, fp32 [N, N] -- reference: utils/train_siamese.py:53
    (also test/instance_avg.py:12).  terms=3: fp32-grade split-operand product."""
    ops._need_cuda(emb)
    emb = ops._f32c(emb)
    hi = ops.to_bf16(emb, 0)
    if terms == 1:
        return ops.gemm_nt(hi, hi)
    lo = ops.to_bf16(emb, 1)
    return ops.gemm_nt_split(hi, lo, hi, lo)


class MiningIndex(object):
    """Embeddings of the reference set prepared for negative selection.

    terms=1 (default): ONE bf16 tcgen05 product per tile screens all couples; every couple carries
    a completeness certificate (8 sigma of the measured screen noise), and the couples it
    rejects -- and only those -- are re-screened with fp32-grade split operands (three products
    per tile) and, failing that, exhaustively: the same three lines as the search
    (isb_topk_search -> isb_topk_resolve -> isb_topk_exhaustive).  Costs one 4-byte
    device->host read per call.  terms=3: every couple goes through the split-operand screen
    (no host read; exhaustive pass in the same call)."""

    def __init__(self, emb, labels, terms=1):
        ops._need_cuda(emb)
        if terms not in (1, 3):
            raise IsbError("terms must be 1 or 3")
        emb = ops._f32c(emb)
        self.N, self.dim = emb.shape
        pad = (-self.dim) % 8
        if pad:
            emb = torch.nn.functional.pad(emb, (0, pad))
        self.emb = emb.contiguous()
        self.labels = labels.to(device=emb.device, dtype=torch.int32).contiguous()
        if self.labels.numel() != self.N:
            raise IsbError("one label id per embedding expected")
        self.terms = terms
        self.emb_hi = ops.to_bf16(self.emb, 0)
        self._emb_lo = ops.to_bf16(self.emb, 1) if terms == 3 else None
        # sigma floor of the plain screen's certificate: expected bf16 noise for the largest rows
        self.sigma_floor = SIGMA_BF16 * float(self.emb.square().sum(1).max()) / self.emb.size(1) ** 0.5
        self.last_bruteforce = None      # device int32 [1]: couples answered by the exhaustive pass
        self.last_second_line = 0        # couples the plain screen's certificate rejected (terms=1)

    def _lo(self):
        if self._emb_lo is None:
            self._emb_lo = ops.to_bf16(self.emb, 1)
        return self._emb_lo

    def _select(self, anchors, positives, semi_hard, split, unc_rows, n_unc, out):
        L = _lib.lib()
        D, P = self.emb.size(1), anchors.numel()
        nbytes = L.isb_select_negatives_workspace_bytes(P, self.N, D, 1 if split else 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.emb.device)
        neg_idx, neg_sim, pos_sim = out
        _lib.check(L.isb_select_negatives(self.emb.data_ptr(), self.emb_hi.data_ptr(),
                                          self._lo().data_ptr() if split else 0, self.emb_hi.size(1), self.N, D,
                                          self.labels.data_ptr(), anchors.data_ptr(), positives.data_ptr(), P,
                                          1 if semi_hard else 0, SCREEN_EPS_3TERM if split else 0.0,
                                          0.0 if split else self.sigma_floor,
                                          neg_idx.data_ptr(), neg_sim.data_ptr(), pos_sim.data_ptr(),
                                          ops._ptr(unc_rows), n_unc.data_ptr(), ws.data_ptr(), nbytes,
                                          ops._stream()), "isb_select_negatives")

    def select_negatives(self, anchors, positives, semi_hard):
        """One negative per positive couple (anchors[p], positives[p]).

        reference: train/siamese_regions.py:106-129.  Returns (neg_idx [P] int64,
        -1 where every item is excluded -> caller draws a random negative as the
        reference does; neg_sim [P]; pos_sim [P]).
        """
        dev = self.emb.device
        anchors = torch.as_tensor(anchors, dtype=torch.int64, device=dev).contiguous()
        positives = torch.as_tensor(positives, dtype=torch.int64, device=dev).contiguous()
        P = anchors.numel()
        neg_idx = torch.empty(P, dtype=torch.int64, device=dev)
        neg_sim = torch.empty(P, dtype=torch.float32, device=dev)
        pos_sim = torch.empty(P, dtype=torch.float32, device=dev)
        if P == 0:
            return neg_idx, neg_sim, pos_sim
        n_unc = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.terms == 3:
            self._select(anchors, positives, semi_hard, True, None, n_unc, (neg_idx, neg_sim, pos_sim))
            self.last_bruteforce, self.last_second_line = n_unc, 0
            return neg_idx, neg_sim, pos_sim
        unc_rows = torch.empty(P, dtype=torch.int32, device=dev)
        self._select(anchors, positives, semi_hard, False, unc_rows, n_unc, (neg_idx, neg_sim, pos_sim))
        n_bad = int(n_unc.item())            # the one 4-byte D2H read of the exactness guarantee
        self.last_second_line = n_bad
        self.last_bruteforce = torch.zeros(1, dtype=torch.int32, device=dev)
        if n_bad:
            rows = unc_rows[:n_bad].long()
            sub = tuple(torch.empty(n_bad, dtype=t.dtype, device=dev) for t in (neg_idx, neg_sim, pos_sim))
            self._select(anchors[rows].contiguous(), positives[rows].contiguous(), semi_hard, True, None,
                         self.last_bruteforce, sub)
            neg_idx[rows], neg_sim[rows] = sub[0], sub[1]
        return neg_idx, neg_sim, pos_sim


class ShardedMiner(object):
    """Negative selection over several GPUs: the embeddings are replicated
    (every rank holds all of E: 134 MB at 16k x 2048), the positive couples are split contiguously
    across the ranks, every rank selects the negatives of ITS couples (MiningIndex, no exchange on
    the data path), and ONE all-gather of the packed results -- 16 bytes per couple -- gives every
    rank the full answer.  Identical to MiningIndex.select_negatives on one GPU.

    reference: train/siamese_regions.py:106-129 over the matrix of utils/train_siamese.py:48-55
    (single device there)."""

    def __init__(self, emb, labels, rank=0, world_size=1, group=None, terms=1):
        self.rank, self.world_size, self.group = rank, world_size, group
        self.index = self._make_index(emb, labels, terms)

    # hooks (the gloo CPU test replaces them to exercise the plumbing)
    def _make_index(self, emb, labels, terms):
        return MiningIndex(emb, labels, terms=terms)

    def _local_select(self, anchors, positives, semi_hard):
        return self.index.select_negatives(anchors, positives, semi_hard)

    def select_negatives(self, anchors, positives, semi_hard):
        """Collective: every rank passes the SAME couples; returns (neg_idx [P] int64, neg_sim [P],
        pos_sim [P]) of all couples on every rank."""
        anchors = torch.as_tensor(anchors, dtype=torch.int64)
        positives = torch.as_tensor(positives, dtype=torch.int64)
        P = anchors.numel()
        lo, hi = shard_bounds(P, self.world_size)[self.rank]
        neg, nsim, psim = self._local_select(anchors[lo:hi], positives[lo:hi], semi_hard)
        if self.world_size == 1:
            return neg, nsim, psim
        # pack (neg int64 | the two fp32 similarities as one 64-bit word) -> one collective
        sims = torch.stack([nsim.float(), psim.float()], 1).contiguous().view(torch.int64)     # [p, 1]
        packed = torch.cat([neg.view(-1, 1), sims.view(-1, 1)], 1)                              # [p, 2]
        full = all_gather_rows(packed, P, self.rank, self.world_size, self.group)
        sims = full[:, 1].contiguous().view(torch.float32).view(-1, 2)
        return full[:, 0].contiguous(), sims[:, 0].contiguous(), sims[:, 1].contiguous()
