"""Descriptor database + top-k search (single GPU and row-sharded multi-GPU).

The reference recomputes ``sim = torch.mm(test_emb, ref_emb.t())`` on every run
(test/siamese_regions_test.py:76, utils/train_siamese.py:70) and ranks it on the
host (utils/metrics.py).  Here the reference set lives in HBM as a
``DescriptorIndex`` (fp32 rows for the exact re-rank + a bf16 copy for the
tensor-core screen) and is searched without materialising the Q x N matrix.

Multi-GPU: the database is split row-wise across the
ranks of a ``torch.distributed`` group, one process per GPU; every rank
searches its shard for the full (replicated) query batch, the per-shard
``[Q, k]`` (score, global index) lists are exchanged with ONE all-gather over
NCCL/NVLink, and every rank merges them (``isb_topk_merge``).
"""

import torch

from . import ops
from ._lib import IsbError


def shard_bounds(n_rows, world_size):
    """Row range [lo, hi) of every rank: as even as possible, contiguous."""
    base, extra = divmod(n_rows, world_size)
    bounds, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


class DescriptorIndex(object):
    """A block of database descriptors resident on one GPU.

    db: [N, D] fp32 CUDA tensor (rows L2-normalised by the caller, as
    model/siamese.py:222 guarantees); row_offset = global index of row 0.
    """

    def __init__(self, db, row_offset=0):
        if not db.is_cuda or db.dtype != torch.float32 or db.dim() != 2:
            raise IsbError("DescriptorIndex needs a [N, D] float32 CUDA tensor")
        d = db.size(1)
        self.dim = d
        self.pad = (-d) % 8          # the kernels want D % 8 == 0: zero-pad (scores unchanged)
        if self.pad:
            db = torch.nn.functional.pad(db, (0, self.pad))
        self.db_f32 = db.contiguous()
        self.db_bf16 = ops.to_bf16(self.db_f32)
        self._db_lo = None            # lo bf16 term, built the first time a row needs resolving
        self.row_offset = int(row_offset)
        self._ws = None
        self.stats = {}               # rows searched / resolved fp32-grade / resolved exhaustively

    def __len__(self):
        return self.db_f32.size(0)

    def _queries(self, q):
        if not q.is_cuda:
            raise IsbError("queries must be a CUDA tensor (no CPU fallback)")
        if q.size(1) != self.dim:
            raise IsbError("query dim %d != index dim %d" % (q.size(1), self.dim))
        if self.pad:
            q = torch.nn.functional.pad(q, (0, self.pad))
        return q.contiguous()

    def _workspace(self, Q, k, margin):
        need = ops._lib.lib().isb_topk_search_workspace_bytes(Q, len(self), self.db_f32.size(1), k, margin)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.db_f32.device)
        return self._ws

    def _lo(self):
        if self._db_lo is None:
            self._db_lo = ops.to_bf16(self.db_f32, 1, ld=self.db_bf16.size(1))
        return self._db_lo

    def search(self, q, k, margin=None, events=None, exact=True):
        """(scores [Q, k] fp32, idx [Q, k] int64 global), best first.

        exact=True: rows whose candidate list the bf16 screen cannot certify
        complete are re-screened with fp32-grade operands / exhaustively
        (ops.resolve_uncertified); costs one 4-byte device->host read per call.

        events: optional list; when given, CUDA events bracketing the screen
        stage (the tcgen05 GEMM + streaming top-k) are appended as a pair so a
        caller can time the dominant kernel on the launching stream.
        """
        q = self._queries(q)
        Q = q.size(0)
        k_eff = min(k, len(self))
        if margin is None:
            margin = min(ops.DEFAULT_MARGIN, ops.MAX_CANDIDATES - k_eff)
        if k_eff + margin > ops.MAX_CANDIDATES:
            raise IsbError("k + margin must be <= %d" % ops.MAX_CANDIDATES)
        scores = torch.empty((Q, k_eff), dtype=torch.float32, device=q.device)
        idx = torch.empty((Q, k_eff), dtype=torch.int64, device=q.device)
        if Q == 0:
            return scores, idx
        ws = self._workspace(Q, k_eff, margin)
        L = ops._lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        N, D = self.db_f32.shape
        if events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ops._lib.check(L.isb_topk_screen(q.data_ptr(), Q, self.db_bf16.data_ptr(), N, D,
                                         self.db_bf16.size(1), k_eff, margin, ws.data_ptr(),
                                         ws.numel(), st), "isb_topk_screen")
        if events is not None:
            e1.record()
            events.append((e0, e1))
        unc_rows = torch.empty(Q, dtype=torch.int32, device=q.device) if exact else None
        n_unc = torch.zeros(1, dtype=torch.int32, device=q.device) if exact else None
        ops._lib.check(L.isb_topk_rerank(q.data_ptr(), Q, self.db_f32.data_ptr(), N, D, k_eff, margin,
                                         self.row_offset, scores.data_ptr(), idx.data_ptr(),
                                         ops._ptr(unc_rows), ops._ptr(n_unc),
                                         ws.data_ptr(), ws.numel(), st), "isb_topk_rerank")
        if exact:
            ops.resolve_uncertified(q, self.db_f32, self.db_bf16, self._lo, k_eff, margin, self.row_offset,
                                    scores, idx, unc_rows, n_unc, self.stats)
        return scores, idx


class ShardedIndex(object):
    """Row-sharded database over a torch.distributed process group.

    Every rank constructs it with ITS shard (rows [lo, hi) of the global
    database, see shard_bounds) and calls search() collectively with the same
    queries.  world_size 1 degenerates to DescriptorIndex.search.
    """

    def __init__(self, local_db, n_total, rank=0, world_size=1, group=None):
        self.rank, self.world_size, self.group = rank, world_size, group
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(self.n_total, world_size)[rank]
        if local_db.size(0) != self.hi - self.lo:
            raise IsbError("rank %d: shard has %d rows, expected %d" %
                           (rank, local_db.size(0), self.hi - self.lo))
        self.local = self._make_local(local_db, self.lo)

    # hooks (the gloo CPU tests replace them to exercise the plumbing)
    def _make_local(self, local_db, row_offset):
        return DescriptorIndex(local_db, row_offset)

    def _local_search(self, q, k, events=None):
        return self.local.search(q, k, events=events)

    def _merge(self, cand_scores, cand_idx):
        return ops.topk_merge(cand_scores, cand_idx)

    def search(self, q, k, events=None):
        import torch.distributed as dist
        k_local = min(k, self.hi - self.lo)
        s, i = self._local_search(q, k_local, events)
        if self.world_size == 1:
            return s, i
        if k_local < k:  # a shard smaller than k: pad with invalid entries
            pad = k - k_local
            s = torch.cat([s, s.new_full((s.size(0), pad), float("-inf"))], 1)
            i = torch.cat([i, i.new_full((i.size(0), pad), -1)], 1)
        Q = s.size(0)
        gs = torch.empty((self.world_size, Q, k), dtype=s.dtype, device=s.device)
        gi = torch.empty((self.world_size, Q, k), dtype=i.dtype, device=i.device)
        # the one exchange step of the path: k candidates per query per shard
        # (concatenated [R*Q, k] view: the layout both the NCCL and the gloo backend accept)
        dist.all_gather_into_tensor(gs.view(-1, k), s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi.view(-1, k), i.contiguous(), group=self.group)
        ms, mi = self._merge(gs, gi)
        kk = min(k, self.n_total)
        return ms[:, :kk], mi[:, :kk]
