"""Descriptor database + top-k search (single GPU and row-sharded multi-GPU).

The reference recomputes ``sim = torch.mm(test_emb, ref_emb.t())`` on every run
(test/siamese_regions_test.py:76, utils/train_siamese.py:70) and ranks it on the
host (utils/metrics.py).  Here the reference set lives in HBM as a
``DescriptorIndex`` (fp32 rows for the exact re-rank + a bf16 copy for the
tensor-core screen) and is searched without materialising the Q x N matrix.

Multi-GPU: the database is split row-wise across the
ranks of a ``torch.distributed`` group, one process per GPU; every rank screens
its shard for the full (replicated) query batch.  Candidate exchange: the
shards all-gather the screen scores of their k + margin candidates (NCCL over
NVLink), derive the same global threshold, re-rank exactly only their own
candidates above it, all-gather the packed per-shard lists and merge them with
the global completeness certificate (``isb_topk_merge_certified``).
"""

import torch

from . import ops
from ._lib import IsbError
from .sharding import shard_bounds  # noqa: F401  (re-exported: callers import it from here)


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
        self._ws_pipe = [None, None]  # deferred searches alternate between two workspaces ...
        self._tail_done = [None, None]   # ... each guarded by the event that ends its re-rank
        self._slot = 0
        self._tail = None             # side stream of the deferred (pipelined) searches
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

    def _workspace(self, Q, k, margin, slot=None):
        need = ops._lib.lib().isb_topk_search_workspace_bytes(Q, len(self), self.db_f32.size(1), k, margin)
        if slot is not None:
            if self._ws_pipe[slot] is None or self._ws_pipe[slot].numel() < need:
                self._ws_pipe[slot] = torch.empty(need, dtype=torch.uint8, device=self.db_f32.device)
            return self._ws_pipe[slot]
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.db_f32.device)
        return self._ws

    def tail_stream(self):
        """Side stream on which a deferred search runs everything after its screen."""
        if self._tail is None:
            self._tail = torch.cuda.Stream(device=self.db_f32.device)
        return self._tail

    def _lo(self):
        if self._db_lo is None:
            self._db_lo = ops.to_bf16(self.db_f32, 1, ld=self.db_bf16.size(1))
        return self._db_lo

    def _screen(self, q, k_eff, margin, events=None, slot=None):
        """Stage 1 (isb_topk_screen): bf16 tcgen05 GEMM + streaming top-(k+margin) into the
        candidate pool of the workspace, which is returned.  q: padded fp32 queries."""
        Q = q.size(0)
        ws = self._workspace(Q, k_eff, margin, slot)
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
        return ws

    def candidates(self, q, k, kc, events=None):
        """Sharded search, local stage 1: screen this shard and list its kc best screen
        entries per query: (cand_screen [Q, kc] fp32, cand_col [Q, kc] int32 local row;
        -inf / -1 where the shard has fewer than kc rows)."""
        q = self._queries(q)
        k_eff = min(k, len(self))
        self._workspace(q.size(0), k_eff, min(kc, len(self)) - k_eff)
        return ops.topk_candidates(q, self.db_bf16, k, kc, workspace=self._ws, events=events)

    def rerank_owned(self, q, k, cand_screen, cand_col, thr):
        """Sharded search, local stage 2: exact scores of this shard's candidates at or above
        the global threshold, packed for the second all-gather (ops.topk_rerank_owned)."""
        return ops.topk_rerank_owned(self._queries(q), self.db_f32, k, cand_screen, cand_col, thr)

    def search(self, q, k, margin=None, events=None, exact=True, defer=False):
        """(scores [Q, k] fp32, idx [Q, k] int64 global), best first.

        defer=True: returns (scores, idx, ticket) without any host synchronisation; the
        caller resolves the ticket (ops.ExactnessTicket) before it consumes the result --
        after it has queued the next batch, so the GPU never idles on the round trip.
        The screen runs on the caller's stream, the exact re-rank (HBM gathers, no tensor
        cores) on a side stream: queued back to back, the re-rank of batch i runs UNDER the
        screen of batch i + 1 (two workspaces alternate).

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
            return (scores, idx, None) if defer else (scores, idx)
        L = ops._lib.lib()
        N, D = self.db_f32.shape
        cur = torch.cuda.current_stream()
        pipelined = defer and exact
        slot = None
        if pipelined:
            slot = self._slot
            self._slot ^= 1
            if self._tail_done[slot] is not None:
                cur.wait_event(self._tail_done[slot])      # the re-rank that last read this workspace
        ws = self._screen(q, k_eff, margin, events, slot)
        tail = cur
        if pipelined:
            screened = torch.cuda.Event()
            screened.record(cur)
            tail = self.tail_stream()
            tail.wait_event(screened)
            for t in (q, scores, idx):
                t.record_stream(tail)
        with torch.cuda.stream(tail):
            unc_rows = torch.empty(Q, dtype=torch.int32, device=q.device) if exact else None
            n_unc = torch.zeros(1, dtype=torch.int32, device=q.device) if exact else None
            ops._lib.check(L.isb_topk_rerank(q.data_ptr(), Q, self.db_f32.data_ptr(), N, D, k_eff, margin,
                                             self.row_offset, scores.data_ptr(), idx.data_ptr(),
                                             ops._ptr(unc_rows), ops._ptr(n_unc),
                                             ws.data_ptr(), ws.numel(), tail.cuda_stream), "isb_topk_rerank")
            ticket = None
            if exact:
                ticket = ops.resolve_uncertified(q, self.db_f32, self.db_bf16, self._lo, k_eff, margin,
                                                 self.row_offset, scores, idx, unc_rows, n_unc, self.stats,
                                                 defer=defer)
            if pipelined:
                self._tail_done[slot] = torch.cuda.Event()
                self._tail_done[slot].record(tail)
                for t in (unc_rows, n_unc):
                    t.record_stream(cur)                   # a fix-up at resolve() time runs on the caller's stream
        return (scores, idx, ticket) if defer else (scores, idx)


class _NoStream(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


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
        self.device = local_db.device
        self.stats = {}
        self._row_offsets = None
        # reduced candidate lists (see list_length); switched off by the first batch whose rows
        # the reduction sends to the second line in numbers (a database sorted by instance puts a
        # query's neighbours on ONE shard)
        self.reduce_lists = True
        self.local = self._make_local(local_db, self.lo)

    # hooks (the gloo CPU tests replace them to exercise the plumbing)
    def _make_local(self, local_db, row_offset):
        return DescriptorIndex(local_db, row_offset)

    def _local_search(self, q, k, events=None, defer=False):
        return self.local.search(q, k, events=events, defer=defer)

    def _merge(self, cand_scores, cand_idx):
        return ops.topk_merge(cand_scores, cand_idx)

    def _local_candidates(self, q, k, kc, events=None):
        return self.local.candidates(q, k, kc, events)

    def _global_threshold(self, all_screen, kth):
        return ops.topk_global_threshold(all_screen, kth)

    def list_length(self, kc):
        """Candidates a shard lists per query.  The global top kc = k + margin spreads over the R
        shards (kc / R per shard on average when rows are placed independently of their content),
        so a shard lists mean + 6 sigma + 8 of them instead of kc: its screen keeps fewer rows
        per query -- 48 instead of 128 at R = 8, which is 9 % of the screen's time on a 125k-row
        shard (threshold warm-up, buffer compactions).  A list that turns out to have been cut
        above the global threshold is detected (isb_topk_rerank_owned) and the row goes to the
        second line, so the result does not depend on the assumption -- only the speed does."""
        R = self.world_size
        if R == 1 or not self.reduce_lists:
            return kc
        mean = kc / float(R)
        return min(kc, int(mean + 6.0 * mean ** 0.5 + 0.999) + 8)

    def _rerank_owned(self, q, k, cand_screen, cand_col, thr):
        return self.local.rerank_owned(q, k, cand_screen, cand_col, thr)

    def _merge_certified(self, packed_all, thr, k):
        if self._row_offsets is None:
            self._row_offsets = torch.tensor([lo for lo, _ in shard_bounds(self.n_total, self.world_size)],
                                             dtype=torch.int64, device=packed_all.device)
        return ops.topk_merge_certified(packed_all, self._row_offsets, thr, k)

    def _gather(self, t):
        """all-gather of a [Q, c] tensor -> [R, Q, c] (the [R*Q, c] view is the layout both
        the NCCL and the gloo backend accept)."""
        import torch.distributed as dist
        out = torch.empty((self.world_size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1, t.size(-1)), t.contiguous(), group=self.group)
        return out

    def upload_queries(self, q_host):
        """Pinned host queries -> device on every rank: each rank copies 1/R of the rows over
        its own PCIe link and the slices are all-gathered over NVLink (instead of R full
        host->device copies)."""
        if self.world_size == 1 or q_host.size(0) < self.world_size:
            return q_host.to(self.device, non_blocking=True)
        lo, hi = shard_bounds(q_host.size(0), self.world_size)[self.rank]
        return self.gather_queries(q_host[lo:hi].to(self.device, non_blocking=True), q_host.size(0))

    def gather_queries(self, part, n_queries):
        """part: this rank's shard_bounds slice of the query rows, already on the device (e.g.
        uploaded on a copy stream while the previous batch was searched) -> all n_queries rows
        on every rank (one NVLink all-gather)."""
        from .sharding import all_gather_rows
        return all_gather_rows(part, n_queries, self.rank, self.world_size, self.group)

    def search(self, q, k, events=None, exchange=True, defer=False):
        """(scores [Q, kk] fp32, idx [Q, kk] int64 global), kk = min(k, n_total), best first,
        identical on every rank.  q: the same queries on every rank (CUDA tensor, or a pinned
        host tensor -> upload_queries).

        defer=True: returns (scores, idx, ticket): no host synchronisation inside; resolve the
        ticket (collectively, on every rank) before consuming the result -- typically after the
        next batch has been queued.

        exchange=True (default): candidate exchange -- the shards all-gather the screen scores
        of their candidates, agree on the global (k + margin)-th best and re-rank only their
        own candidates above it, so the exact re-rank costs k + margin gathers per query in
        total instead of per shard; then ONE all-gather of the per-shard lists (packed: scores,
        local rows, noise statistics) and the merge with the global completeness certificate.  exchange=False: every shard runs the
        full single-GPU search (screen + re-rank of k + margin) before the all-gather."""
        if not q.is_cuda and self.device.type == "cuda":
            q = self.upload_queries(q)
        if self.world_size == 1:
            return self._local_search(q, min(k, self.hi - self.lo), events, defer)
        if not exchange:
            r = self._search_replicated_rerank(q, k, events)
            return r + (None,) if defer else r
        kk = min(k, self.n_total)
        kc = min(kk + ops.DEFAULT_MARGIN, ops.MAX_CANDIDATES)   # rank of the global threshold
        kl = self.list_length(kc)                               # entries listed per shard
        cand_screen, cand_col = self._local_candidates(q, min(kk, kl), kl, events)
        # defer=True: everything after the screen (two all-gathers, threshold, re-rank of the owned
        # candidates, merge: no tensor cores, ~1 ms) runs on a side stream, i.e. UNDER the screen
        # of the next batch when searches are queued back to back
        side = None
        if defer and q.is_cuda:
            cur = torch.cuda.current_stream()
            screened = torch.cuda.Event()
            screened.record(cur)
            side = self.local.tail_stream()
            side.wait_event(screened)
            for t in (q, cand_screen, cand_col):
                t.record_stream(side)
        with (torch.cuda.stream(side) if side is not None else _NoStream()):
            thr = self._global_threshold(self._gather(cand_screen), kc)
            packed = self._rerank_owned(q, kk, cand_screen, cand_col, thr)
            ms, mi, unc_rows, n_unc = self._merge_certified(self._gather(packed), thr, kk)
            if side is not None:
                for t in (ms, mi, unc_rows, n_unc):
                    t.record_stream(cur)

        def fixup(n_bad):
            self.stats["rows"] = self.stats.get("rows", 0) + q.size(0)
            self.stats["resolved_locally_exact"] = self.stats.get("resolved_locally_exact", 0) + n_bad
            if kl < kc and n_bad > max(2, q.size(0) // 50):
                self.reduce_lists = False   # n_bad is the same on every rank: so is the switch
            if n_bad:
                # rows the global certificate rejects (the same on every rank): every shard answers
                # them with its own certified search (fp32-grade re-screen / exhaustive as needed),
                # plain merge
                rows = unc_rows[:n_bad].long().sort().values   # same order on every rank
                rs, ri = self._search_replicated_rerank(q.index_select(0, rows), k, None)
                ms.index_copy_(0, rows, rs)
                mi.index_copy_(0, rows, ri)

        # the one 4-byte D2H read of the exactness guarantee (deferred: off the critical path)
        with (torch.cuda.stream(side) if side is not None else _NoStream()):
            ticket = self._ticket(n_unc, fixup)
        if defer:
            return ms, mi, ticket
        ticket.resolve()
        return ms, mi

    def _ticket(self, n_unc, fixup):
        return ops.ExactnessTicket(n_unc, fixup)

    def _search_replicated_rerank(self, q, k, events=None):
        k_local = min(k, self.hi - self.lo)
        s, i = self._local_search(q, k_local, events)
        if k_local < k:  # a shard smaller than k: pad with invalid entries
            pad = k - k_local
            s = torch.cat([s, s.new_full((s.size(0), pad), float("-inf"))], 1)
            i = torch.cat([i, i.new_full((i.size(0), pad), -1)], 1)
        # the one exchange step of this variant: k results per query per shard
        ms, mi = self._merge(self._gather(s), self._gather(i))
        kk = min(k, self.n_total)
        return ms[:, :kk], mi[:, :kk]
