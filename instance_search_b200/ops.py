"""Tensor-level wrappers of the C ABI (ctypes -> libisb.so).  ``torch_ops.py`` registers them
as torch custom ops (``torch.ops.isb.*``).

PyTorch is plumbing here: it owns device memory and the current stream; every
computation below happens in libisb.so.  CUDA tensors only -- a CPU tensor is
an error, not a fallback.
"""

import torch

from . import _lib
from ._lib import IsbError

DEFAULT_MARGIN = 28      # candidates screened beyond k (k + margin <= 128)
MAX_CANDIDATES = 128


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise IsbError("instance_search_b200 ops take CUDA tensors only (got %s); "
                           "there is no CPU fallback" % (getattr(t, "device", type(t)),))


def _f32c(t):
    if t.dtype != torch.float32:
        raise IsbError("expected a float32 tensor, got %s" % t.dtype)
    return t.contiguous()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ---------------------------------------------------------------- row operators
def l2norm_rows(x, eps=1e-10):
    """y = x / sqrt(sum(x^2, 1) + eps). reference: model/custom_modules.py:52-57"""
    _need_cuda(x)
    if x.dim() != 2:
        raise IsbError("l2norm_rows expects a 2-D tensor")
    x = _f32c(x)
    y = torch.empty_like(x)
    _lib.check(_lib.lib().isb_l2norm_rows(x.data_ptr(), x.size(0), x.size(1), float(eps),
                                          y.data_ptr(), _stream()), "isb_l2norm_rows")
    return y


def shift_rows(x, param):
    """y = x + param. reference: model/custom_modules.py:16-18"""
    _need_cuda(x, param)
    if x.dim() != 2 or param.numel() != x.size(1):
        raise IsbError("shift_rows: x [M, F] and param [F] expected")
    x, param = _f32c(x), _f32c(param)
    y = torch.empty_like(x)
    _lib.check(_lib.lib().isb_shift_rows(x.data_ptr(), param.data_ptr(), x.size(0), x.size(1),
                                         y.data_ptr(), _stream()), "isb_shift_rows")
    return y


def to_bf16(x, part=0, ld=None):
    """fp32 [rows, cols] -> bf16 [rows, ld] (ld = cols rounded up to 8, zero padded).
    part 0/1/2 = hi / lo / lo2 term of the bf16 expansion of x."""
    _need_cuda(x)
    x = _f32c(x)
    rows, cols = x.shape
    if ld is None:
        ld = (cols + 7) // 8 * 8
    y = torch.empty((rows, ld), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().isb_f32_to_bf16(x.data_ptr(), rows, cols, cols, y.data_ptr(), ld,
                                          int(part), _stream()), "isb_f32_to_bf16")
    return y


def gemm_nt(a, b, bias=None, splits=1, k=None):
    """C[M, N] fp32 = a[M, :K] . b[N, :K]^T (+ bias). a, b: bf16, row-major,
    leading dimensions multiples of 8."""
    _need_cuda(a, b, bias)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise IsbError("gemm_nt takes bf16 operands (see to_bf16)")
    a, b = a.contiguous(), b.contiguous()
    K = min(a.size(1), b.size(1)) if k is None else k
    M, N = a.size(0), b.size(0)
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    L = _lib.lib()
    ws_bytes = L.isb_gemm_nt_workspace_bytes(M, N, K, splits)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=a.device)
    if bias is not None:
        bias = _f32c(bias)
    _lib.check(L.isb_gemm_nt(a.data_ptr(), a.size(1), b.data_ptr(), b.size(1), M, N, K,
                             _ptr(bias), c.data_ptr(), N, int(splits), ws.data_ptr(), ws_bytes,
                             _stream()), "isb_gemm_nt")
    return c


def gemm_nt_split(a_hi, a_lo, b_hi, b_lo, bias=None, splits=1, k=None):
    """fp32-grade C[M, N] = A . B^T from split operands (to_bf16 parts 0 and 1):
    a_hi.b_hi + a_lo.b_hi + a_hi.b_lo accumulated in one TMEM tile."""
    _need_cuda(a_hi, a_lo, b_hi, b_lo, bias)
    for t in (a_hi, a_lo, b_hi, b_lo):
        if t.dtype != torch.bfloat16 or not t.is_contiguous():
            raise IsbError("gemm_nt_split takes contiguous bf16 operands (see to_bf16)")
    if a_hi.shape != a_lo.shape or b_hi.shape != b_lo.shape:
        raise IsbError("gemm_nt_split: hi/lo shape mismatch")
    K = min(a_hi.size(1), b_hi.size(1)) if k is None else k
    M, N = a_hi.size(0), b_hi.size(0)
    c = torch.empty((M, N), dtype=torch.float32, device=a_hi.device)
    L = _lib.lib()
    ws_bytes = L.isb_gemm_nt_workspace_bytes(M, N, K, splits)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=a_hi.device)
    if bias is not None:
        bias = _f32c(bias)
    _lib.check(L.isb_gemm_nt_split(a_hi.data_ptr(), a_lo.data_ptr(), a_hi.size(1), b_hi.data_ptr(),
                                   b_lo.data_ptr(), b_hi.size(1), M, N, K, _ptr(bias), c.data_ptr(), N,
                                   int(splits), ws.data_ptr(), ws_bytes, _stream()), "isb_gemm_nt_split")
    return c


# ---------------------------------------------------------------------- search
class ExactnessTicket(object):
    """The host side of the exactness guarantee, deferred.  The completeness certificate leaves
    the number of rows it rejected in device memory; reading it is the one device->host
    round trip of a search.  A ticket holds that count on its way to pinned host memory (copy
    queued behind the search, an event after it) so that the caller can queue MORE work --
    the next batch's search -- before it waits: ``resolve()`` waits for the event and, on the
    rare batch that has uncertified rows, re-runs them (fp32-grade re-screen / exhaustive pass)
    and patches the result tensors in place.  A result must not be consumed before its ticket is
    resolved."""

    def __init__(self, n_unc, fixup):
        if n_unc.is_cuda:
            self._host = torch.empty(1, dtype=torch.int32).pin_memory()
            self._host.copy_(n_unc, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()
        else:   # host tensors (the gloo plumbing tests): nothing to wait for
            self._host, self._event = n_unc.reshape(-1)[:1].clone(), None
        self._fixup = fixup
        self.n_uncertified = None

    def resolve(self):
        if self.n_uncertified is None:
            if self._event is not None:
                self._event.synchronize()
            self.n_uncertified = int(self._host[0])
            self._fixup(self.n_uncertified)
            self._fixup = None
        return self.n_uncertified


def resolve_uncertified(q, db_f32, db_bf16, db_lo, k, margin, idx_offset, scores, idx, unc_rows, n_unc,
                        stats=None, defer=False):
    """Host side of the exactness guarantee (include/isb.h, stage 2b): read the
    number of rows the bf16 screen could not certify (ONE 4-byte D2H read; it is 0
    on ordinary data), re-screen those with fp32-grade split operands, and search
    exhaustively whatever is still uncertified.  db_lo: bf16 lo term of db_f32 or a
    zero-argument callable that builds it on first need.  defer=True: nothing is read
    here; returns an ExactnessTicket to resolve later."""
    L = _lib.lib()
    N, D = db_f32.shape

    def fixup(n1):
        n2 = 0
        if n1 > 0:
            lo = db_lo() if callable(db_lo) else db_lo
            rows = unc_rows[:n1].clone()
            nb = L.isb_topk_resolve_workspace_bytes(n1, N, D)
            ws = torch.empty(nb, dtype=torch.uint8, device=q.device)
            _lib.check(L.isb_topk_resolve(q.data_ptr(), db_f32.data_ptr(), db_bf16.data_ptr(), lo.data_ptr(),
                                          N, D, db_bf16.size(1), int(k), int(margin), int(idx_offset),
                                          rows.data_ptr(), n1, scores.data_ptr(), idx.data_ptr(),
                                          unc_rows.data_ptr(), n_unc.data_ptr(), ws.data_ptr(), nb, _stream()),
                       "isb_topk_resolve")
            n2 = int(n_unc.item())
            if n2 > 0:
                rows = unc_rows[:n2].clone()
                nb = L.isb_topk_exhaustive_workspace_bytes(n2, N, int(k))
                ws = torch.empty(nb, dtype=torch.uint8, device=q.device)
                _lib.check(L.isb_topk_exhaustive(q.data_ptr(), db_f32.data_ptr(), N, D, int(k), int(idx_offset),
                                                 rows.data_ptr(), n2, scores.data_ptr(), idx.data_ptr(),
                                                 ws.data_ptr(), nb, _stream()), "isb_topk_exhaustive")
        if stats is not None:
            stats["rows"] = stats.get("rows", 0) + q.size(0)
            stats["resolved_fp32_grade"] = stats.get("resolved_fp32_grade", 0) + n1
            stats["resolved_exhaustive"] = stats.get("resolved_exhaustive", 0) + n2

    ticket = ExactnessTicket(n_unc, fixup)
    if defer:
        return ticket
    ticket.resolve()
    return None


def topk_search(q, db_f32, db_bf16, k, margin=None, idx_offset=0, workspace=None, exact=True,
                db_lo=None, stats=None):
    """Top-k rows of db by q . db (best first), index-exact w.r.t. fp64 scores.

    replaces ``torch.mm(q, db.t())`` + sort/max (test/siamese_regions_test.py:76,
    utils/metrics.py:11,13,33).  q [Q, D] fp32, db_f32 [N, D] fp32,
    db_bf16 = to_bf16(db_f32).  Returns (scores [Q, k] fp32, idx [Q, k] int64).
    exact=True (default) runs the completeness certificate and resolves the rows
    that fail it (one 4-byte device->host read per call); exact=False returns the
    re-ranked screen result without the guarantee and without any host sync.
    """
    _need_cuda(q, db_f32, db_bf16)
    q, db_f32 = _f32c(q), _f32c(db_f32)
    Q, D = q.shape
    N = db_f32.size(0)
    if db_f32.size(1) != D or db_bf16.size(0) != N or db_bf16.dtype != torch.bfloat16:
        raise IsbError("topk_search: shape/dtype mismatch between q, db_f32 and db_bf16")
    if margin is None:
        margin = min(DEFAULT_MARGIN, MAX_CANDIDATES - k)
    margin = max(0, min(margin, MAX_CANDIDATES - k))
    scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
    if Q == 0:
        return scores, idx
    L = _lib.lib()
    ws_bytes = L.isb_topk_search_workspace_bytes(Q, N, D, k, margin)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    unc_rows = torch.empty(Q, dtype=torch.int32, device=q.device) if exact else None
    n_unc = torch.zeros(1, dtype=torch.int32, device=q.device) if exact else None
    _lib.check(L.isb_topk_search(q.data_ptr(), Q, db_f32.data_ptr(), db_bf16.data_ptr(), N, D,
                                 db_bf16.size(1), int(k), int(margin), int(idx_offset),
                                 scores.data_ptr(), idx.data_ptr(), _ptr(unc_rows), _ptr(n_unc),
                                 workspace.data_ptr(), workspace.numel(), _stream()), "isb_topk_search")
    if exact:
        if db_lo is None:
            db_lo = lambda: to_bf16(db_f32, 1, ld=db_bf16.size(1))  # noqa: E731
        resolve_uncertified(q, db_f32, db_bf16, db_lo, k, margin, idx_offset, scores, idx, unc_rows,
                            n_unc, stats)
    return scores, idx


def topk_search_workspace(Q, N, D, k, margin, device):
    n = _lib.lib().isb_topk_search_workspace_bytes(Q, N, D, k, margin)
    return torch.empty(n, dtype=torch.uint8, device=device)


def topk_candidates(q, db_bf16, k, kc, workspace=None, events=None, dim=None):
    """Sharded search, local stage 1 (isb_topk_screen + isb_topk_candidates): screen one shard and
    list its kc best screen entries per query.  q [Q, D] fp32, db_bf16 [N, ld] (to_bf16 of the
    shard); dim = the true D when the bf16 rows are padded.  Returns (cand_screen [Q, kc] fp32,
    cand_col [Q, kc] int32 local row; -inf / -1 where the shard has fewer than kc rows)."""
    _need_cuda(q, db_bf16)
    q = _f32c(q)
    Q, D = q.shape
    N = db_bf16.size(0)
    k_eff = min(k, N)
    margin = min(kc, N) - k_eff
    cand_screen = torch.empty((Q, kc), dtype=torch.float32, device=q.device)
    cand_col = torch.empty((Q, kc), dtype=torch.int32, device=q.device)
    if Q == 0:
        return cand_screen, cand_col
    L = _lib.lib()
    nbytes = L.isb_topk_search_workspace_bytes(Q, N, D, k_eff, margin)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    if events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(L.isb_topk_screen(q.data_ptr(), Q, db_bf16.data_ptr(), N, D, db_bf16.size(1), k_eff, margin,
                                 workspace.data_ptr(), workspace.numel(), _stream()), "isb_topk_screen")
    if events is not None:
        e1.record()
        events.append((e0, e1))
    _lib.check(L.isb_topk_candidates(Q, N, D, k_eff, margin, kc, cand_screen.data_ptr(), cand_col.data_ptr(),
                                     workspace.data_ptr(), workspace.numel(), _stream()), "isb_topk_candidates")
    return cand_screen, cand_col


def topk_rerank_owned(q, db_f32, k, cand_screen, cand_col, thr):
    """Sharded search, local stage 2 (isb_topk_rerank_owned): exact scores of this shard's
    candidates at or above the global threshold, best first, as packed rows [Q, 2k + 2] int32
    words: k scores (fp32 bits, -inf padded) | k local rows (-1 padded) | sum (screen - exact)^2 |
    candidates scored (fp32 bits) -- what the second all-gather exchanges."""
    _need_cuda(q, db_f32, cand_screen, cand_col, thr)
    q, db_f32 = _f32c(q), _f32c(db_f32)
    Q, kc = cand_screen.shape
    N, D = db_f32.shape
    packed = torch.empty((Q, 2 * k + 2), dtype=torch.int32, device=q.device)
    if Q == 0:
        return packed
    _lib.check(_lib.lib().isb_topk_rerank_owned(
        q.data_ptr(), Q, db_f32.data_ptr(), N, D, k, kc, _f32c(cand_screen).data_ptr(),
        cand_col.contiguous().data_ptr(), _f32c(thr).data_ptr(), packed.data_ptr(), _stream()),
        "isb_topk_rerank_owned")
    return packed


def topk_merge(cand_scores, cand_idx):
    """[R, Q, k] per-shard results -> the k best per query (ties -> lower index)."""
    _need_cuda(cand_scores, cand_idx)
    cand_scores = _f32c(cand_scores)
    cand_idx = cand_idx.contiguous()
    if cand_idx.dtype != torch.int64 or cand_scores.shape != cand_idx.shape or cand_idx.dim() != 3:
        raise IsbError("topk_merge: expected [R, Q, k] fp32 scores and int64 indices")
    R, Q, k = cand_scores.shape
    scores = torch.empty((Q, k), dtype=torch.float32, device=cand_scores.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=cand_scores.device)
    _lib.check(_lib.lib().isb_topk_merge(cand_scores.data_ptr(), cand_idx.data_ptr(), R, Q, k,
                                         scores.data_ptr(), idx.data_ptr(), _stream()),
               "isb_topk_merge")
    return scores, idx


def topk_global_threshold(all_screen, kth=0):
    """all_screen [R, Q, kc] fp32 (the all-gathered screen scores of every shard's
    candidates, -inf = none) -> thr [Q]: the kth-best per query (kth <= 0: kc), -inf when
    fewer exist."""
    _need_cuda(all_screen)
    all_screen = _f32c(all_screen)
    R, Q, kc = all_screen.shape
    thr = torch.empty((Q,), dtype=torch.float32, device=all_screen.device)
    _lib.check(_lib.lib().isb_topk_global_threshold(all_screen.data_ptr(), R, Q, kc, int(kth), thr.data_ptr(), _stream()),
               "isb_topk_global_threshold")
    return thr


def topk_merge_certified(packed_all, row_offsets, thr, k):
    """The k best of the shards' packed re-ranked lists + the global completeness
    certificate (include/isb.h).  packed_all [R, Q, 2k + 2] (32-bit words, any 4-byte
    dtype), row_offsets [R] int64 (first global row of every shard), thr [Q].  Returns
    (scores [Q, k], idx [Q, k] int64 global, uncertified_rows [Q] int32, n_uncertified [1])."""
    _need_cuda(packed_all, row_offsets, thr)
    packed_all = packed_all.contiguous()
    R, Q, pw = packed_all.shape
    if packed_all.element_size() != 4 or pw != 2 * k + 2:
        raise IsbError("topk_merge_certified: expected packed_all [R, Q, 2k + 2] of 32-bit words")
    if row_offsets.dtype != torch.int64 or row_offsets.numel() != R or tuple(thr.shape) != (Q,):
        raise IsbError("topk_merge_certified: expected row_offsets [R] int64 and thr [Q]")
    thr = _f32c(thr)
    dev = packed_all.device
    scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
    unc_rows = torch.empty((max(Q, 1),), dtype=torch.int32, device=dev)
    n_unc = torch.empty((1,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().isb_topk_merge_certified(packed_all.data_ptr(), row_offsets.contiguous().data_ptr(),
                                                   thr.data_ptr(), R, Q, int(k), scores.data_ptr(),
                                                   idx.data_ptr(), unc_rows.data_ptr(), n_unc.data_ptr(),
                                                   _stream()), "isb_topk_merge_certified")
    return scores, idx, unc_rows, n_unc


# --------------------------------------------------------------------- metrics
def row_kth_largest(sim, kth=1):
    """(val [Q], idx [Q] int64): the kth largest entry of every row of sim and its
    column -- sim.max(1) / sim.kthvalue(N - kth + 1, 1) of utils/metrics.py:11-13."""
    _need_cuda(sim)
    if sim.dim() != 2 or sim.dtype != torch.float32 or sim.stride(1) != 1:
        raise IsbError("row_kth_largest: [Q, N] float32 rows expected")
    Q, N = sim.shape
    val = torch.empty(Q, dtype=torch.float32, device=sim.device)
    idx = torch.empty(Q, dtype=torch.int64, device=sim.device)
    _lib.check(_lib.lib().isb_row_kth_largest(sim.data_ptr(), Q, N, sim.stride(0), int(kth), val.data_ptr(),
                                              idx.data_ptr(), _stream()), "isb_row_kth_largest")
    return val, idx


def row_ranks(sim, cols):
    """rank [Q, P] int32 of the listed columns (cols [Q, P] int32, -1 = unused) in
    the descending sort of their row -- utils/metrics.py:33 without the sort."""
    _need_cuda(sim, cols)
    if sim.dim() != 2 or sim.dtype != torch.float32 or sim.stride(1) != 1:
        raise IsbError("row_ranks: [Q, N] float32 rows expected")
    cols = cols.to(torch.int32).contiguous()
    Q, N = sim.shape
    if cols.dim() != 2 or cols.size(0) != Q:
        raise IsbError("row_ranks: cols must be [Q, P]")
    rank = torch.empty_like(cols)
    _lib.check(_lib.lib().isb_row_ranks(sim.data_ptr(), Q, N, sim.stride(0), cols.data_ptr(), cols.size(1),
                                        rank.data_ptr(), _stream()), "isb_row_ranks")
    return rank


def instance_avg(emb, label_ids, k=-1):
    """DBA of test/instance_avg.py:7-33 on device; label_ids [N] int32."""
    _need_cuda(emb, label_ids)
    emb = _f32c(emb)
    label_ids = label_ids.to(torch.int32).contiguous()
    out = torch.empty_like(emb)
    ovf = torch.zeros(1, dtype=torch.int32, device=emb.device)
    _lib.check(_lib.lib().isb_instance_avg(emb.data_ptr(), label_ids.data_ptr(), emb.size(0), emb.size(1),
                                           int(k), out.data_ptr(), ovf.data_ptr(), _stream()),
               "isb_instance_avg")
    if int(ovf.item()):
        raise IsbError("instance_avg: an instance has more than 2048 other members")
    return out


# ------------------------------------------------------------- training-side ops
def l2norm_rows_backward(x, grad_out, eps=1e-10):
    """reference: model/custom_modules.py:59-67"""
    _need_cuda(x, grad_out)
    x, grad_out = _f32c(x), _f32c(grad_out)
    gx = torch.empty_like(x)
    _lib.check(_lib.lib().isb_l2norm_rows_backward(x.data_ptr(), grad_out.data_ptr(), x.size(0), x.size(1),
                                                   float(eps), gx.data_ptr(), _stream()),
               "isb_l2norm_rows_backward")
    return gx


def col_sums(g):
    """grad_param of Shift: model/custom_modules.py:23-24"""
    _need_cuda(g)
    g = _f32c(g)
    out = torch.empty(g.size(1), dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().isb_col_sums(g.data_ptr(), g.size(0), g.size(1), out.data_ptr(), _stream()),
               "isb_col_sums")
    return out


def triplet_loss_forward(anchor, pos, neg, margin, size_average=True, normalized=True):
    """(loss [1], clamp [B] uint8). reference: model/custom_modules.py:153-171"""
    _need_cuda(anchor, pos, neg)
    anchor, pos, neg = _f32c(anchor), _f32c(pos), _f32c(neg)
    B, D = anchor.shape
    loss = torch.empty(1, dtype=torch.float32, device=anchor.device)
    row_loss = torch.empty(B, dtype=torch.float32, device=anchor.device)
    clamp = torch.empty(B, dtype=torch.uint8, device=anchor.device)
    _lib.check(_lib.lib().isb_triplet_loss_forward(anchor.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, D,
                                                   float(margin), int(bool(size_average)),
                                                   int(bool(normalized)), loss.data_ptr(),
                                                   row_loss.data_ptr(), clamp.data_ptr(), _stream()),
               "isb_triplet_loss_forward")
    return loss, clamp


def triplet_loss_backward(anchor, pos, neg, clamp, grad_out, size_average=True, normalized=True):
    """reference: model/custom_modules.py:173-203"""
    _need_cuda(anchor, pos, neg, clamp, grad_out)
    anchor, pos, neg = _f32c(anchor), _f32c(pos), _f32c(neg)
    grad_out = _f32c(grad_out.reshape(1))
    B, D = anchor.shape
    ga, gp, gn = torch.empty_like(anchor), torch.empty_like(anchor), torch.empty_like(anchor)
    _lib.check(_lib.lib().isb_triplet_loss_backward(anchor.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, D,
                                                    clamp.data_ptr(), grad_out.data_ptr(),
                                                    int(bool(size_average)), int(bool(normalized)),
                                                    ga.data_ptr(), gp.data_ptr(), gn.data_ptr(), _stream()),
               "isb_triplet_loss_backward")
    return ga, gp, gn
