"""The whole C-ABI surface as torch custom ops ``torch.ops.isb.*`` (north_star: "exposed as
torch custom ops over a thin C-ABI extension").

One op per entry point / per public stage of the path, CUDA-only: no CPU kernel is
registered, so a CPU tensor raises ``NotImplementedError`` from the dispatcher (no
fallback).  Every op has a ``register_fake`` shape function, so the ops trace under
FakeTensorMode / torch.export without a GPU; ``tests/test_torch_ops_cpu.py`` checks the
registrations and the shape functions, ``tests/test_gpu_torch_ops.py`` calls every op on
the device through ``torch.ops.isb`` and compares it with the oracle.

Tensor-core operands are the bf16 terms made by ``isb::f32_to_bf16`` (part 0 = hi,
part 1 = lo); an op that may work without a term takes it as ``Optional[Tensor]``.
"""

from typing import Optional, Tuple

import torch
from torch.library import custom_op

from . import mining as _mining
from . import ops as _ops
from . import regions as _regions

T = torch.Tensor
OPS = []   # names of every registered op (tests iterate over it)


def _op(name, fake):
    """custom_op(..., device_types='cuda') + register_fake in one decorator."""
    def deco(fn):
        op = custom_op("isb::" + name, mutates_args=(), device_types="cuda")(fn)
        op.register_fake(fake)
        OPS.append(name)
        return op
    return deco


def _head_from_parts(cls_w, cls_w_hi, cls_w_lo, cls_b, shift, lin_w_hi, lin_w_lo, lin_b, cls_w_absmax=0.0):
    """A regions.HeadWeights over already converted operands (no conversion work)."""
    hw = object.__new__(_regions.HeadWeights)
    hw.terms = 1 if lin_w_lo is None else 3
    hw.cls_w, hw.cls_b, hw.cls_w_hi, hw.cls_w_lo = cls_w, cls_b, cls_w_hi, cls_w_lo
    hw.cls_w_absmax = float(cls_w_absmax)
    hw.shift, hw.lin_b, hw.lin_w_hi, hw.lin_w_lo = shift, lin_b, lin_w_hi, lin_w_lo
    if lin_w_hi is not None:
        hw.D, hw.KinP = lin_w_hi.shape
    else:
        hw.D, hw.KinP = 0, (0 if shift is None else (shift.numel() + 7) // 8 * 8)
    hw.Kin = shift.numel() if shift is not None else hw.KinP
    return hw


# ---------------------------------------------------------------- row operators (a1, a2, f3)
@_op("l2norm_rows", lambda x, eps: torch.empty_like(x))
def l2norm_rows(x: T, eps: float) -> T:
    return _ops.l2norm_rows(x, eps)


@_op("l2norm_rows_backward", lambda x, grad_out, eps: torch.empty_like(x))
def l2norm_rows_backward(x: T, grad_out: T, eps: float) -> T:
    return _ops.l2norm_rows_backward(x, grad_out, eps)


@_op("shift_rows", lambda x, param: torch.empty_like(x))
def shift_rows(x: T, param: T) -> T:
    return _ops.shift_rows(x, param)


@_op("col_sums", lambda g: g.new_empty((g.size(1),)))
def col_sums(g: T) -> T:
    return _ops.col_sums(g)


def _bf16_fake(x, part, ld):
    ld = ld if ld > 0 else (x.size(1) + 7) // 8 * 8
    return x.new_empty((x.size(0), ld), dtype=torch.bfloat16)


@_op("f32_to_bf16", _bf16_fake)
def f32_to_bf16(x: T, part: int, ld: int) -> T:
    return _ops.to_bf16(x, part, ld if ld > 0 else None)


# ---------------------------------------------------------------- dense contractions
@_op("gemm_nt", lambda a, b, splits: a.new_empty((a.size(0), b.size(0)), dtype=torch.float32))
def gemm_nt(a: T, b: T, splits: int) -> T:
    return _ops.gemm_nt(a, b, None, splits)


@_op("gemm_nt_split",
     lambda a_hi, a_lo, b_hi, b_lo, splits: a_hi.new_empty((a_hi.size(0), b_hi.size(0)), dtype=torch.float32))
def gemm_nt_split(a_hi: T, a_lo: T, b_hi: T, b_lo: T, splits: int) -> T:
    return _ops.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, None, splits)


@_op("all_pairs_similarities", lambda emb, terms: emb.new_empty((emb.size(0), emb.size(0))))
def all_pairs_similarities(emb: T, terms: int) -> T:
    return _mining.all_pairs_similarities(emb, terms)


# ---------------------------------------------------------------- search (a8-a10, 8e)
def _topk_fake(q, db_f32, db_bf16, k, margin, idx_offset):
    return q.new_empty((q.size(0), k)), q.new_empty((q.size(0), k), dtype=torch.int64)


@_op("topk_search", _topk_fake)
def topk_search(q: T, db_f32: T, db_bf16: T, k: int, margin: int, idx_offset: int) -> Tuple[T, T]:
    return _ops.topk_search(q, db_f32, db_bf16, k, margin if margin >= 0 else None, idx_offset)


def _cand_fake(q, db_bf16, k, kc):
    return q.new_empty((q.size(0), kc)), q.new_empty((q.size(0), kc), dtype=torch.int32)


@_op("topk_candidates", _cand_fake)
def topk_candidates(q: T, db_bf16: T, k: int, kc: int) -> Tuple[T, T]:
    return _ops.topk_candidates(q, db_bf16, k, kc)


@_op("topk_global_threshold", lambda all_screen, kth=0: all_screen.new_empty((all_screen.size(1),)))
def topk_global_threshold(all_screen: T, kth: int = 0) -> T:
    return _ops.topk_global_threshold(all_screen, kth)


@_op("topk_rerank_owned",
     lambda q, db_f32, k, cand_screen, cand_col, thr: q.new_empty((q.size(0), 2 * k + 2), dtype=torch.int32))
def topk_rerank_owned(q: T, db_f32: T, k: int, cand_screen: T, cand_col: T, thr: T) -> T:
    return _ops.topk_rerank_owned(q, db_f32, k, cand_screen, cand_col, thr)


def _mergec_fake(packed_all, row_offsets, thr, k):
    Q = packed_all.size(1)
    return (thr.new_empty((Q, k)), thr.new_empty((Q, k), dtype=torch.int64),
            thr.new_empty((max(Q, 1),), dtype=torch.int32), thr.new_empty((1,), dtype=torch.int32))


@_op("topk_merge_certified", _mergec_fake)
def topk_merge_certified(packed_all: T, row_offsets: T, thr: T, k: int) -> Tuple[T, T, T, T]:
    return _ops.topk_merge_certified(packed_all, row_offsets, thr, k)


@_op("topk_merge", lambda cs, ci: (cs.new_empty(cs.shape[1:]), ci.new_empty(ci.shape[1:])))
def topk_merge(cand_scores: T, cand_idx: T) -> Tuple[T, T]:
    return _ops.topk_merge(cand_scores, cand_idx)


# ---------------------------------------------------------------- region-descriptor head (a3-a6)
def _rsel_fake(x, cls_w, cls_w_hi, cls_w_lo, cls_b, fh, fw, k, margin, exact_mode):
    B, ncls = x.size(0), cls_w.size(0)
    f = lambda *s: x.new_empty(s)                                                   # noqa: E731
    return (x.new_empty((B, k), dtype=torch.int64), x.new_empty((B,), dtype=torch.int32), f(B, ncls, k), f(B, k),
            f(B, k), f(B), x.new_empty((1 + B,), dtype=torch.int32))


@_op("region_select", _rsel_fake)
def region_select(x: T, cls_w: T, cls_w_hi: T, cls_w_lo: T, cls_b: T, fh: int, fw: int, k: int, margin: int,
                  exact_mode: bool) -> Tuple[T, T, T, T, T, T, T]:
    hw = _head_from_parts(cls_w, cls_w_hi, cls_w_lo, cls_b, None, None, None, None)
    return _regions.region_select(x, hw, k, (fh, fw), margin, exact_mode)


def _rgather_fake(x, idx, nsel, win_norm, shift, fh, fw, k_sum, terms):
    B, C, k = x.size(0), x.size(1), idx.size(1)
    KinP = (C * fh * fw + 7) // 8 * 8
    u = x.new_empty((B, KinP), dtype=torch.bfloat16)
    return u, (torch.empty_like(u) if terms == 3 else x.new_empty((0,), dtype=torch.bfloat16)), x.new_empty((B, k, C))


@_op("region_gather", _rgather_fake)
def region_gather(x: T, idx: T, nsel: T, win_norm: T, shift: T, fh: int, fw: int, k_sum: int,
                  terms: int) -> Tuple[T, T, T]:
    hw = _head_from_parts(None, None, None, None, shift, None, None, None)
    hw.terms, hw.Kin = terms, x.size(1) * fh * fw
    hw.KinP = (hw.Kin + 7) // 8 * 8
    U_hi, U_lo, win_mean = _regions.region_gather(x, hw, idx.size(1), (fh, fw), idx, nsel, win_norm, k_sum=k_sum)
    return U_hi, (U_lo if U_lo is not None else x.new_empty((0,), dtype=torch.bfloat16)), win_mean


def _rlogits_fake(win_mean, cls_w, cls_b, k, nsel, idx, norm, approx_max, runner_up, cls_w_absmax):
    B, ncls = win_mean.size(0), cls_w.size(0)
    i32 = lambda *s: win_mean.new_empty(s, dtype=torch.int32)                        # noqa: E731
    return (win_mean.new_empty((B, k), dtype=torch.int64), win_mean.new_empty((B, k)), i32(B),
            win_mean.new_empty((B, ncls, k)), i32(B), i32(1), i32(1 + B))


@_op("region_logits", _rlogits_fake)
def region_logits(win_mean: T, cls_w: T, cls_b: T, k: int, nsel: T, idx: T, norm: T, approx_max: T, runner_up: T,
                  cls_w_absmax: float) -> Tuple[T, T, T, T, T, T, T]:
    hw = _head_from_parts(cls_w, None, None, cls_b, None, None, None, None, cls_w_absmax)
    return _regions.region_logits(win_mean, hw, k, nsel, idx, norm, approx_max, runner_up)


@_op("descriptor_finalize", lambda y, bias, nsel, eps: torch.empty_like(y))
def descriptor_finalize(y: T, bias: Optional[T], nsel: Optional[T], eps: float) -> T:
    return _regions.descriptor_finalize(y, bias, nsel, eps)


def _rdesc_fake(x, cls_w, cls_w_hi, cls_w_lo, cls_b, shift, lin_w_hi, lin_w_lo, lin_b, fh, fw, k, cls_w_absmax):
    B, ncls, D = x.size(0), cls_w.size(0), lin_w_hi.size(0)
    return (x.new_empty((B, D)), x.new_empty((B, ncls, k)), x.new_empty((B, k), dtype=torch.int64),
            x.new_empty((B,), dtype=torch.int32))


@_op("region_descriptors", _rdesc_fake)
def region_descriptors(x: T, cls_w: T, cls_w_hi: T, cls_w_lo: T, cls_b: T, shift: T, lin_w_hi: T,
                       lin_w_lo: Optional[T], lin_b: Optional[T], fh: int, fw: int, k: int,
                       cls_w_absmax: float) -> Tuple[T, T, T, T]:
    """The whole head of RegionDescriptorNet.forward_single after the trunk
    (model/siamese.py:187-222) for a batch of feature maps: (desc, cls_out, idx, nsel)."""
    hw = _head_from_parts(cls_w, cls_w_hi, cls_w_lo, cls_b, shift, lin_w_hi, lin_w_lo, lin_b, cls_w_absmax)
    return _regions.region_descriptors(x, hw, k, (fh, fw))


@_op("global_descriptors", lambda x, shift, lin_w_hi, lin_w_lo, lin_b: x.new_empty((x.size(0), lin_w_hi.size(0))))
def global_descriptors(x: T, shift: T, lin_w_hi: T, lin_w_lo: Optional[T], lin_b: Optional[T]) -> T:
    """DescriptorNet head (model/siamese.py:117-122)."""
    hw = _head_from_parts(None, None, None, None, shift, lin_w_hi, lin_w_lo, lin_b)
    return _regions.global_descriptors(x, hw)


def _cstats_fake(x, idx, nsel, g_u, fh, fw):
    B, C, k = x.size(0), x.size(1), idx.size(1)
    return x.new_empty((B, k)), x.new_empty((B, k)), x.new_empty((B, k, C))


@_op("region_crop_stats", _cstats_fake)
def region_crop_stats(x: T, idx: T, nsel: T, g_u: Optional[T], fh: int, fw: int) -> Tuple[T, T, T]:
    """Backward glue of the fused head: (|crop|^2, <crop, g_u>, window means) per selected window."""
    return _regions.crop_stats(x, idx.size(1), (fh, fw), idx, nsel, g_u, want_means=True)


@_op("region_scatter_grad", lambda x, idx, nsel, g_u, n2, dot, g_mean, fh, fw, eps: torch.empty_like(x))
def region_scatter_grad(x: T, idx: T, nsel: T, g_u: Optional[T], n2: T, dot: T, g_mean: Optional[T], fh: int,
                        fw: int, eps: float) -> T:
    """Backward glue of the fused head: gradient w.r.t. the feature map (gather form)."""
    return _regions.scatter_grad(x, idx.size(1), (fh, fw), idx, nsel, g_u, n2, dot, g_mean, eps)


# ---------------------------------------------------------------- mining (a11-a13)
def _neg_fake(emb, labels, anchors, positives, semi_hard, terms):
    P = anchors.size(0)
    return emb.new_empty((P,), dtype=torch.int64), emb.new_empty((P,)), emb.new_empty((P,))


@_op("select_negatives", _neg_fake)
def select_negatives(emb: T, labels: T, anchors: T, positives: T, semi_hard: bool, terms: int) -> Tuple[T, T, T]:
    """One (semi-)hard negative per positive couple (train/siamese_regions.py:106-129).  Builds the
    bf16 operands of ``emb`` on every call: keep a mining.MiningIndex for repeated use."""
    return _mining.MiningIndex(emb, labels, terms=terms).select_negatives(anchors, positives, semi_hard)


# ---------------------------------------------------------------- metrics / DBA (a9, a10, f1, f2)
@_op("row_kth_largest",
     lambda sim, kth: (sim.new_empty((sim.size(0),)), sim.new_empty((sim.size(0),), dtype=torch.int64)))
def row_kth_largest(sim: T, kth: int) -> Tuple[T, T]:
    return _ops.row_kth_largest(sim, kth)


@_op("row_ranks", lambda sim, cols: cols.new_empty(cols.shape, dtype=torch.int32))
def row_ranks(sim: T, cols: T) -> T:
    return _ops.row_ranks(sim, cols)


@_op("instance_avg", lambda emb, label_ids, k: torch.empty_like(emb))
def instance_avg(emb: T, label_ids: T, k: int) -> T:
    return _ops.instance_avg(emb, label_ids, k)


# ---------------------------------------------------------------- training-side (f3)
@_op("triplet_loss_forward",
     lambda a, p, n, margin, size_average, normalized: (a.new_empty((1,)), a.new_empty((a.size(0),), dtype=torch.uint8)))
def triplet_loss_forward(anchor: T, pos: T, neg: T, margin: float, size_average: bool,
                         normalized: bool) -> Tuple[T, T]:
    return _ops.triplet_loss_forward(anchor, pos, neg, margin, size_average, normalized)


@_op("triplet_loss_backward",
     lambda a, p, n, clamp, g, size_average, normalized: (torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)))
def triplet_loss_backward(anchor: T, pos: T, neg: T, clamp: T, grad_out: T, size_average: bool,
                          normalized: bool) -> Tuple[T, T, T]:
    return _ops.triplet_loss_backward(anchor, pos, neg, clamp, grad_out, size_average, normalized)
