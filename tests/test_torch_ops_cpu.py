"""CPU: the torch custom-op surface (torch.ops.isb.*).  Every compute entry point of the C ABI
is reachable through an op; every op has a fake (shape) function that traces without a GPU;
no op has a CPU kernel (a CPU tensor is an error, not a fallback)."""

import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from instance_search_b200 import _lib, torch_ops

# C-ABI entry point -> the op(s) that reach it (queries / helpers that do no device work: None)
ENTRY_TO_OP = {
    "isb_abi_version": None, "isb_last_error": None, "isb_check_device": None,
    "isb_set_option": None, "isb_get_option": None,
    "isb_l2norm_rows": "l2norm_rows", "isb_shift_rows": "shift_rows", "isb_f32_to_bf16": "f32_to_bf16",
    "isb_topk_search_workspace_bytes": None, "isb_topk_search": "topk_search",
    "isb_topk_resolve_workspace_bytes": None, "isb_topk_resolve": "topk_search",
    "isb_topk_exhaustive_workspace_bytes": None, "isb_topk_exhaustive": "topk_search",
    "isb_topk_screen": "topk_candidates", "isb_topk_rerank": "topk_search",
    "isb_topk_merge": "topk_merge", "isb_topk_candidates": "topk_candidates",
    "isb_topk_global_threshold": "topk_global_threshold", "isb_topk_rerank_owned": "topk_rerank_owned",
    "isb_topk_merge_certified": "topk_merge_certified",
    "isb_region_select_workspace_bytes": None, "isb_region_select": "region_select",
    "isb_region_logits": "region_logits", "isb_region_gather": "region_gather",
    "isb_descriptor_finalize": "descriptor_finalize",
    "isb_region_crop_stats": "region_crop_stats", "isb_region_scatter_grad": "region_scatter_grad",
    "isb_select_negatives_workspace_bytes": None, "isb_select_negatives": "select_negatives",
    "isb_row_kth_largest": "row_kth_largest", "isb_row_ranks": "row_ranks", "isb_instance_avg": "instance_avg",
    "isb_l2norm_rows_backward": "l2norm_rows_backward", "isb_col_sums": "col_sums",
    "isb_triplet_loss_forward": "triplet_loss_forward", "isb_triplet_loss_backward": "triplet_loss_backward",
    "isb_gemm_nt_workspace_bytes": None, "isb_gemm_nt": "gemm_nt", "isb_gemm_nt_split": "gemm_nt_split",
}


def test_every_entry_point_is_reachable_through_an_op():
    missing = [e for e in _lib.SIGNATURES if e not in ENTRY_TO_OP]
    assert not missing, "C-ABI entry points without an op mapping: %s" % missing
    for entry, op in ENTRY_TO_OP.items():
        assert entry in _lib.SIGNATURES, entry
        if op is not None:
            assert op in torch_ops.OPS and hasattr(torch.ops.isb, op), (entry, op)
    # the composite ops of the path on top of the entry points
    for op in ("region_descriptors", "global_descriptors", "all_pairs_similarities"):
        assert op in torch_ops.OPS


def _f(*shape, dtype=torch.float32):
    return torch.empty(*shape, dtype=dtype, device="cuda")


def test_fake_shape_functions_trace_without_a_gpu():
    o = torch.ops.isb
    bf, i32, i64 = torch.bfloat16, torch.int32, torch.int64
    B, C, H, W, ncls, k, D, fh = 4, 64, 14, 14, 20, 6, 32, 7
    Kin = C * 49
    with FakeTensorMode():
        x = _f(5, 64)
        assert o.l2norm_rows(x, 1e-10).shape == (5, 64)
        assert o.l2norm_rows_backward(x, x, 1e-10).shape == (5, 64)
        assert o.shift_rows(x, _f(64)).shape == (5, 64) and o.col_sums(x).shape == (64,)
        h = o.f32_to_bf16(_f(5, 100), 0, 0)
        assert h.shape == (5, 104) and h.dtype == bf
        assert o.gemm_nt(_f(7, 64, dtype=bf), _f(9, 64, dtype=bf), 1).shape == (7, 9)
        assert o.gemm_nt_split(*([_f(7, 64, dtype=bf)] * 2 + [_f(9, 64, dtype=bf)] * 2), 1).shape == (7, 9)
        assert o.all_pairs_similarities(_f(11, 64), 3).shape == (11, 11)
        s, i = o.topk_search(_f(3, 64), _f(100, 64), _f(100, 64, dtype=bf), 10, -1, 0)
        assert s.shape == (3, 10) and i.dtype == i64
        cs, cc = o.topk_candidates(_f(3, 64), _f(100, 64, dtype=bf), 10, 38)
        assert cs.shape == (3, 38) and cc.dtype == i32
        thr = o.topk_global_threshold(_f(2, 3, 38))
        assert thr.shape == (3,)
        packed = o.topk_rerank_owned(_f(3, 64), _f(100, 64), 10, cs, cc, thr)
        assert packed.shape == (3, 22) and packed.dtype == i32
        ms, mi, ur, nu = o.topk_merge_certified(_f(2, 3, 22, dtype=i32), _f(2, dtype=i64), thr, 10)
        assert ms.shape == (3, 10) and mi.dtype == i64 and nu.shape == (1,)
        ms, mi = o.topk_merge(_f(2, 3, 10), _f(2, 3, 10, dtype=i64))
        assert ms.shape == (3, 10) and mi.shape == (3, 10)
        # region head
        fm = _f(B, C, H, W)
        cw, cwb = _f(ncls, C), _f(ncls, C, dtype=bf)
        idx, nsel, cls_out, wn, am, ru, nunc = o.region_select(fm, cw, cwb, cwb, _f(ncls), fh, fh, 8, 10, False)
        assert idx.shape == (B, 8) and cls_out.shape == (B, ncls, 8) and nunc.shape == (1 + B,)
        uh, ul, wm = o.region_gather(fm, idx, nsel, wn, _f(Kin), fh, fh, k, 3)
        assert uh.shape == (B, Kin) and uh.dtype == bf and ul.shape == uh.shape and wm.shape == (B, 8, C)
        r = o.region_logits(wm, cw, _f(ncls), k, nsel, idx, wn, am, ru, 1.0)
        assert r[0].shape == (B, k) and r[3].shape == (B, ncls, k)
        assert o.descriptor_finalize(_f(B, D), _f(D), nsel, 1e-10).shape == (B, D)
        assert o.descriptor_finalize(_f(B, D), None, None, 1e-10).shape == (B, D)
        n2, dot, means = o.region_crop_stats(fm, idx, nsel, _f(B, Kin), fh, fh)
        assert n2.shape == (B, 8) and means.shape == (B, 8, C)
        assert o.region_scatter_grad(fm, idx, nsel, _f(B, Kin), n2, dot, None, fh, fh, 1e-10).shape == fm.shape
        lw = _f(D, Kin, dtype=bf)
        d, c, i, n = o.region_descriptors(fm, cw, cwb, cwb, _f(ncls), _f(Kin), lw, lw, _f(D), fh, fh, k, 1.0)
        assert d.shape == (B, D) and c.shape == (B, ncls, k) and i.shape == (B, k) and n.dtype == i32
        assert o.global_descriptors(_f(B, C, 7, 7), _f(Kin), lw, None, None).shape == (B, D)
        # mining, metrics, DBA, loss
        neg, ns, ps = o.select_negatives(_f(50, 64), _f(50, dtype=i32), _f(9, dtype=i64), _f(9, dtype=i64), True, 1)
        assert neg.shape == (9,) and neg.dtype == i64 and ps.shape == (9,)
        v, j = o.row_kth_largest(_f(6, 40), 2)
        assert v.shape == (6,) and j.dtype == i64
        assert o.row_ranks(_f(6, 40), _f(6, 3, dtype=i32)).shape == (6, 3)
        assert o.instance_avg(_f(50, 64), _f(50, dtype=i32), -1).shape == (50, 64)
        loss, clamp = o.triplet_loss_forward(_f(8, 16), _f(8, 16), _f(8, 16), 0.2, True, True)
        assert loss.shape == (1,) and clamp.dtype == torch.uint8
        g = o.triplet_loss_backward(_f(8, 16), _f(8, 16), _f(8, 16), clamp, loss, True, True)
        assert all(t.shape == (8, 16) for t in g)


def test_ops_have_no_cpu_kernel():
    x = torch.zeros(3, 8)
    with pytest.raises(NotImplementedError):
        torch.ops.isb.l2norm_rows(x, 1e-10)
    with pytest.raises(NotImplementedError):
        torch.ops.isb.topk_search(x, x, x.bfloat16(), 2, -1, 0)
    with pytest.raises(NotImplementedError):
        torch.ops.isb.select_negatives(x, torch.zeros(3, dtype=torch.int32), torch.zeros(1, dtype=torch.int64),
                                       torch.zeros(1, dtype=torch.int64), False, 1)
