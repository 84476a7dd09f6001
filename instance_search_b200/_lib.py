"""ctypes binding of libisb.so (the C ABI declared in include/isb.h).

There is no fallback: if the library is missing or a call fails, the product
path raises.  Nothing here imports ``oracle``.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libisb.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p
c_size = ctypes.c_size_t

ABI_VERSION = 4   # == ISB_ABI_VERSION in include/isb.h

# name -> (restype, argtypes); mirrors include/isb.h one to one
SIGNATURES = {
    "isb_abi_version": (c_int, []),
    "isb_last_error": (ctypes.c_char_p, []),
    "isb_check_device": (c_int, []),
    "isb_l2norm_rows": (c_int, [c_ptr, c_i64, c_i64, c_f32, c_ptr, c_ptr]),
    "isb_shift_rows": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "isb_f32_to_bf16": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_int, c_ptr]),
    "isb_topk_search_workspace_bytes": (c_size, [c_i64, c_i64, c_i64, c_int, c_int]),
    "isb_topk_search": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_int,
                                c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "isb_topk_resolve_workspace_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "isb_topk_resolve": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_int, c_i64,
                                 c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "isb_topk_exhaustive_workspace_bytes": (c_size, [c_i64, c_i64, c_int]),
    "isb_topk_exhaustive": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, c_i64, c_ptr, c_i64, c_ptr,
                                    c_ptr, c_ptr, c_size, c_ptr]),
    "isb_topk_screen": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_int, c_int, c_ptr, c_size,
                                c_ptr]),
    "isb_topk_rerank": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_int, c_int, c_i64, c_ptr, c_ptr,
                                c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "isb_topk_merge": (c_int, [c_ptr, c_ptr, c_int, c_i64, c_int, c_ptr, c_ptr, c_ptr]),
    "isb_topk_candidates": (c_int, [c_i64, c_i64, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_size,
                                    c_ptr]),
    "isb_topk_global_threshold": (c_int, [c_ptr, c_int, c_i64, c_int, c_ptr, c_ptr]),
    "isb_topk_rerank_owned": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                      c_ptr, c_ptr]),
    "isb_topk_merge_certified": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                         c_ptr]),
    "isb_region_select_workspace_bytes": (c_size, [c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                                  c_int, c_int]),
    "isb_region_select": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr,
                                  c_i64, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "isb_region_logits": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr,
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_ptr,
                                  c_ptr, c_ptr, c_ptr]),
    "isb_region_gather": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_ptr,
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "isb_descriptor_finalize": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_f32, c_ptr, c_ptr]),
    "isb_select_negatives_workspace_bytes": (c_size, [c_i64, c_i64, c_i64, c_int]),
    "isb_select_negatives": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64,
                                     c_int, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "isb_row_kth_largest": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr]),
    "isb_row_ranks": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_int, c_ptr, c_ptr]),
    "isb_instance_avg": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr]),
    "isb_l2norm_rows_backward": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_f32, c_ptr, c_ptr]),
    "isb_col_sums": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "isb_triplet_loss_forward": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_f32, c_int, c_int, c_ptr,
                                         c_ptr, c_ptr, c_ptr]),
    "isb_triplet_loss_backward": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_int, c_int,
                                          c_ptr, c_ptr, c_ptr, c_ptr]),
    "isb_gemm_nt_workspace_bytes": (c_size, [c_i64, c_i64, c_i64, c_int]),
    "isb_gemm_nt": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64,
                            c_int, c_ptr, c_size, c_ptr]),
    "isb_gemm_nt_split": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr,
                                  c_ptr, c_i64, c_int, c_ptr, c_size, c_ptr]),
}


class IsbError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libisb.so once; raise (never fall back) when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IsbError(
            "libisb.so not found at %s -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C instance_search_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    l = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if l.isb_abi_version() != ABI_VERSION:
        raise IsbError("libisb.so ABI version mismatch")
    _lib = l
    return l


def check(rc, what):
    if rc != 0:
        msg = lib().isb_last_error().decode("utf-8", "replace")
        raise IsbError("%s failed (code %d): %s" % (what, rc, msg))
