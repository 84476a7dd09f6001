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

ABI_VERSION = 6   # == ISB_ABI_VERSION in include/isb.h

# name -> (restype, argtypes); mirrors include/isb.h one to one
SIGNATURES = {
    "isb_abi_version": (c_int, []),
    "isb_last_error": (ctypes.c_char_p, []),
    "isb_check_device": (c_int, []),
    "isb_set_option": (c_int, [c_int, c_int]),
    "isb_get_option": (c_int, [c_int]),
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
    "isb_topk_global_threshold": (c_int, [c_ptr, c_int, c_i64, c_int, c_int, c_ptr, c_ptr]),
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
    "isb_region_crop_stats": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                      c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "isb_region_scatter_grad": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                        c_i64, c_ptr, c_ptr, c_ptr, c_f32, c_ptr, c_ptr]),
    "isb_descriptor_finalize": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_f32, c_ptr, c_ptr]),
    "isb_select_negatives_workspace_bytes": (c_size, [c_i64, c_i64, c_i64, c_int]),
    "isb_select_negatives": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64,
                                     c_int, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size,
                                     c_ptr]),
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


# ISB_OPT_* of include/isb.h
OPTIONS = {
    "screen_pair": 0, "screen_seed": 1, "screen_wavesync": 2, "mining_kc": 3, "pool_stages": 4, "pool_g": 5,
    "pool_generic_geom": 6, "region_pool_tc": 7, "tc_debug": 8, "gather_cw": 9, "gather_g": 10,
    "gather_stages": 11, "gemm_pair": 12, "mining_progressive": 13, "gather_small": 14, "screen_groups": 15, "resc_splits": 16, "region_top_select": 17,
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


def set_option(name, value):
    """Set a tuning option of the library (include/isb.h, ISB_OPT_*); value None or -1
    restores the default.  The library reads nothing from the environment."""
    if name not in OPTIONS:
        raise IsbError("unknown option %r (known: %s)" % (name, ", ".join(sorted(OPTIONS))))
    check(lib().isb_set_option(OPTIONS[name], -1 if value is None else int(value)), "isb_set_option")


def get_option(name):
    """The explicitly set value of an option, or None when its default is in force."""
    v = lib().isb_get_option(OPTIONS[name])
    return None if v < 0 else v


class options(object):
    """``with options(screen_pair=0): ...`` -- set options for a block, restore afterwards."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False
