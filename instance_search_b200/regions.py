"""Region-descriptor head on the GPU (functional form).

Everything RegionDescriptorNet.forward_single does after the trunk
(model/siamese.py:187-222), for a whole batch of feature maps at once; the
result equals looping the reference's batch-1 forward over the images
(train/siamese_regions.py:31-38).  All arithmetic is in libisb.so.
"""

import torch

from . import _lib, ops
from ._lib import IsbError

WINDOW_MARGIN = 10   # windows re-scored exactly beyond k (k + margin <= 32)
LARGE_MAP_WINDOWS = 128   # maps with more windows than this always use the full 32 candidates


class HeadWeights(object):
    """Device-resident, tensor-core-ready copies of the head's parameters.

    cls_w [ncls, C], cls_b [ncls]    classifier.0.{weight,bias} (1x1 conv view)
    shift [C*fh*fw]                  feature_reduc1.1.param
    lin_w [D, C*fh*fw], lin_b [D]    feature_reduc1.2.{weight,bias}
    terms: 1 = plain bf16 projection; 3 = split operands (hi + lo bf16 terms, three
    tcgen05 products per tile), an fp32-grade product (default: descriptors feed an
    index-exact search).
    cls_w / lin_w may be None (DescriptorNet has no classifier).
    """

    def __init__(self, cls_w, cls_b, shift, lin_w, lin_b, terms=3):
        if terms not in (1, 3):
            raise IsbError("terms must be 1 or 3")
        self.terms = terms
        self.cls_w = None if cls_w is None else ops._f32c(cls_w.reshape(cls_w.size(0), -1))
        self.cls_b = None if cls_b is None else ops._f32c(cls_b)
        self.cls_w_hi = None if cls_w is None else ops.to_bf16(self.cls_w, 0)
        self.cls_w_lo = None if cls_w is None else ops.to_bf16(self.cls_w, 1)
        self.cls_w_absmax = 0.0 if cls_w is None else float(self.cls_w.abs().max())
        self.shift = ops._f32c(shift)
        self.lin_b = None if lin_b is None else ops._f32c(lin_b)
        lin_w = ops._f32c(lin_w)
        self.D, self.Kin = lin_w.shape
        self.KinP = (self.Kin + 7) // 8 * 8   # every bf16 term is zero-padded to 8 columns
        self.lin_w_hi = ops.to_bf16(lin_w, 0)
        self.lin_w_lo = ops.to_bf16(lin_w, 1) if terms == 3 else None


def _splits_for(M, N, K):
    """Split K so the projection fills the 148 SMs (tiles are 128 x 256)."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    kb = (K + 63) // 64
    return max(1, min(kb, 148 // max(1, tiles)))


def _project(U_hi, U_lo, hw, B):
    """y = u . W^T  (nn.Linear(100352, D) without the bias, model/siamese.py:180)"""
    splits = _splits_for(B, hw.D, hw.KinP * hw.terms)
    if hw.terms == 1:
        return ops.gemm_nt(U_hi, hw.lin_w_hi, splits=splits)
    return ops.gemm_nt_split(U_hi, U_lo, hw.lin_w_hi, hw.lin_w_lo, splits=splits)


_PROBE_CACHE = {}


def region_pool_probe(x, hw, k, fsize, margin=WINDOW_MARGIN):
    """Measurement probe: the pooling pass of isb_region_select alone (exact_mode = -1: x read
    once, window means written as bf16 hi + lo, per-pixel energy).  Buffers are cached between
    calls so that a timed call is one memset and one kernel.  Returns (bytes read, bytes written)."""
    ops._need_cuda(x)
    B, C, H, W = x.shape
    fh, fw = fsize
    ncls = hw.cls_w.size(0)
    margin = max(0, min(margin, 32 - k))
    L = _lib.lib()
    key = (B, C, H, W, ncls, k, margin, x.device)
    if key not in _PROBE_CACHE:
        dev = x.device
        nbytes = L.isb_region_select_workspace_bytes(B, C, H, W, ncls, fh, fw, k, margin)
        _PROBE_CACHE.clear()
        _PROBE_CACHE[key] = (
            torch.empty((B, k), dtype=torch.int64, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
            torch.empty((B, ncls, k), dtype=torch.float32, device=dev),
            torch.empty((B, k), dtype=torch.float32, device=dev), torch.empty((B, k), dtype=torch.float32, device=dev),
            torch.empty((B,), dtype=torch.float32, device=dev), torch.zeros(1 + B, dtype=torch.int32, device=dev),
            torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev), nbytes)
    idx, nsel, cls_out, win_norm, approx_max, runner_up, n_unc, ws, nbytes = _PROBE_CACHE[key]
    _lib.check(L.isb_region_select(x.data_ptr(), B, C, H, W, hw.cls_w.data_ptr(), hw.cls_w_hi.data_ptr(),
                                   hw.cls_w_lo.data_ptr(), hw.cls_w_hi.size(1), hw.cls_b.data_ptr(), ncls,
                                   fh, fw, k, margin, -1, idx.data_ptr(), nsel.data_ptr(),
                                   cls_out.data_ptr(), win_norm.data_ptr(), approx_max.data_ptr(),
                                   runner_up.data_ptr(), n_unc.data_ptr(), ws.data_ptr(), nbytes,
                                   ops._stream()), "isb_region_select")
    nwin = (H - fh + 1) * (W - fw + 1)
    return 4 * B * C * H * W, 2 * 2 * B * nwin * ((C + 7) // 8 * 8)


def region_select(x, hw, k, fsize, margin=WINDOW_MARGIN, exact_mode=False):
    """a3, first stage: (idx [B,k] int64, nsel [B] int32, cls_out [B,ncls,k], win_norm [B,k],
    approx_max [B,k], runner_up [B], n_uncertified [1] int32).  exact_mode: the slow
    fp64 second line (candidates = 32), cls_out / order final."""
    ops._need_cuda(x)
    x = ops._f32c(x)
    B, C, H, W = x.shape
    fh, fw = fsize
    ncls = hw.cls_w.size(0)
    nwin = (H - fh + 1) * (W - fw + 1)
    if nwin > LARGE_MAP_WINDOWS:
        margin = 32 - k   # dense maps: the screen's candidate list must reach further down
    margin = 32 - k if exact_mode is True else max(0, min(margin, 32 - k))
    dev = x.device
    idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    nsel = torch.empty((B,), dtype=torch.int32, device=dev)
    cls_out = torch.empty((B, ncls, k), dtype=torch.float32, device=dev)
    win_norm = torch.empty((B, k), dtype=torch.float32, device=dev)
    approx_max = torch.empty((B, k), dtype=torch.float32, device=dev)
    runner_up = torch.empty((B,), dtype=torch.float32, device=dev)
    # count, then the uncertified images; the library zeroes the count, entries beyond it are never read
    n_unc = torch.empty(1 + B, dtype=torch.int32, device=dev)
    L = _lib.lib()
    nbytes = L.isb_region_select_workspace_bytes(B, C, H, W, ncls, fh, fw, k, margin)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    _lib.check(L.isb_region_select(x.data_ptr(), B, C, H, W, hw.cls_w.data_ptr(), hw.cls_w_hi.data_ptr(),
                                   hw.cls_w_lo.data_ptr(), hw.cls_w_hi.size(1), hw.cls_b.data_ptr(), ncls,
                                   fh, fw, k, margin, int(exact_mode), idx.data_ptr(), nsel.data_ptr(),
                                   cls_out.data_ptr(), win_norm.data_ptr(), approx_max.data_ptr(),
                                   runner_up.data_ptr(), n_unc.data_ptr(), ws.data_ptr(), nbytes,
                                   ops._stream()), "isb_region_select")
    return idx, nsel, cls_out, win_norm, approx_max, runner_up, n_unc


RUNNER_UPS = 2   # windows beyond k whose logits are also recomputed in true fp32


def region_gather(x, hw, k, fsize, idx, nsel, win_norm, k_sum=None, want_means=True, out=None,
                  image_list=None, n_list=None):
    """a4 operand: (U_hi, U_lo or None, win_mean [B,k,C] or None).  idx / win_norm are
    [B, k]; the first min(nsel, k_sum) windows of every image are summed.  image_list /
    n_list (device int32): fix-up pass over the listed images only, into ``out``."""
    x = ops._f32c(x)
    B, C, H, W = x.shape
    fh, fw = fsize
    if C * fh * fw != hw.Kin:
        raise IsbError("feature map channels x window (%d) != projection in_features (%d)" %
                       (C * fh * fw, hw.Kin))
    if out is None:
        U_hi = torch.empty((B, hw.KinP), dtype=torch.bfloat16, device=x.device)
        U_lo = torch.empty_like(U_hi) if hw.terms == 3 else None
    else:
        U_hi, U_lo = out
    win_mean = torch.empty((B, k, C), dtype=torch.float32, device=x.device) if want_means else None
    _lib.check(_lib.lib().isb_region_gather(x.data_ptr(), B, C, H, W, fh, fw, k, k if k_sum is None else k_sum,
                                            ops._ptr(image_list), ops._ptr(n_list), idx.data_ptr(),
                                            nsel.data_ptr(), win_norm.data_ptr(), hw.shift.data_ptr(),
                                            U_hi.data_ptr(), ops._ptr(U_lo), hw.KinP, ops._ptr(win_mean),
                                            ops._stream()), "isb_region_gather")
    return U_hi, U_lo, win_mean


def region_logits(win_mean, hw, k, nsel_in, idx_in, norm_in, approx_max, runner_up, approx_cls=None):
    """Exact fp32 logits of the ke = win_mean.size(1) scored windows -> the final k.
    Returns (idx [B,k], win_norm [B,k], nsel [B], cls_out [B,ncls,k], changed_list [B] int32,
    n_changed [1], n_uncertified [1]).  approx_cls (region_select's cls_out): class-max
    only -- cls_out is not computed and returned as None."""
    B, ke, C = win_mean.shape
    ncls = hw.cls_w.size(0)
    dev = win_mean.device
    idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    norm = torch.empty((B, k), dtype=torch.float32, device=dev)
    nsel = torch.empty((B,), dtype=torch.int32, device=dev)
    cls_out = torch.empty((B, ncls, k), dtype=torch.float32, device=dev) if approx_cls is None else None
    changed = torch.empty((B,), dtype=torch.int32, device=dev)
    n_changed = torch.empty(1, dtype=torch.int32, device=dev)        # zeroed by the library
    n_unc = torch.empty(1 + B, dtype=torch.int32, device=dev)        # count (zeroed by the library), then the images
    _lib.check(_lib.lib().isb_region_logits(win_mean.data_ptr(), hw.cls_w.data_ptr(), hw.cls_b.data_ptr(), B, C,
                                            ncls, ke, k, nsel_in.data_ptr(), approx_max.data_ptr(),
                                            runner_up.data_ptr(), idx_in.data_ptr(), norm_in.data_ptr(),
                                            idx.data_ptr(), norm.data_ptr(), nsel.data_ptr(), ops._ptr(cls_out),
                                            ops._ptr(approx_cls), hw.cls_w_absmax,
                                            changed.data_ptr(), n_changed.data_ptr(), n_unc.data_ptr(),
                                            ops._stream()), "isb_region_logits")
    return idx, norm, nsel, cls_out, changed, n_changed, n_unc


def descriptor_finalize(y, bias, nsel, eps=1e-10):
    """a4 (bias) + a5: desc = l2norm(y + nsel * bias); nsel None = 1 (DescriptorNet).
    reference: model/siamese.py:220-222."""
    ops._need_cuda(y, bias, nsel)
    y = ops._f32c(y)
    B, D = y.shape
    desc = torch.empty((B, D), dtype=torch.float32, device=y.device)
    if B == 0:
        return desc
    if nsel is not None:
        nsel = nsel.to(torch.int32).contiguous()
    _lib.check(_lib.lib().isb_descriptor_finalize(y.data_ptr(), B, D, ops._ptr(None if bias is None else ops._f32c(bias)),
                                                  ops._ptr(nsel), float(eps), desc.data_ptr(), ops._stream()),
               "isb_descriptor_finalize")
    return desc


def region_project(U_hi, U_lo, hw, nsel, extras=None):
    """a4 projection + a5: desc = l2norm(u . W^T + nsel * bias).  extras (dict): receives
    'y' = u . W^T, what the backward of the final normalisation needs."""
    y = _project(U_hi, U_lo, hw, U_hi.size(0))
    if extras is not None:
        extras["y"] = y
    return descriptor_finalize(y, hw.lin_b, nsel)


def region_head(x, hw, k, fsize, margin=WINDOW_MARGIN, want_cls_out=True):
    """The certified fast path, all asynchronous: returns (U_hi, U_lo, idx, nsel, cls_out,
    n_uncertified [2, 1 + B] device int32: per certificate, the count and then the images it
    rejected).  want_cls_out=False (eval): cls_out is None and only the contending classes of
    every window are scored in fp32."""
    ke = min(32, k + RUNNER_UPS)
    # 1. screen all windows, fp32-grade scores of the candidates, the best ke of them
    idx_e, nsel_e, approx_cls, norm_e, approx_e, runner_up, n1 = region_select(x, hw, ke, fsize, margin)
    # 2. operand from the first k, exact means of all ke
    U_hi, U_lo, win_mean = region_gather(x, hw, ke, fsize, idx_e, nsel_e, norm_e, k_sum=k)
    # 3. true-fp32 logits -> final k, exact order; images whose k changed are listed
    idx, norm, nsel, cls_out, changed, n_changed, n2 = region_logits(
        win_mean, hw, k, nsel_e, idx_e, norm_e, approx_e, runner_up, None if want_cls_out else approx_cls)
    # 4. fix-up: re-gather the (rare) images where a runner-up entered the top k
    region_gather(x, hw, k, fsize, idx, nsel, norm, want_means=False, out=(U_hi, U_lo), image_list=changed,
                  n_list=n_changed)
    return U_hi, U_lo, idx, nsel, cls_out, torch.stack([n1, n2])


def region_descriptors_async(x, hw, k, fsize, margin=WINDOW_MARGIN, want_cls_out=True, extras=None):
    """The certified fast path with nothing read back: (desc, cls_out, idx, nsel,
    n_uncertified [2, 1 + B] device int32).  The caller checks n_uncertified[:, 0] whenever
    it likes (e.g. after queueing the next batch) and, if it is non-zero, patches the
    listed images with region_descriptors_fix.  extras (dict): receives the projection's
    operand and output ('U_hi', 'U_lo', 'y') for the backward pass."""
    U_hi, U_lo, idx, nsel, cls_out, n_unc = region_head(x, hw, k, fsize, margin, want_cls_out)
    if extras is not None:
        extras["U_hi"], extras["U_lo"] = U_hi, U_lo
    desc = region_project(U_hi, U_lo, hw, nsel, extras)
    return desc, cls_out, idx, nsel, n_unc


class GraphedRegionDescriptors(object):
    """The certified fast path (region_descriptors_async) captured ONCE into a CUDA graph for a fixed
    input buffer: every replay() re-runs the whole chain (pool -> classifier screen -> reselect ->
    gather -> logits -> fix-up -> projection -> finalize: ~10 launches and 3 memsets) with one
    graph launch on whatever ``x`` holds by then.  The chain is launch-bound at small maps (35 us of
    its 350 at 14 x 14 are gaps between kernels); a replay is 3-4 % faster than the eager calls and
    costs the host one call instead of six.  Outputs are the SAME tensors on every replay: consume
    (or copy) them before the next one.  The caller checks ``n_uncertified[:, 0]`` as after
    region_descriptors_async and patches the listed images with region_descriptors_fix.

        graphed = GraphedRegionDescriptors(x_buffer, hw, k, (7, 7))
        for batch in loader:
            x_buffer.copy_(trunk(batch))          # or have the trunk write into x_buffer
            desc, cls_out, idx, nsel, n_unc = graphed.replay()
    """

    def __init__(self, x, hw, k, fsize, margin=WINDOW_MARGIN, want_cls_out=False, warmup=2):
        ops._need_cuda(x)
        if not x.is_contiguous() or x.dtype != torch.float32:
            raise IsbError("GraphedRegionDescriptors: x must be a contiguous fp32 CUDA tensor (the graph reads it in place)")
        self.x = x
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):             # warm-up outside the capture: allocator, function attributes
            for _ in range(max(1, warmup)):
                region_descriptors_async(x, hw, k, fsize, margin, want_cls_out)
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = region_descriptors_async(x, hw, k, fsize, margin, want_cls_out)

    def replay(self):
        self.graph.replay()
        return self.outputs


def uncertified_images(n_unc):
    """Sorted image indices listed by the certificates (host sync).  n_unc: [2, 1 + B]."""
    host = n_unc.cpu()
    bad = set()
    for row in host:
        bad.update(int(v) for v in row[1:1 + int(row[0])])
    return sorted(bad)


def region_descriptors_exact(x, hw, k, fsize, extras=None):
    """The fp64-exact second line (rounds of up to 32 candidate windows re-scored from the fp32
    inputs until the completeness certificate holds or the whole map has been scored): what
    uncertified images are redone with.  Complete by construction; the one case it cannot
    finish (k fills all 32 candidate slots of a larger map) raises."""
    idx, nsel, cls_out, win_norm, _, _, n_unc = region_select(x, hw, k, fsize, exact_mode=True)
    if k >= 32 and int(n_unc[0]):
        raise IsbError("region_descriptors_exact: %d image(s) could not be certified with k = %d "
                       "(k must leave free candidate slots: k < 32)" % (int(n_unc[0]), k))
    U_hi, U_lo, _ = region_gather(x, hw, k, fsize, idx, nsel, win_norm, want_means=False)
    if extras is not None:
        extras["U_hi"], extras["U_lo"] = U_hi, U_lo
    return region_project(U_hi, U_lo, hw, nsel, extras), cls_out, idx, nsel


def region_descriptors_fix(x, hw, k, fsize, bad, desc, cls_out, idx, nsel, extras=None):
    """Redo the images ``bad`` (list of batch indices) with the exact second line and patch
    the fast path's outputs (and the backward's operands in ``extras``) in place."""
    sel = torch.tensor(bad, dtype=torch.int64, device=x.device)
    sub = {} if extras is not None else None
    d2, c2, i2, n2 = region_descriptors_exact(x.index_select(0, sel).contiguous(), hw, k, fsize, sub)
    desc.index_copy_(0, sel, d2)
    idx.index_copy_(0, sel, i2)
    nsel.index_copy_(0, sel, n2)
    if cls_out is not None:
        cls_out.index_copy_(0, sel, c2)
    if extras is not None:
        for name in ("U_hi", "U_lo", "y"):
            if extras.get(name) is not None:
                extras[name].index_copy_(0, sel, sub[name])


def region_descriptors(x, hw, k, fsize, margin=WINDOW_MARGIN, exact=True, stats=None, want_cls_out=True,
                       extras=None):
    """x [B, C, H, W] trunk feature maps -> (desc [B, D], cls_out [B, ncls, k],
    idx [B, k], nsel [B]).  reference: model/siamese.py:187-223 per image.

    exact=True: the certificates of the fast path (screen completeness, selection
    against the unscored runner-ups) are read back (ONE small D2H read, issued after
    the projection has been queued so the GPU does not idle during the round trip);
    the images they reject -- and only those -- are redone with the fp64-exact second
    line.  exact=False skips the read-back (no host sync)."""
    desc, cls_out, idx, nsel, n_unc = region_descriptors_async(x, hw, k, fsize, margin, want_cls_out, extras)
    if exact:
        n_bad = int(n_unc[:, 0].sum().item())
        if stats is not None:
            stats["batches"] = stats.get("batches", 0) + 1
            stats["batches_resolved_exactly"] = stats.get("batches_resolved_exactly", 0) + (1 if n_bad else 0)
            stats["images_resolved_exactly"] = stats.get("images_resolved_exactly", 0)
        if n_bad:
            bad = uncertified_images(n_unc)
            if stats is not None:
                stats["images_resolved_exactly"] += len(bad)
            region_descriptors_fix(x, hw, k, fsize, bad, desc, cls_out, idx, nsel, extras)
    return desc, cls_out, idx, nsel


# ------------------------------------------------------------------ backward glue (f3)
def crop_stats(x, k, fsize, idx, nsel, g_u=None, want_means=True):
    """Per selected window: (|crop|^2 [B, k], <crop, g_u[b]> [B, k], window means [B, k, C] or
    None) -- isb_region_crop_stats.  g_u [B, ld] fp32 (gradient w.r.t. the projection operand)."""
    ops._need_cuda(x, idx, nsel, g_u)
    x = ops._f32c(x)
    B, C, H, W = x.shape
    n2 = torch.empty((B, k), dtype=torch.float32, device=x.device)
    dot = torch.empty((B, k), dtype=torch.float32, device=x.device)
    means = torch.empty((B, k, C), dtype=torch.float32, device=x.device) if want_means else None
    if B == 0:
        return n2, dot, means
    if g_u is not None:
        g_u = ops._f32c(g_u)
    _lib.check(_lib.lib().isb_region_crop_stats(
        x.data_ptr(), B, C, H, W, fsize[0], fsize[1], k, idx.contiguous().data_ptr(),
        nsel.to(torch.int32).contiguous().data_ptr(), ops._ptr(g_u), 0 if g_u is None else g_u.size(1),
        n2.data_ptr(), dot.data_ptr(), ops._ptr(means), ops._stream()), "isb_region_crop_stats")
    return n2, dot, means


def scatter_grad(x, k, fsize, idx, nsel, g_u, n2, dot, g_mean, eps=1e-10):
    """g_x [B, C, H, W]: gradient of the head w.r.t. the feature map from the gradient of the
    projection operand (g_u, through the crops' L2 normalisation) and of the window means
    (g_mean [B, k, C], through the sliding mean) -- isb_region_scatter_grad."""
    ops._need_cuda(x, idx, nsel, g_u, g_mean)
    x = ops._f32c(x)
    B, C, H, W = x.shape
    g_x = torch.empty_like(x)
    if B == 0:
        return g_x
    if g_u is not None:
        g_u = ops._f32c(g_u)
    if g_mean is not None:
        g_mean = ops._f32c(g_mean)
    _lib.check(_lib.lib().isb_region_scatter_grad(
        x.data_ptr(), B, C, H, W, fsize[0], fsize[1], k, idx.contiguous().data_ptr(),
        nsel.to(torch.int32).contiguous().data_ptr(), ops._ptr(g_u), 0 if g_u is None else g_u.size(1),
        ops._f32c(n2).data_ptr(), ops._f32c(dot).data_ptr(), ops._ptr(g_mean), float(eps), g_x.data_ptr(),
        ops._stream()), "isb_region_scatter_grad")
    return g_x


def global_descriptors(x, hw):
    """DescriptorNet head: flatten -> L2 -> Shift -> Linear -> L2.
    reference: model/siamese.py:117-122.  x [B, C, fh, fw]."""
    ops._need_cuda(x)
    x = ops._f32c(x)
    B = x.size(0)
    flat = x.reshape(B, -1)
    if flat.size(1) != hw.Kin:
        raise IsbError("flattened features (%d) != projection in_features (%d)" % (flat.size(1), hw.Kin))
    u = ops.shift_rows(ops.l2norm_rows(flat), hw.shift)
    y = _project(ops.to_bf16(u, 0), ops.to_bf16(u, 1) if hw.terms == 3 else None, hw, B)
    return descriptor_finalize(y, hw.lin_b, None)


def classif_embeddings(x, hw, fsize):
    """Embeddings of the classification track's sub-window net (TuneClassifSub): for every image the
    class scores at the window with the highest maximal activation, L2-normalised -- [B, ncls].
    reference: train/classif_regions.py:107-132 (one image per forward there).  The window is the
    head's exact top-1 selection (same certified path as region_descriptors with k = 1), the scores
    are its true-fp32 logits."""
    ops._need_cuda(x)
    k = 1
    ke = min(32, k + RUNNER_UPS)
    idx_e, nsel_e, _, norm_e, approx_e, runner_up, n1 = region_select(x, hw, ke, fsize)
    # exact means of the k + 2 best windows (the gather's by-product) -> true fp32 logits -> the top-1
    gw = object.__new__(HeadWeights)
    gw.terms, gw.shift = 1, torch.zeros(x.size(1) * fsize[0] * fsize[1], dtype=torch.float32, device=x.device)
    gw.Kin = gw.shift.numel()
    gw.KinP = (gw.Kin + 7) // 8 * 8
    _, _, win_mean = region_gather(x, gw, ke, fsize, idx_e, nsel_e, norm_e, k_sum=k)
    idx, _, nsel, cls_out, _, _, n2 = region_logits(win_mean, hw, k, nsel_e, idx_e, norm_e, approx_e, runner_up)
    bad = uncertified_images(torch.stack([n1, n2]))
    if bad:   # near-tied windows: the fp64-exact second line for those images
        sel = torch.tensor(bad, dtype=torch.int64, device=x.device)
        i2, _, c2, _, _, _, _ = region_select(x.index_select(0, sel).contiguous(), hw, k, fsize, exact_mode=True)
        cls_out.index_copy_(0, sel, c2)
        idx.index_copy_(0, sel, i2)
    return ops.l2norm_rows(cls_out[:, :, 0].contiguous()), idx[:, 0]
