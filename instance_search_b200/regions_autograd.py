"""Autograd for the fused region-descriptor head: the training path of RegionDescriptorNet
(reference: model/siamese.py:225-229 -> forward_single :185-223, three times per triplet) with the
backward passes of the reference's custom Functions (model/custom_modules.py:20-25 Shift,
:59-67 NormalizeL2) and of torch's conv / linear / avg-pool, for the WHOLE batch at once.

Forward = regions.region_descriptors (the certified fused head).  The window selection
(class-max, top-k) is piecewise constant, so the backward differentiates, per image b with
selected windows i < n_b:

    u_b = sum_i crop_i / sqrt(|crop_i|^2 + eps) + n_b * shift          (NormalizeL2, Shift, sum)
    z_b = W u_b + n_b * bias ;  desc_b = z_b / sqrt(|z_b|^2 + eps)      (Linear, NormalizeL2)
    cls_out[b, :, i] = Wc mean_i + bc                                  (AvgPool, 1x1 conv)

Dense parts are split-operand tcgen05 GEMMs (fp32-grade): g_u = g_z W, dW = g_z^T u,
dWc = g_cls^T mean, g_mean = g_cls Wc.  The rest is two bandwidth kernels
(isb_region_crop_stats, isb_region_scatter_grad).  No per-image Python loop, no cuBLAS / cuDNN.
"""

import torch

from . import ops, regions


def _split(t):
    """fp32 [M, K] -> (hi, lo) bf16 operands [M, K rounded to 8]"""
    return ops.to_bf16(t, 0), ops.to_bf16(t, 1)


def _mm_nt(a, b):
    """fp32-grade a [M, K] . b [N, K]^T on the tensor cores"""
    a_hi, a_lo = _split(a.contiguous())
    b_hi, b_lo = _split(b.contiguous())
    M, N, K = a.size(0), b.size(0), a_hi.size(1)
    return ops.gemm_nt_split(a_hi, a_lo, b_hi, b_lo, splits=regions._splits_for(M, N, 3 * K))


class RegionHeadFunction(torch.autograd.Function):
    """(desc [B, D], cls_out [B, ncls, k]) = head(x [B, C, H, W]; cls_w, cls_b, shift, lin_w, lin_b)."""

    @staticmethod
    def forward(ctx, x, cls_w, cls_b, shift, lin_w, lin_b, hw, k, fsize):
        extras = {}
        desc, cls_out, idx, nsel = regions.region_descriptors(x.detach(), hw, k, fsize, extras=extras)
        nf = nsel.to(torch.float32).unsqueeze(1)
        z = extras["y"] if lin_b is None else extras["y"] + nf * lin_b.detach()
        ctx.k, ctx.fsize = k, tuple(fsize)
        ctx.has_lin_b = lin_b is not None
        ctx.save_for_backward(x, cls_w, lin_w, z, extras["U_hi"], extras["U_lo"], idx, nsel)
        ctx.mark_non_differentiable(idx, nsel)
        return desc, cls_out, idx, nsel

    @staticmethod
    def backward(ctx, g_desc, g_cls, _gi, _gn):
        x, cls_w, lin_w, z, U_hi, U_lo, idx, nsel = ctx.saved_tensors
        k, fsize = ctx.k, ctx.fsize
        need_x, need_cw, need_cb, need_s, need_w, need_b = ctx.needs_input_grad[:6]
        B, C = x.size(0), x.size(1)
        Kin = C * fsize[0] * fsize[1]
        nf = nsel.to(torch.float32).unsqueeze(1)
        g_x = g_cw = g_cb = g_s = g_w = g_b = None

        # ---- descriptor branch: desc <- z <- u <- crops
        g_u = None
        if g_desc is not None and (need_x or need_s or need_w or need_b):
            g_z = ops.l2norm_rows_backward(z, g_desc.contiguous())            # custom_modules.py:59-67
            if need_b and ctx.has_lin_b:
                g_b = ops.col_sums(g_z * nf)                                   # bias enters n_b times
            if need_w:
                u = U_hi[:, :Kin].float() if U_lo is None else U_hi[:, :Kin].float() + U_lo[:, :Kin].float()
                g_w = _mm_nt(g_z.t(), u.t())                                   # dW = g_z^T u   [D, Kin]
            if need_x or need_s:
                g_u = _mm_nt(g_z, lin_w.detach().t())                          # g_u = g_z W    [B, Kin]
                if need_s:
                    g_s = ops.col_sums(g_u * nf)                               # custom_modules.py:23-24, n_b times

        # ---- classifier branch: cls_out <- window means <- x
        g_mean = None
        live = (torch.arange(k, device=x.device).unsqueeze(0) < nsel.unsqueeze(1))   # [B, k]
        want_cls = g_cls is not None and (need_x or need_cw or need_cb)
        means = None
        n2 = dot = None
        if need_x or (want_cls and need_cw):
            n2, dot, means = regions.crop_stats(x, k, fsize, idx, nsel, g_u if need_x else None,
                                                want_means=want_cls and need_cw)
        if want_cls:
            G = (g_cls.permute(0, 2, 1) * live.unsqueeze(2)).reshape(B * k, -1)      # [B*k, ncls], dead slots 0
            if need_cb:
                g_cb = ops.col_sums(G.contiguous())
            if need_cw:
                g_cw = _mm_nt(G.t(), means.reshape(B * k, C).t()).reshape(cls_w.shape)   # [ncls, C]
            if need_x:
                g_mean = _mm_nt(G, cls_w.detach().reshape(cls_w.size(0), -1).t()).reshape(B, k, C)
        if need_x:
            g_x = regions.scatter_grad(x, k, fsize, idx, nsel, g_u, n2, dot, g_mean)
        return g_x, g_cw, g_cb, g_s, g_w, g_b, None, None, None


def region_head(x, conv, shift, lin, hw, k, fsize):
    """(desc, cls_out) of the fused head with autograd; conv / shift / lin are the modules
    classifier[0], feature_reduc1[1], feature_reduc1[2] of RegionDescriptorNet."""
    desc, cls_out, _, _ = RegionHeadFunction.apply(x, conv.weight, conv.bias, shift.param, lin.weight, lin.bias,
                                                   hw, k, tuple(fsize))
    return desc, cls_out
