"""CPU model test of the fused head's backward (instance_search_b200/regions_autograd.py):
the decomposition the CUDA path uses -- l2norm backward on z, the four dense products, the
per-window crop statistics (contract of isb_region_crop_stats) and the gather-form crop
gradient (contract of isb_region_scatter_grad) -- restated in torch and compared with autograd
through the oracle's forward_single.  Validates the algebra without a GPU; the kernels
themselves are checked on the device against the composed path (tests/test_gpu_dropin.py)."""

import torch

import oracle


def crop_stats_model(x, k, fsize, idx, nsel, g_u):
    """contract of isb_region_crop_stats"""
    B, C, H, W = x.shape
    fh, fw = fsize
    Wo = W - fw + 1
    n2, dot, means = torch.zeros(B, k), torch.zeros(B, k), torch.zeros(B, k, C)
    for b in range(B):
        for i in range(int(nsel[b])):
            r0, c0 = int(idx[b, i]) // Wo, int(idx[b, i]) % Wo
            crop = x[b, :, r0:r0 + fh, c0:c0 + fw]
            n2[b, i] = (crop.double() ** 2).sum()
            dot[b, i] = (crop.reshape(-1).double() * g_u[b].double()).sum()
            means[b, i] = crop.mean(dim=(1, 2))
    return n2, dot, means


def scatter_grad_model(x, k, fsize, idx, nsel, g_u, n2, dot, g_mean, eps=1e-10):
    """contract of isb_region_scatter_grad (gather form)"""
    B, C, H, W = x.shape
    fh, fw = fsize
    Wo = W - fw + 1
    g_x = torch.zeros_like(x)
    for b in range(B):
        for i in range(int(nsel[b])):
            r0, c0 = int(idx[b, i]) // Wo, int(idx[b, i]) % Wo
            n = (n2[b, i] + eps).sqrt()
            crop = x[b, :, r0:r0 + fh, c0:c0 + fw]
            g = g_u[b].view(C, fh, fw) / n - crop * dot[b, i] / n ** 3
            g = g + g_mean[b, i].view(C, 1, 1) / (fh * fw)
            g_x[b, :, r0:r0 + fh, c0:c0 + fw] += g
    return g_x


def test_backward_decomposition_equals_autograd_of_the_oracle_forward():
    g = torch.Generator().manual_seed(0)
    B, C, H, W, ncls, D, k, fs = 3, 5, 9, 10, 4, 6, 4, (7, 7)
    Kin = C * 49
    x = torch.relu(torch.randn(B, C, H, W, generator=g)).requires_grad_(True)
    cls_w = (torch.randn(ncls, C, generator=g) / C ** 0.5).requires_grad_(True)
    cls_b = (0.01 * torch.randn(ncls, generator=g)).requires_grad_(True)
    shift = (0.01 * torch.randn(Kin, generator=g)).requires_grad_(True)
    lin_w = (torch.randn(D, Kin, generator=g) / Kin ** 0.5).requires_grad_(True)
    lin_b = (0.01 * torch.randn(D, generator=g)).requires_grad_(True)
    desc, cls_out, idx, nsel = oracle.region_descriptor_forward(x, cls_w, cls_b, shift, lin_w, lin_b, k, fs)
    g_desc, g_cls = torch.randn(B, D, generator=g), torch.randn(B, ncls, k, generator=g)
    (desc * g_desc).sum().backward(retain_graph=True)
    (cls_out * g_cls).sum().backward()

    # ---- the decomposition of regions_autograd.RegionHeadFunction.backward, in torch
    with torch.no_grad():
        Wo = W - 6
        u = torch.zeros(B, Kin)
        for b in range(B):
            for i in range(int(nsel[b])):
                r0, c0 = int(idx[b, i]) // Wo, int(idx[b, i]) % Wo
                u[b] += oracle.normalize_l2(x[b, :, r0:r0 + 7, c0:c0 + 7].reshape(1, -1))[0]
            u[b] += float(nsel[b]) * shift
        nf = nsel.float().unsqueeze(1)
        z = u @ lin_w.t() + nf * lin_b
    zr = z.clone().requires_grad_(True)
    oracle.normalize_l2(zr).backward(g_desc)
    g_z = zr.grad                                              # isb_l2norm_rows_backward(z, g_desc)
    with torch.no_grad():
        g_b = (g_z * nf).sum(0)
        g_w = g_z.t() @ u
        g_u = g_z @ lin_w
        g_s = (g_u * nf).sum(0)
        n2, dot, means = crop_stats_model(x, k, fs, idx, nsel, g_u)
        live = torch.arange(k).unsqueeze(0) < nsel.unsqueeze(1)
        G = (g_cls.permute(0, 2, 1) * live.unsqueeze(2)).reshape(B * k, ncls)
        g_cb = G.sum(0)
        g_cw = G.t() @ means.reshape(B * k, C)
        g_mean = (G @ cls_w).reshape(B, k, C)
        g_x = scatter_grad_model(x, k, fs, idx, nsel, g_u, n2, dot, g_mean)
    for name, got, want in (("lin_b", g_b, lin_b.grad), ("lin_w", g_w, lin_w.grad), ("shift", g_s, shift.grad),
                            ("cls_b", g_cb, cls_b.grad), ("cls_w", g_cw, cls_w.grad), ("x", g_x, x.grad)):
        scale = want.abs().max().item()
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5 * scale), (name, (got - want).abs().max().item(), scale)
