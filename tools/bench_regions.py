"""Stage-by-stage timing of the region-descriptor head (BASELINE configs[1]):
B x 2048 x H x W fp32 feature maps -> [B, D] descriptors.  Prints one JSON line per
map size with per-stage CUDA-event times, region descriptors/s, the HBM fraction of
the bandwidth-bound stages and the tensor fraction of the projection.
Synthetic inputs per SURVEY.md 8d.  Development / profiling probe; the contract
bench is bench.py."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import _lib, regions  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--C", type=int, default=2048)
ap.add_argument("--sizes", default="14,32")
ap.add_argument("--ncls", type=int, default=464)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--k", type=int, default=6)
ap.add_argument("--terms", type=int, default=3)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cls-out", type=int, default=0, help="1: also produce cls_out (training); 0: eval, class-max only")
ap.add_argument("--variants", default="", help="';'-separated tuning variants, each 'option=val,option=val' (pool_stages, "
                "pool_generic_geom, gather_cw / _g / _stages: _lib.OPTIONS); one JSON line per variant and map size")
a = ap.parse_args()

dev = torch.device("cuda:0")
peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peaks = json.load(open(pp))

g = torch.Generator(device=dev).manual_seed(1234 + 2)
Kin = a.C * 49
cls_w = torch.randn(a.ncls, a.C, device=dev, generator=g) / a.C ** 0.5
cls_b = 0.01 * torch.randn(a.ncls, device=dev, generator=g)
shift = 0.01 * torch.randn(Kin, device=dev, generator=g)
lin_w = torch.randn(a.D, Kin, device=dev, generator=g) / Kin ** 0.5
lin_b = 0.01 * torch.randn(a.D, device=dev, generator=g)
hw = regions.HeadWeights(cls_w, cls_b, shift, lin_w, lin_b, terms=a.terms)
del lin_w
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > L2


def ev():
    return torch.cuda.Event(enable_timing=True)


TUNABLES = ("pool_stages", "pool_generic_geom", "gather_cw", "gather_g", "gather_stages", "gather_small", "resc_splits", "tc_debug", "region_top_select")
variants = [v for v in a.variants.split(";")] if a.variants else [""]
for hw_size, variant in [(int(s), v) for s in a.sizes.split(",") for v in variants]:
    H = W = hw_size
    for name in (TUNABLES if a.variants else ()):
        _lib.set_option(name, None)
    for kv in [t for t in variant.split(",") if t]:
        name, val = kv.split("=")
        _lib.set_option(name, int(val))   # read by libisb's launch plans at every call
    x = torch.relu(torch.randn(a.B, a.C, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(77 + hw_size)))
    stages = {"select": [], "gather": [], "logits_fixup": [], "project": [], "total": []}
    changed_total = 0
    uncert = 0
    for it in range(a.warmup + a.iters):
        flush.zero_()   # L2 flush between iterations (x at 14x14 is 411 MB, at 32x32 2.1 GB)
        e = [ev() for _ in range(5)]
        ke = a.k + regions.RUNNER_UPS
        e[0].record()
        idx_e, nsel_e, approx_cls, norm_e, approx_e, runner_up, n1 = regions.region_select(x, hw, ke, (7, 7))
        e[1].record()
        U_hi, U_lo, win_mean = regions.region_gather(x, hw, ke, (7, 7), idx_e, nsel_e, norm_e, k_sum=a.k)
        e[2].record()
        idx, norm, nsel, cls_out, changed, n_changed, n2 = regions.region_logits(
            win_mean, hw, a.k, nsel_e, idx_e, norm_e, approx_e, runner_up, None if a.cls_out else approx_cls)
        regions.region_gather(x, hw, a.k, (7, 7), idx, nsel, norm, want_means=False, out=(U_hi, U_lo),
                              image_list=changed, n_list=n_changed)
        e[3].record()
        desc = regions.region_project(U_hi, U_lo, hw, nsel)
        e[4].record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            for n, (s, t) in zip(["select", "gather", "logits_fixup", "project"], zip(e[:-1], e[1:])):
                stages[n].append(s.elapsed_time(t))
            stages["total"].append(e[0].elapsed_time(e[4]))
            uncert += int((n1[0] + n2[0]).item())
            changed_total += int(n_changed.item())
    med = {n: sorted(v)[len(v) // 2] for n, v in stages.items()}
    # streamed: 8 batches back to back over two alternating inputs, certificates read at the end
    x2 = torch.relu(torch.randn(a.B, a.C, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(78 + hw_size)))
    st = []
    for it in range(2 + 5):
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record()
        flags = [regions.region_descriptors_async((x, x2)[i & 1], hw, a.k, (7, 7), want_cls_out=not not a.cls_out)[4]
                 for i in range(8)]
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            st.append(e0.elapsed_time(e1) / 8)
    med["streamed_per_batch"] = sorted(st)[len(st) // 2]
    del x2
    nwin = (H - 6) * (W - 6)
    kk = min(nwin, a.k)
    units = a.B * kk                                  # region descriptors per batch (SURVEY 8d unit)
    # algorithmic bytes of the bandwidth-bound part (sum-before-projection form):
    #   x read once + classifier weights + the bf16 operand written once (hi, + lo when terms == 3)
    op_bytes = 2 * a.B * Kin * (2 if a.terms == 3 else 1)
    bw_bytes = 4 * a.B * a.C * H * W + 4 * a.ncls * a.C + op_bytes
    bw_ms = med["select"] + med["gather"]
    proj_flops = 2.0 * a.B * Kin * a.D * a.terms     # tensor-pipe flops actually issued
    alg_flops = 2.0 * a.B * Kin * a.D                # one fp32-grade product
    print(json.dumps({
        "workload": "region descriptors, B=%d C=%d %dx%d map, ncls=%d, D=%d, k=%d, terms=%d, cls_out=%d" %
                    (a.B, a.C, H, W, a.ncls, a.D, a.k, a.terms, a.cls_out),
        "variant": variant,
        "ms": med, "uncertified_images": uncert, "regathered_images": changed_total, "region_desc_per_s": units / (med["total"] * 1e-3),
        "images_per_s": a.B / (med["total"] * 1e-3),
        "pool_select_gather": {"algorithmic_bytes": bw_bytes, "ms": bw_ms,
                               "achieved_gbs": bw_bytes / (bw_ms * 1e-3) / 1e9,
                               "frac_of_measured_hbm": bw_bytes / (bw_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "projection": {"issued_tflops": proj_flops / (med["project"] * 1e-3) / 1e12,
                       "algorithmic_tflops": alg_flops / (med["project"] * 1e-3) / 1e12,
                       "frac_of_measured_bf16_burst": proj_flops / (med["project"] * 1e-3) / 1e12 / peaks["bf16_tflops"]},
    }), flush=True)
    del x
