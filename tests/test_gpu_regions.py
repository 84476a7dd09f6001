"""GPU parity: region-descriptor head (a3-a6) through the C ABI vs the oracle and
the fixtures minted from the reference's own RegionDescriptorNet / DescriptorNet."""

import pytest
import torch

import oracle
from conftest import load_golden
from parity import check_descriptors

pytestmark = pytest.mark.gpu

# descriptors: per-row L2 error <= parity.DESC_L2_TOL and 1 - cosine <= 1e-10 against the
# oracle (the projection runs as a 3-product bf16 expansion, fp32-grade): check_descriptors
CLS_RTOL = 1e-5   # cls_out: fp32 dot products of the exact window means (isb_region_logits)
CLS_ATOL = 2e-6   # fp32 summation-order noise on logits that cancel to ~0


@pytest.fixture(scope="module")
def R():
    from instance_search_b200 import regions
    return regions


def _hw(R, g, dev="cuda", terms=3):
    return R.HeadWeights(g["cls_w"].to(dev), g["cls_b"].to(dev), g["shift"].to(dev),
                         g["lin_w"].to(dev), g["lin_b"].to(dev), terms=terms)


@pytest.mark.parametrize("tag", ["7x7", "8x8", "9x12", "14x14"])
def test_region_tiny_golden(R, tag):
    g = load_golden("region_tiny")
    fs = tuple(int(v) for v in g["fsize"])
    hw = _hw(R, g)
    d, c, i, n = R.region_descriptors(g["x_" + tag].cuda(), hw, g["k"], fs)
    assert torch.equal(i.cpu(), g["idx_" + tag])                 # index-exact windows
    assert torch.allclose(c.cpu(), g["cls_out_" + tag], rtol=CLS_RTOL, atol=CLS_ATOL)
    check_descriptors(d, g["desc_" + tag])
    nw = (g["x_" + tag].size(2) - 6) * (g["x_" + tag].size(3) - 6)
    assert n.tolist() == [min(nw, g["k"])] * 3
    assert torch.allclose(d.norm(dim=1).cpu(), torch.ones(3), atol=1e-5)


def test_region_resnet18_head_golden(R):
    g = load_golden("region_resnet18_head")
    hw = _hw(R, g)
    d, c, i, n = R.region_descriptors(g["x"].cuda(), hw, g["k"], tuple(int(v) for v in g["fsize"]))
    assert torch.equal(i.cpu(), g["idx"])
    assert torch.allclose(c.cpu(), g["cls_out"], rtol=CLS_RTOL, atol=CLS_ATOL)
    check_descriptors(d, g["desc"])


def test_descriptor_head_golden(R):
    g = load_golden("descriptor_tiny")
    hw = R.HeadWeights(None, None, g["shift"].cuda(), g["lin_w"].cuda(), g["lin_b"].cuda())
    d = R.global_descriptors(g["x"].cuda(), hw)
    check_descriptors(d, g["desc"])


def _synthetic(B, C, H, W, ncls, D, seed):
    # relu(randn) maps, Wc ~ randn/sqrt(C), W ~ randn/sqrt(Kin)
    g = torch.Generator().manual_seed(seed)
    Kin = C * 49
    return dict(x=torch.relu(torch.randn(B, C, H, W, generator=g)),
                cls_w=torch.randn(ncls, C, generator=g) / C ** 0.5,
                cls_b=0.01 * torch.randn(ncls, generator=g),
                shift=0.01 * torch.randn(Kin, generator=g),
                lin_w=torch.randn(D, Kin, generator=g) / Kin ** 0.5,
                lin_b=0.01 * torch.randn(D, generator=g))


@pytest.mark.parametrize("B,C,H,W,ncls,D,k", [
    (5, 64, 14, 14, 20, 32, 6),
    (3, 100, 10, 17, 7, 24, 6),      # C not a multiple of the channel block, ragged map
    (2, 2048, 14, 14, 464, 64, 6),   # the reference's ResNet-152 head shape (config 2a), small D
    (2, 256, 32, 32, 464, 32, 6),    # 676 windows per image (config 2b map size)
    (4, 64, 9, 9, 10, 16, 12),       # k larger than the number of windows (9)
    (3, 100, 14, 14, 12, 24, 6),     # small-map gather kernel, C not a multiple of its 32-channel step
    (3, 72, 12, 16, 9, 16, 5),       # small-map gather kernel, run-time map width
    (2, 40, 7, 8, 5, 16, 6),         # 2 windows, 56 pixels
])
def test_region_random_vs_oracle(R, B, C, H, W, ncls, D, k):
    s = _synthetic(B, C, H, W, ncls, D, seed=B * 1000 + C)
    hw = _hw(R, s)
    d, c, i, n = R.region_descriptors(s["x"].cuda(), hw, k, (7, 7))
    od, oc, oi, on = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"],
                                                      s["lin_w"], s["lin_b"], k, (7, 7))
    assert torch.equal(n.cpu().long(), on)
    assert torch.equal(i.cpu(), oi)
    assert torch.allclose(c.cpu(), oc, rtol=CLS_RTOL, atol=CLS_ATOL)
    check_descriptors(d, od)


def test_region_plain_bf16_projection_is_close(R):
    # terms=1 (north_star's plain bf16 x bf16 projection): same windows, looser descriptors
    s = _synthetic(4, 128, 14, 14, 30, 64, seed=77)
    d3, _, i3, _ = R.region_descriptors(s["x"].cuda(), _hw(R, s, terms=3), 6, (7, 7))
    d1, _, i1, _ = R.region_descriptors(s["x"].cuda(), _hw(R, s, terms=1), 6, (7, 7))
    assert torch.equal(i1, i3)
    assert (d1 - d3).abs().max().item() < 5e-3
    assert (d1 * d3).sum(1).min().item() > 1 - 1e-4      # cosine between the two


def test_region_exact_second_line_matches_fast_path_and_oracle(R):
    # the fp64 second line (exact_mode) alone, and the stats of the certified fast path
    s = _synthetic(3, 96, 12, 15, 25, 24, seed=5)
    hw = _hw(R, s)
    x = s["x"].cuda()
    stats = {}
    d, c, i, n = R.region_descriptors(x, hw, 6, (7, 7), stats=stats)
    assert stats == {"batches": 1, "batches_resolved_exactly": 0, "images_resolved_exactly": 0}
    i2, n2, c2, wn2, _, _, _ = R.region_select(x, hw, 6, (7, 7), exact_mode=True)
    torch.cuda.synchronize()
    od, oc, oi, on = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"],
                                                      s["lin_w"], s["lin_b"], 6, (7, 7))
    assert torch.equal(i2.cpu(), oi) and torch.equal(i.cpu(), oi)
    assert torch.allclose(c2.cpu(), oc, rtol=CLS_RTOL, atol=CLS_ATOL)
    assert torch.allclose(c.cpu(), oc, rtol=CLS_RTOL, atol=CLS_ATOL)
    U_hi, U_lo, _ = R.region_gather(x, hw, 6, (7, 7), i2, n2, wn2, want_means=False)
    check_descriptors(R.region_project(U_hi, U_lo, hw, n2), od)


def test_region_near_tied_windows_fall_to_the_exact_line(R):
    # a constant map: every window has the same score -> no certificate can hold ->
    # the batch is redone exactly; ties -> lower window index first (the oracle's
    # topk does not define a tie order, so only the set of scores is compared)
    C, H, W, ncls, D = 32, 10, 10, 6, 16
    s = _synthetic(2, C, H, W, ncls, D, seed=6)
    s["x"] = torch.ones(2, C, H, W) * 0.5
    hw = _hw(R, s)
    stats = {}
    d, c, i, n = R.region_descriptors(s["x"].cuda(), hw, 6, (7, 7), stats=stats)
    assert stats["batches_resolved_exactly"] == 1
    assert i.cpu().tolist() == [[0, 1, 2, 3, 4, 5]] * 2
    od, oc, _, _ = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"], s["lin_w"],
                                                    s["lin_b"], 6, (7, 7))
    check_descriptors(d, od)      # identical crops: same descriptor
    assert torch.allclose(c.cpu(), oc, rtol=CLS_RTOL, atol=CLS_ATOL)


def test_region_only_the_uncertified_images_are_redone(R):
    # one constant (all windows tied) image inside an ordinary batch: the certificates list
    # that image alone, it alone goes through the exact second line, and every image of the
    # batch still matches the oracle
    C, H, W, ncls, D = 32, 12, 11, 6, 16
    s = _synthetic(5, C, H, W, ncls, D, seed=8)
    s["x"][3] = 0.25
    hw = _hw(R, s)
    stats = {}
    d, c, i, n = R.region_descriptors(s["x"].cuda(), hw, 6, (7, 7), stats=stats)
    assert stats == {"batches": 1, "batches_resolved_exactly": 1, "images_resolved_exactly": 1}
    od, oc, oi, on = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"], s["lin_w"],
                                                      s["lin_b"], 6, (7, 7))
    keep = [0, 1, 2, 4]
    assert torch.equal(i.cpu()[keep], oi[keep])
    assert i.cpu()[3].tolist() == [0, 1, 2, 3, 4, 5]           # ties -> lower window index first
    check_descriptors(d, od)
    assert torch.allclose(c.cpu()[keep], oc[keep], rtol=CLS_RTOL, atol=CLS_ATOL)


@pytest.mark.parametrize("B,C,H,W,ncls,D", [(6, 64, 14, 14, 20, 32), (3, 256, 20, 24, 464, 32)])
def test_region_eval_path_without_cls_out(R, B, C, H, W, ncls, D):
    # eval (descriptor only): class-max from the contending classes only -- same windows, same descriptor
    s = _synthetic(B, C, H, W, ncls, D, seed=31)
    hw = _hw(R, s)
    x = s["x"].cuda()
    d0, c0, i0, n0 = R.region_descriptors(x, hw, 6, (7, 7))
    d1, c1, i1, n1 = R.region_descriptors(x, hw, 6, (7, 7), want_cls_out=False)
    assert c1 is None and torch.equal(i1, i0) and torch.equal(n1, n0) and torch.equal(d1, d0)


@pytest.mark.parametrize("H,W", [(13, 13), (16, 20)])
def test_region_exact_line_is_complete_on_maps_with_more_than_32_windows(R, H, W):
    # window scores separated by far less than the bf16 screen can see (a nearly constant map with a
    # 2e-3 ramp under one bf16 ulp) on a map with 49 / 140 windows: the first round's 32 candidates are an arbitrary
    # subset, so the exact second line must keep scoring until its certificate holds -- it may
    # never return a list it has flagged incomplete (round 1 did)
    C, ncls, D, k = 32, 6, 16, 6
    s = _synthetic(2, C, H, W, ncls, D, seed=9)
    # raster ramp of amplitude 1.8e-3 on 0.5 (below half a bf16 ulp of 0.5, so every pooled value
    # rounds to the same bf16 number), ascending on image 0 and descending on image 1: consecutive
    # windows differ by 1.8e-3 / (H*W) = 5e-6 .. 1e-5 in the fp32 pooled mean
    ramp = (torch.arange(H * W, dtype=torch.float64) / (H * W)).view(1, 1, H, W)
    ramp = torch.cat([ramp, 0.9 * (1.0 - ramp)], 0)
    s["x"] = (0.5 + 1.8e-3 * ramp).float().expand(2, C, H, W).contiguous()
    s["cls_w"] = s["cls_w"].abs() / s["cls_w"].abs().sum(1, keepdim=True)     # slope 1 for every class
    hw = _hw(R, s)
    x = s["x"].cuda()
    od, oc, oi, on = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"], s["lin_w"],
                                                      s["lin_b"], k, (7, 7))
    # the oracle's own top-k must be decided well above fp32 noise for the comparison to mean anything
    c = torch.nn.functional.conv2d(torch.nn.functional.avg_pool2d(s["x"], 7, stride=1),
                                   s["cls_w"].view(ncls, C, 1, 1), s["cls_b"]).max(1).values.reshape(2, -1)
    top = c.sort(dim=1, descending=True).values
    assert float((top[:, :k] - top[:, 1:k + 1]).min()) > 2e-6
    i2, n2, c2, wn2, _, _, n_unc = R.region_select(x, hw, k, (7, 7), exact_mode=True)
    assert int(n_unc[0]) == 0                                   # complete by construction
    assert torch.equal(i2.cpu(), oi)
    assert torch.allclose(c2.cpu(), oc, rtol=CLS_RTOL, atol=CLS_ATOL)
    stats = {}
    d, c1, i1, n1 = R.region_descriptors(x, hw, k, (7, 7), stats=stats)       # the certified front door
    assert torch.equal(i1.cpu(), oi)
    check_descriptors(d, od)


@pytest.mark.parametrize("B,C,H,W,k,k_sum", [(5, 200, 14, 14, 8, 6), (3, 64, 12, 16, 5, 5), (4, 2048, 14, 14, 8, 6),
                                             (2, 96, 8, 8, 8, 3), (3, 256, 32, 32, 8, 6), (2, 100, 20, 24, 7, 7),
                                             (4, 72, 32, 28, 3, 2)])
def test_small_map_gather_kernel_equals_the_channel_stream_kernel(R, B, C, H, W, k, k_sum):
    # the plane-block kernel for maps of <= 256 pixels and the row-range kernel for maps of <= 1024
    # pixels (option gather_small, default on) against the general kernel.  Small maps: the operand
    # sums the same terms in the same order -> bit-identical bf16 hi / lo; the row-range kernel adds
    # the windows pairwise -> equal to fp32 rounding (hi + lo compared at 2^-16 of the largest entry).
    # The window means add their 49 terms column-wise instead of row-wise -> equal to fp32 rounding
    from instance_search_b200 import _lib
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.relu(torch.randn(B, C, H, W, generator=g)).cuda()
    nwin = (H - 6) * (W - 6)
    hw = R.HeadWeights(None, None, 0.01 * torch.randn(C * 49, generator=g).cuda(),
                       torch.randn(8, C * 49, generator=g).cuda(), None)
    idx = torch.stack([torch.randperm(nwin, generator=g)[:k] if nwin >= k else
                       torch.arange(k) % nwin for _ in range(B)]).cuda()
    nsel = torch.tensor([min(k, nwin) - (b % 2) for b in range(B)], dtype=torch.int32).cuda()
    norm = (1.0 + torch.rand(B, k, generator=g)).cuda()
    with _lib.options(gather_small=0):
        h0, l0, m0 = R.region_gather(x, hw, k, (7, 7), idx, nsel, norm, k_sum=k_sum)
    h1, l1, m1 = R.region_gather(x, hw, k, (7, 7), idx, nsel, norm, k_sum=k_sum)
    if H * W <= 256:
        assert torch.equal(h0.view(torch.int16), h1.view(torch.int16))
        assert torch.equal(l0.view(torch.int16), l1.view(torch.int16))
    else:
        u0, u1 = h0.float() + l0.float(), h1.float() + l1.float()
        assert float((u0 - u1).abs().max()) <= 2.0 ** -16 * float(u0.abs().max())
        assert float((h0.float() - h1.float()).abs().max()) <= 2.0 ** -7 * float(u0.abs().max())
    for b in range(B):
        n = int(nsel[b])
        assert torch.allclose(m0[b, :n], m1[b, :n], rtol=2e-6, atol=1e-7)
    # ... and the fix-up form (image list, no means) rewrites exactly the listed rows
    lst = torch.tensor([B - 1, 0] + [0] * (B - 2), dtype=torch.int32).cuda()
    n_list = torch.tensor([2], dtype=torch.int32).cuda()
    h2, l2 = torch.zeros_like(h1), torch.zeros_like(l1)
    R.region_gather(x, hw, k, (7, 7), idx, nsel, norm, k_sum=k_sum, want_means=False, out=(h2, l2),
                    image_list=lst, n_list=n_list)
    assert torch.equal(h2[0].view(torch.int16), h1[0].view(torch.int16))
    assert torch.equal(l2[B - 1].view(torch.int16), l1[B - 1].view(torch.int16))
    if B > 2:
        assert int(h2[1].view(torch.int16).abs().sum()) == 0


@pytest.mark.parametrize("B,C,H,W,ncls,k", [(6, 128, 14, 14, 40, 6), (3, 64, 20, 24, 464, 6), (4, 96, 12, 12, 2, 4),
                                            (2, 256, 32, 32, 100, 6)])
def test_one_kernel_reselect_equals_the_four_launch_chain(R, B, C, H, W, ncls, k):
    # option region_top_select (default on): candidates + fp32-grade re-score of each window's best three
    # classes + final selection in ONE kernel, against the chain with the tensor-core re-score of all
    # classes: same windows, same certificates, descriptors equal to fp32 rounding, and the oracle's
    from instance_search_b200 import _lib
    s = _synthetic(B, C, H, W, ncls, 24, seed=B * 7 + ncls)
    hw = _hw(R, s)
    x = s["x"].cuda()
    with _lib.options(region_top_select=0):
        d0, c0, i0, n0 = R.region_descriptors(x, hw, k, (7, 7), want_cls_out=False)
        e0 = R.region_select(x, hw, k + 2, (7, 7))
    st = {}
    d1, c1, i1, n1 = R.region_descriptors(x, hw, k, (7, 7), want_cls_out=False, stats=st)
    e1 = R.region_select(x, hw, k + 2, (7, 7))
    assert torch.equal(i0, i1) and torch.equal(n0, n1)
    assert torch.equal(e0[0], e1[0])                                    # the k + 2 windows handed to the gather
    assert torch.allclose(e0[3], e1[3], rtol=1e-6)                      # crop norms
    assert torch.allclose(e0[4], e1[4], rtol=1e-5, atol=2e-6)           # fp32-grade class-max
    assert int(e1[6][0]) == 0 and st.get("images_resolved_exactly", 0) == 0
    check_descriptors(d1, d0)
    od, oc, oi, on = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"], s["lin_w"],
                                                      s["lin_b"], k, (7, 7))
    assert torch.equal(i1.cpu(), oi)
    check_descriptors(d1, od)


def test_one_kernel_reselect_reports_windows_with_more_than_three_contending_classes(R):
    # five identical, dominant classifier rows: the four best classes of every window are tied, so the class-max
    # cannot be settled from the best three -> every image is listed, the exact second line answers,
    # and the result is still the oracle's
    s = _synthetic(4, 64, 14, 14, 12, 16, seed=41)
    s["cls_w"][0:5] = s["cls_w"][0].abs() * 3.0      # positive on relu features: these five classes win, tied
    s["cls_b"][0:5] = 0.5
    hw = _hw(R, s)
    x = s["x"].cuda()
    n_unc = R.region_select(x, hw, 8, (7, 7))[6]
    assert int(n_unc[0]) == 4
    st = {}
    d, c, i, n = R.region_descriptors(x, hw, 6, (7, 7), stats=st)
    assert st["images_resolved_exactly"] == 4
    od, oc, oi, on = oracle.region_descriptor_forward(s["x"], s["cls_w"], s["cls_b"], s["shift"], s["lin_w"],
                                                      s["lin_b"], 6, (7, 7))
    o_max = oc.max(1).values
    assert torch.allclose(c.cpu().max(1).values, o_max, rtol=CLS_RTOL, atol=CLS_ATOL)
    assert torch.equal(i.cpu(), oi)
    check_descriptors(d, od)


def test_nan_features_do_not_leak_into_other_images(R):
    # a NaN in one image's map poisons that image's logits only: the windows chosen for the other
    # images are those of the clean batch (the top-class epilogue maps a NaN logit to "no class")
    s = _synthetic(3, 64, 14, 14, 20, 16, seed=77)
    hw = _hw(R, s)
    clean = R.region_select(s["x"].cuda(), hw, 8, (7, 7))
    dirty_x = s["x"].clone()
    dirty_x[1, 5, 3, 4] = float("nan")
    dirty = R.region_select(dirty_x.cuda(), hw, 8, (7, 7))
    torch.cuda.synchronize()
    for b in (0, 2):
        assert torch.equal(clean[0][b], dirty[0][b])
        assert torch.equal(clean[4][b], dirty[4][b])


def test_graphed_region_descriptors_replay_equals_the_eager_path(R):
    # the whole chain captured into one CUDA graph: bit-identical to the eager calls, also after
    # the input buffer has been refilled
    s = _synthetic(6, 128, 14, 14, 40, 32, seed=5)
    hw = _hw(R, s)
    x = s["x"].cuda().contiguous()
    graphed = R.GraphedRegionDescriptors(x, hw, 6, (7, 7))
    for seed in (5, 6):
        x.copy_(_synthetic(6, 128, 14, 14, 40, 32, seed=seed)["x"])
        d, c, i, n, unc = graphed.replay()
        torch.cuda.synchronize()
        d0, c0, i0, n0, unc0 = R.region_descriptors_async(x, hw, 6, (7, 7), want_cls_out=False)
        assert torch.equal(d, d0) and torch.equal(i, i0) and torch.equal(n, n0)
        assert int(unc[:, 0].sum()) == int(unc0[:, 0].sum()) == 0
        od, oc, oi, on = oracle.region_descriptor_forward(x.cpu(), s["cls_w"], s["cls_b"], s["shift"], s["lin_w"],
                                                          s["lin_b"], 6, (7, 7))
        assert torch.equal(i.cpu(), oi)
        check_descriptors(d, od)
