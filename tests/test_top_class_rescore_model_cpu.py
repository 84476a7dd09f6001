"""CPU model of the class-max re-score of the region head (csrc/isb_regions.cu RowTopEpilogue /
region_select_top_kernel): the bf16 screen keeps the four best class values of every window and the
classes of the best three; the class-max is re-scored from those three, and the window is certified
when the re-scored maximum clears the fourth-best SCREEN value by 8 sigma of the screen noise.
Property: a certified window's class-max is the true one; windows whose leading classes are tied
inside the noise are not certified; the packed (value, class) key orders like the value.  No GPU."""

import torch


def rescore(true_logits, noise_sigma, g):
    screen = true_logits + noise_sigma * torch.randn(true_logits.shape, generator=g)
    top = screen.topk(min(4, screen.size(1)), dim=1)
    m4 = top.values[:, 3] if screen.size(1) >= 4 else torch.full((screen.size(0),), float("-inf"))
    best3 = top.indices[:, :3]
    rescored = true_logits.gather(1, best3)                 # fp32-grade logits of the best three classes
    cm = rescored.max(1).values
    d = top.values[:, :3] - rescored
    sigma = d.pow(2).mean().sqrt().clamp_min(0.5 * noise_sigma)
    certified = (cm - m4) > 8.0 * sigma
    return cm, certified


def test_certified_windows_have_the_true_class_max():
    g = torch.Generator().manual_seed(1)
    W, J, sigma = 20000, 464, 1e-3
    logits = 0.41 * torch.randn(W, J, generator=g)          # the benchmark's logit scale
    cm, cert = rescore(logits, sigma, g)
    true = logits.max(1).values
    assert bool((cm[cert] == true[cert]).all())
    assert float(cert.float().mean()) > 0.995                # class gaps ~0.1 >> 8 sigma: nearly all certify
    # noise as large as the class gaps: most windows are NOT certified, the certified ones still right
    cm2, cert2 = rescore(logits, 0.05, g)
    assert bool((cm2[cert2] == true[cert2]).all())
    assert float(cert2.float().mean()) < 0.5


def test_tied_leading_classes_are_not_certified():
    g = torch.Generator().manual_seed(2)
    logits = 0.41 * torch.randn(500, 40, generator=g)
    logits[:, :5] = logits.max(1, keepdim=True).values + 1.0     # five classes tied at the top
    cm, cert = rescore(logits, 1e-3, g)
    assert not bool(cert.any())
    logits[:, 3:5] -= 0.5                                         # three tied: the best three contain the maximum
    cm, cert = rescore(logits, 1e-3, g)
    assert bool(cert.all()) and bool((cm == logits.max(1).values).all())


def test_packed_value_class_key_orders_like_the_value():
    # the epilogue replaces the low 9 mantissa bits of a value by the class index and inserts the
    # packed float with fmin / fmax: the order of two packed keys is the order of the values unless
    # they agree in all but those 9 bits (relative difference < 2^-14)
    g = torch.Generator().manual_seed(3)
    v = torch.randn(100000, generator=g) * 3.0
    cls = torch.randint(0, 464, (100000,), generator=g, dtype=torch.int32)
    packed = ((v.view(torch.int32) & ~0x1FF) | cls).view(torch.float32)
    assert bool(torch.isfinite(packed).all())
    assert float(((packed - v).abs() / v.abs()).max()) < 2.0 ** -14
    a, b = v[:-1], v[1:]
    decided = (a - b).abs() > 2.0 ** -13 * torch.maximum(a.abs(), b.abs())
    assert bool(((packed[:-1] > packed[1:]) == (a > b))[decided].all())
    assert bool((((packed.view(torch.int32) & 0x1FF)) == cls).all())
