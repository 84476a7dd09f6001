"""Timing of all-pairs similarities + per-couple negative selection (BASELINE
configs[2]): N = 16384 descriptors of dimension 2048, labels i // 16, one couple per
anchor.  Development / profiling probe; the contract bench is bench.py."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import _lib, mining  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=16384)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--per", type=int, default=16)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--terms", type=int, default=1)
ap.add_argument("--options", default="", help="library options, e.g. 'mining_kc=8'")
a = ap.parse_args()
for kv in [t for t in a.options.split(",") if t]:
    name, val = kv.split("=")
    _lib.set_option(name, int(val))
dev = torch.device("cuda:0")
peaks = {"bf16_tflops": 1590.0}
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peaks = json.load(open(pp))
g = torch.Generator(device=dev).manual_seed(1234 + 3)
lab = torch.arange(a.N, device=dev) // a.per
centers = torch.randn(int(lab.max()) + 1, a.D, device=dev, generator=g)
E = centers[lab] + 0.5 * torch.randn(a.N, a.D, device=dev, generator=g)
E = E / E.norm(dim=1, keepdim=True)
anchors = torch.arange(a.N, device=dev)
positives = (anchors // a.per) * a.per + (anchors % a.per + 1) % a.per
idx = mining.MiningIndex(E, lab.int(), terms=a.terms)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {"workload": "negative mining, N=%d D=%d, %d per label, one couple per anchor, terms=%d" %
                   (a.N, a.D, a.per, a.terms), "options": a.options}
sus = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
flops = 2.0 * a.N * a.N * a.D
for semi in (True, False):
    ms = timeit(lambda: idx.select_negatives(anchors, positives, semi))
    out["semi_hard" if semi else "hard"] = {
        "ms": ms, "anchors_per_s": a.N / (ms * 1e-3), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
        "frac_of_sustained_peak_algorithmic": flops / (ms * 1e-3) / 1e12 / sus,
        "second_line_couples": idx.last_second_line, "bruteforce_rows": int(idx.last_bruteforce)}
ms = timeit(lambda: mining.all_pairs_similarities(E, terms=a.terms))
out["all_pairs_matrix"] = {"ms": ms, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                           "issued_tflops": flops * a.terms / (ms * 1e-3) / 1e12}
print(json.dumps(out), flush=True)
