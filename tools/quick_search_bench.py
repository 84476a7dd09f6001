"""Development probe (not the contract bench): time isb_topk_search at a given
scale on one GPU and spot-check a few queries against torch fp64 on the GPU."""
import argparse
import json
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--Q", type=int, default=10000)
ap.add_argument("--N", type=int, default=1000000)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--check", type=int, default=32)
ap.add_argument("--options", default="", help="library options, e.g. 'screen_wavesync=0,screen_seed=1'")
a = ap.parse_args()
for kv in [t for t in a.options.split(",") if t]:
    name, val = kv.split("=")
    _lib.set_option(name, int(val))

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
db = torch.randn(a.N, a.D, device=dev, generator=g)
db /= db.norm(dim=1, keepdim=True)
q = torch.randn(a.Q, a.D, device=dev, generator=g)
q /= q.norm(dim=1, keepdim=True)
t0 = time.time()
db16 = ops.to_bf16(db)
torch.cuda.synchronize()
print("to_bf16 %.1f ms" % ((time.time() - t0) * 1e3))
margin = min(28, 128 - a.k)
ws = ops.topk_search_workspace(a.Q, a.N, a.D, a.k, margin, dev)
print("workspace %.1f MB" % (ws.numel() / 1e6))
times = []
for it in range(a.iters + 1):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    s, i = ops.topk_search(q, db, db16, a.k, margin, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
print("times ms", ["%.2f" % t for t in times])
best = min(times[1:])
flops = 2.0 * a.Q * a.N * a.D
print(json.dumps({"Q": a.Q, "N": a.N, "D": a.D, "k": a.k, "options": a.options, "ms": best, "qps": a.Q / best * 1e3,
                  "tflops": flops / best / 1e9}))
if a.check:
    c = min(a.check, a.Q)
    sim = q[:c].double() @ db.double().t()
    v, ix = sim.topk(a.k, dim=1)
    ok_i = (ix == i[:c]).all(dim=1).float().mean().item()
    print("rows index-exact vs fp64: %.3f ; max |score diff| %.3g" %
          (ok_i, (v - s[:c].double()).abs().max().item()))
    if ok_i < 1.0:
        bad = (ix != i[:c]).any(dim=1).nonzero().flatten()[:5].tolist()
        for r in bad:
            pos = (ix[r] != i[r]).nonzero().flatten()[:6].tolist()
            print(" row", r, "pos", pos, "want", ix[r, pos].tolist(), "got", i[r, pos].tolist())
