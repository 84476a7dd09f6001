"""Where the roles of the CTA-pair screen kernel wait.  Needs the debug build:

    make -C instance_search_b200/csrc timeline          # -> instance_search_b200/libisb_timeline.so
    python tools/screen_timeline.py --N 125000

Runs one search and prints, per slot of isb_timeline (isb_gemm_core.cuh), the mean over the
leader / peer CTAs as cycles and as a share of the kernel's own duration."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--Q", type=int, default=10000)
ap.add_argument("--N", type=int, default=125000)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libisb_timeline.so")   # the instrumented build
L = _lib.lib()
try:
    L.isb_debug_timeline
except AttributeError:
    sys.exit("libisb_timeline.so lacks isb_debug_timeline: build it with `make timeline`")
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
db = torch.randn(a.N, a.D, device=dev, generator=g)
db /= db.norm(dim=1, keepdim=True)
q = torch.randn(a.Q, a.D, device=dev, generator=g)
q /= q.norm(dim=1, keepdim=True)
db16 = ops.to_bf16(db)
margin = min(28, 128 - a.k)
for _ in range(2):
    ops.topk_search(q, db, db16, a.k, margin)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (256 * 8))()
_lib.check(L.isb_debug_timeline(buf), "isb_debug_timeline")
t = torch.tensor(list(buf), dtype=torch.float64).view(256, 8)[:148]
names = ["kernel total", "producer: wave gate", "producer: free stage", "MMA: accumulator drained",
         "MMA: stage loaded", "epilogue w2: accumulator ready", "epilogue w2: tile()", "tiles"]
for who, rows in (("leader CTAs", t[0::2]), ("peer CTAs", t[1::2])):
    total = rows[:, 0].mean().item()
    print(who, "(mean over %d)" % rows.size(0))
    for s, name in enumerate(names):
        v = rows[:, s].mean().item()
        print("  %-32s %14.0f  %5.1f %%" % (name, v, 100.0 * v / total if s < 7 and total else float("nan")))
