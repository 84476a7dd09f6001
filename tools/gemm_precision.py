"""Where the error of the split-operand projection comes from: relative L2 error (per row, vs an
fp64 product of the fp32 inputs) of gemm_nt_split at the whitening shape, as a function of the
number of split-K partitions (= length of the fp32 accumulation chain inside the tensor core),
and of the plain 1-term product and torch's fp32 matmul for scale."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=256)
ap.add_argument("--N", type=int, default=2048)
ap.add_argument("--K", type=int, default=100352)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
# the projection's operands: U = sum of 6 unit-norm crops (+ small shift), W ~ randn / sqrt(K)
U = torch.relu(torch.randn(a.M, a.K, device=dev, generator=g))
U = 6.0 * U / U.norm(dim=1, keepdim=True) + 0.01 * torch.randn(a.K, device=dev, generator=g)
W = torch.randn(a.N, a.K, device=dev, generator=g) / a.K ** 0.5
ref = torch.empty(a.M, a.N, dtype=torch.float64, device=dev)
for s in range(0, a.N, 256):
    ref[:, s:s + 256] = U.double() @ W[s:s + 256].double().t()


def err(y):
    e = (y.double() - ref).norm(dim=1) / ref.norm(dim=1)
    return {"max_rel_l2": float(e.max()), "mean_rel_l2": float(e.mean())}


out = {}
u_hi, u_lo, w_hi, w_lo = ops.to_bf16(U, 0), ops.to_bf16(U, 1), ops.to_bf16(W, 0), ops.to_bf16(W, 1)
# error of the split representation itself (no accumulation error): fp64 product of the bf16 terms
rep = torch.zeros_like(ref)
for s in range(0, a.N, 256):
    wh, wl = w_hi[s:s + 256, :a.K].double(), w_lo[s:s + 256, :a.K].double()
    uh, ul = u_hi[:, :a.K].double(), u_lo[:, :a.K].double()
    rep[:, s:s + 256] = uh @ wh.t() + ul @ wh.t() + uh @ wl.t()
out["three_terms_in_fp64"] = err(rep)
for splits in (1, 4, 9, 18, 37, 74, 148):
    out["split3_splits%d" % splits] = err(ops.gemm_nt_split(u_hi, u_lo, w_hi, w_lo, splits=splits, k=a.K))
out["plain1_splits9"] = err(ops.gemm_nt(u_hi, w_hi, splits=9, k=a.K))
torch.backends.cuda.matmul.allow_tf32 = False
out["torch_fp32_matmul"] = err(U @ W.t())
print(json.dumps(out, indent=1))
