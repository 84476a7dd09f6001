"""Kernel-level timing of the candidate-exchange stages on ONE GPU: R shards of the
headline database held side by side and stepped in lock-step (the all-gathers are
torch.stack).  Prints per-stage CUDA-event times of ONE rank's work.  Development probe."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from instance_search_b200 import ops  # noqa: E402
from instance_search_b200.search import DescriptorIndex, shard_bounds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=8)
ap.add_argument("--rows", type=int, default=1000000)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
bounds = shard_bounds(a.rows, a.R)
shards = [DescriptorIndex(bench.make_rows_slice(a.rows, a.dim, bench.SEED, dev, lo, hi), lo) for lo, hi in bounds]
q = bench.make_rows(a.queries, a.dim, bench.SEED + 100, dev)
offs = torch.tensor([lo for lo, _ in bounds], dtype=torch.int64, device=dev)
kc = min(a.k + ops.DEFAULT_MARGIN, ops.MAX_CANDIDATES)


def ev():
    return torch.cuda.Event(enable_timing=True)


acc = {}
for it in range(2 + a.iters):
    cands = [sh.candidates(q, a.k, kc) for sh in shards[1:]]
    torch.cuda.synchronize()
    e = [ev() for _ in range(6)]
    scr = []
    e[0].record()
    c0 = shards[0].candidates(q, a.k, kc, events=scr)
    e[1].record()
    all_screen = torch.stack([c0[0]] + [c[0] for c in cands])
    e[2].record()
    thr = ops.topk_global_threshold(all_screen)
    e[3].record()
    p0 = shards[0].rerank_owned(q, a.k, c0[0], c0[1], thr)
    e[4].record()
    packed_all = torch.stack([p0] + [sh.rerank_owned(q, a.k, c[0], c[1], thr) for sh, c in zip(shards[1:], cands)])
    torch.cuda.synchronize()
    e5a, e5b = ev(), ev()
    e5a.record()
    s, i, unc_rows, n_unc = ops.topk_merge_certified(packed_all, offs, thr, a.k)
    e5b.record()
    torch.cuda.synchronize()
    if it >= 2:
        t = {"screen(cast+fill+gemm)": scr[0][0].elapsed_time(scr[0][1]), "candidates_total": e[0].elapsed_time(e[1]),
             "global_threshold": e[2].elapsed_time(e[3]), "rerank_owned": e[3].elapsed_time(e[4]),
             "merge_certified": e5a.elapsed_time(e5b)}
        for k_, v in t.items():
            acc.setdefault(k_, []).append(v)
print("R=%d rows/shard=%d Q=%d k=%d kc=%d uncertified=%d" % (a.R, bounds[0][1], a.queries, a.k, kc, int(n_unc.item())))
for k_, v in acc.items():
    print("%-26s %8.3f ms" % (k_, sorted(v)[len(v) // 2]))
