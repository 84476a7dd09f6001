"""Per-stage CUDA-event timing of ShardedIndex.search under torchrun (development probe):
wraps the stage hooks and the all-gathers of the candidate exchange.
    python -m torch.distributed.run --nproc-per-node N tools/profile_sharded.py [--rows 1000000]"""
import argparse
import collections
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from instance_search_b200.search import ShardedIndex, shard_bounds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1000000)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
lo, hi = shard_bounds(a.rows, world)[rank]
index = ShardedIndex(bench.make_rows_slice(a.rows, a.dim, bench.SEED, dev, lo, hi), a.rows, rank, world)
q = bench.make_rows(a.queries, a.dim, bench.SEED + 100, dev)
times = collections.OrderedDict()
pending = []


def wrap(name):
    fn = getattr(index, name)

    def timed(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args, **kw)
        e1.record()
        pending.append((name, e0, e1))
        return out
    setattr(index, name, timed)


for n in ("_local_candidates", "_gather", "_global_threshold", "_rerank_owned", "_merge_certified"):
    wrap(n)
for it in range(3 + a.steps):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    index.search(q, a.k)
    e1.record()
    torch.cuda.synchronize()
    if it >= 3:
        times.setdefault("total", []).append(e0.elapsed_time(e1))
        seen = collections.Counter()
        for name, s, e in pending:
            seen[name] += 1
            times.setdefault("%s#%d" % (name, seen[name]), []).append(s.elapsed_time(e))
    pending.clear()
if rank == 0:
    for k_, v in times.items():
        print("%-24s %8.3f ms" % (k_, sum(v) / len(v)))
dist.barrier()
dist.destroy_process_group()
