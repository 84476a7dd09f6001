"""BASELINE.json configs[4]: end-to-end instance search on N GPUs of one box --
ResNet-152 trunk (PyTorch, random init, the reference's 448-px input -> 14 x 14 maps) ->
region descriptors (fused CUDA head, D = 512, k = 6) -> top-100 over a synthetic
10M x 512-d database sharded row-wise (candidate exchange over NCCL).

    python -m torch.distributed.run --nproc-per-node N tools/e2e_instance_search.py [--db-rows 10000000]

Images are data-parallel (no exchange); the query descriptors are all-gathered ([Q, 512])
and searched collectively.  Prints ONE JSON line on rank 0.  Side measurement: the
contract bench is bench.py."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from instance_search_b200 import regions  # noqa: E402
from instance_search_b200.model.siamese import RegionDescriptorNet  # noqa: E402
from instance_search_b200.search import ShardedIndex, shard_bounds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--db-rows", type=int, default=10000000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--images-per-gpu", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--px", type=int, default=448)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True

torch.manual_seed(0)
trunk = torchvision.models.resnet152(weights=None, num_classes=464)
net = RegionDescriptorNet(trunk, 6, a.dim, (7, 7)).to(dev).eval()
lo, hi = shard_bounds(a.db_rows, world)[rank]
index = ShardedIndex(bench.make_rows_slice(a.db_rows, a.dim, 1234 + 5, dev, lo, hi), a.db_rows, rank, world)

g = torch.Generator().manual_seed(100 + rank)
mean = torch.tensor([0.36, 0.30, 0.28]).view(1, 3, 1, 1)
std = torch.tensor([0.21, 0.20, 0.20]).view(1, 3, 1, 1)
images = ((torch.rand(a.images_per_gpu, 3, a.px, a.px, generator=g) - mean) / std).pin_memory()


def ev():
    return torch.cuda.Event(enable_timing=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


t_trunk, t_head, t_search, t_total = [], [], [], []
for it in range(1 + a.steps):
    barrier()
    e = [ev() for _ in range(4)]
    trunk_ms = head_ms = 0.0
    e[0].record()
    descs = []
    marks = []
    with torch.no_grad():
        for s in range(0, a.images_per_gpu, a.batch):
            x = images[s:s + a.batch].to(dev, non_blocking=True)
            m0, m1, m2 = ev(), ev(), ev()
            m0.record()
            fmap = net.features(x)                      # PyTorch trunk (north_star: stays in PyTorch)
            m1.record()
            # the fused CUDA head, exactly what RegionDescriptorNet.forward runs in eval mode
            descs.append(regions.region_descriptors(fmap, net._head(), net.k, net.feature_size2d,
                                                    want_cls_out=False)[0])
            m2.record()
            marks.append((m0, m1, m2))
    d_local = torch.cat(descs)
    e[1].record()
    if world > 1:
        q = torch.empty((world * d_local.size(0), a.dim), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(q, d_local)
    else:
        q = d_local
    e[2].record()
    scores, idx = index.search(q, a.k)
    e[3].record()
    barrier()
    if it >= 1:
        t_trunk.append(sum(m0.elapsed_time(m1) for m0, m1, _ in marks))
        t_head.append(sum(m1.elapsed_time(m2) for _, m1, m2 in marks))
        t_search.append(e[2].elapsed_time(e[3]))
        t_total.append(e[0].elapsed_time(e[3]))

med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
ms = torch.tensor([med(t_trunk), med(t_head), med(t_search), med(t_total)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    tr, hd, se, tot = [float(v) for v in ms.tolist()]
    n_img = a.images_per_gpu * world
    print(json.dumps({
        "workload": "end-to-end instance search (BASELINE configs[4]): ResNet-152 trunk (PyTorch fp32, %d px) -> region "
                    "descriptors (D=%d, k=6) -> top-%d over %d x %d-d database, %d GPU(s)" % (a.px, a.dim, a.k, a.db_rows, a.dim, world),
        "n_gpus": world, "images": n_img, "ms": {"trunk": tr, "region_head": hd, "search": se, "total": tot},
        "images_per_s_end_to_end": n_img / (tot * 1e-3),
        "region_head_images_per_s": n_img / (hd * 1e-3),
        "search_queries_per_s": n_img / (se * 1e-3),
        "head_share_of_descriptor_stage": hd / (tr + hd),
        "unit_norm_ok": bool(torch.allclose(q.norm(dim=1), torch.ones(q.size(0), device=dev), atol=1e-4)),
        "top1_score_mean": float(scores[:, 0].mean())}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
