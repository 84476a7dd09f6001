"""Real multi-process run of ShardedIndex.search over NCCL (needs >= 2 GPUs; skipped
otherwise): candidate exchange and the replicated-re-rank variant must both equal the
single-GPU search and the oracle, with host (pinned) and device queries."""

import os
import socket
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    from parity import check_topk_against_oracle
    from instance_search_b200.search import DescriptorIndex, ShardedIndex, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    msg = "ok"
    try:
        g = torch.Generator().manual_seed(99)
        N, D, Q, k = 40003, 256, 75, 100
        db = oracle.normalize_l2(torch.randn(N, D, generator=g))
        q = oracle.normalize_l2(torch.randn(Q, D, generator=g))
        db[5000:5300] = q[7]                 # duplicates: the global certificate must reject row 7
        lo, hi = shard_bounds(N, world)[rank]
        index = ShardedIndex(db[lo:hi].to(dev), N, rank, world)
        one_s, one_i = DescriptorIndex(db.to(dev)).search(q.to(dev), k)
        for exchange in (True, False):
            for qq in (q.to(dev), q.pin_memory()):
                s, i = index.search(qq, k, exchange=exchange)
                torch.cuda.synchronize()
                assert torch.equal(i.cpu(), one_i.cpu()), "indices differ from the single-GPU search"
                assert torch.equal(s.cpu(), one_s.cpu()), "scores differ from the single-GPU search"
        assert index.stats["resolved_locally_exact"] >= 2      # row 7, once per exchange call
        keep = [r for r in range(Q) if r != 7]
        check_topk_against_oracle(q[keep], db, k, s[keep], i[keep])
    except Exception as e:  # noqa: BLE001
        msg = "rank %d: %r" % (rank, e)
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write(msg)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_search_nccl(tmp_path):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("rank%d" % r)).read_text() == "ok"
