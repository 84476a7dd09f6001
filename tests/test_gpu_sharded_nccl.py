"""Real multi-process runs over NCCL (need >= 2 GPUs; skipped otherwise):
ShardedIndex.search -- candidate exchange and the replicated-re-rank variant must both equal the
single-GPU search and the oracle, with host (pinned) and device queries, with and without
deferred exactness tickets; mining.ShardedMiner -- couples split across ranks == one GPU ==
oracle; data-parallel get_embeddings -- images split across ranks == one GPU."""

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
        # deferred tickets: two searches queued back to back, resolved afterwards (collectively)
        s1, i1, t1 = index.search(q.to(dev), k, defer=True)
        s2, i2, t2 = index.search(q.to(dev), k, defer=True)
        assert t1.resolve() == 1 and t2.resolve() == 1          # row 7 both times
        torch.cuda.synchronize()
        assert torch.equal(i1.cpu(), one_i.cpu()) and torch.equal(i2.cpu(), one_i.cpu())

        # ---- anchor-sharded mining
        from instance_search_b200 import mining
        Nm, Dm, per = 4096, 128, 8
        lab = torch.arange(Nm) // per
        E = oracle.normalize_l2(torch.randn(Nm // per, Dm, generator=g)[lab] + 0.5 * torch.randn(Nm, Dm, generator=g))
        anchors = torch.randperm(Nm, generator=g)[:333]
        positives = (anchors // per) * per + (anchors % per + 1) % per
        one = mining.MiningIndex(E.to(dev), lab).select_negatives(anchors, positives, True)
        sharded = mining.ShardedMiner(E.to(dev), lab, rank, world).select_negatives(anchors, positives, True)
        for a_, b_ in zip(one, sharded):
            assert torch.equal(a_.cpu(), b_.cpu()), "sharded mining differs from the single-GPU result"
        want = oracle.select_negatives(E @ E.t(), lab, list(zip(anchors.tolist(), positives.tolist())), True)
        assert int((sharded[0].cpu() != want).sum()) <= 1

        # ---- data-parallel embeddings
        from test_host_cpu import ToyNet
        from instance_search_b200.model.siamese import RegionDescriptorNet
        from instance_search_b200.train.siamese_regions import get_embeddings
        torch.manual_seed(5)
        net = RegionDescriptorNet(ToyNet(8, 5), 6, 16, (7, 7))
        net.feature_reduc1[1].param.data.normal_(0, 0.01)
        net = net.to(dev).eval()
        ds = [(torch.randn(3, 32, 32, generator=g), "L%d" % (n % 4), "im%d" % n) for n in range(21)]
        full = get_embeddings(net, ds, 0, 16, batch_size=4, rank=rank, world_size=world)
        ref = get_embeddings(net, ds, 0, 16, batch_size=4)
        assert full.shape == (21, 16) and torch.allclose(full, ref, atol=1e-6)
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
