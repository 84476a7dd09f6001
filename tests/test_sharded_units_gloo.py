"""CPU, world_size 2 and 3 over gloo: the plumbing of the two independent-unit multi-GPU paths
 -- data-parallel embeddings (images split across ranks, one
all-gather of the descriptor blocks) and anchor-sharded negative mining (couples split across
ranks, one all-gather of the packed results).  The kernels are replaced through the classes'
hooks by the oracle, so what runs is the host code of mining.ShardedMiner /
sharding.all_gather_rows exactly as on the GPUs; the kernels are covered by the GPU tests."""

import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, n_couples, out_dir):
    sys.path.insert(0, ROOT)
    import oracle
    from instance_search_b200 import mining
    from instance_search_b200.sharding import all_gather_rows, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(11)
    lab = torch.arange(n_items) // 4
    E = oracle.normalize_l2(torch.randn(n_items // 4 + 1, 24, generator=g)[lab] + 0.5 * torch.randn(n_items, 24, generator=g))
    S = oracle.mining.all_pairs_similarities(E)
    anchors = torch.randint(0, n_items, (n_couples,), generator=g)
    positives = (anchors // 4) * 4 + (anchors % 4 + 1) % 4
    positives = positives.clamp(max=n_items - 1)

    class CpuMiner(mining.ShardedMiner):
        def _make_index(self, emb, labels, terms):
            return None

        def _local_select(self, a, b, semi_hard):
            neg = oracle.select_negatives(S, lab, list(zip(a.tolist(), b.tolist())), semi_hard)
            ok = neg >= 0
            nsim = torch.full((len(a),), -2.0)
            nsim[ok] = S[a[ok], neg[ok]]
            return neg, nsim, S[a, b]

    miner = CpuMiner(E, lab, rank, world)
    res = {}
    for semi in (False, True):
        res[semi] = miner.select_negatives(anchors, positives, semi)

    # data-parallel embeddings: every rank "embeds" its contiguous slice, one all-gather
    lo, hi = shard_bounds(n_items, world)[rank]
    full = all_gather_rows(E[lo:hi].clone(), n_items, rank, world)
    ints = all_gather_rows(torch.arange(lo, hi), n_items, rank, world)      # 1-D rows too
    torch.save({"res": res, "full": full, "ints": ints, "E": E, "S": S, "lab": lab, "anchors": anchors,
                "positives": positives}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items,n_couples", [(2, 64, 37), (3, 50, 10), (3, 40, 2)])
def test_sharded_miner_and_embedding_gather_over_gloo(tmp_path, world, n_items, n_couples):
    import oracle
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_items, n_couples, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    o = outs[0]
    couples = list(zip(o["anchors"].tolist(), o["positives"].tolist()))
    for semi in (False, True):
        want = oracle.select_negatives(o["S"], o["lab"], couples, semi)
        for r in range(world):
            neg, nsim, psim = outs[r]["res"][semi]
            assert torch.equal(neg, want)                                    # identical on every rank
            assert torch.equal(psim, o["S"][o["anchors"], o["positives"]])
            ok = want >= 0
            assert torch.equal(nsim[ok], o["S"][o["anchors"][ok], want[ok]]) and (nsim[~ok] == -2).all()
    for r in range(world):
        assert torch.equal(outs[r]["full"], o["E"])
        assert torch.equal(outs[r]["ints"], torch.arange(n_items))
