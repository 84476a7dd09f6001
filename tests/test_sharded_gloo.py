"""CPU, world_size 2 and 3 over gloo: the host-side plumbing of the row-sharded
search (shard bounds, global index offsets, padding of short shards, the
candidate exchange -- all-gather of screen scores, global threshold, re-rank of
the owned candidates, all-gather of the per-shard lists, certified merge, the
resolve path of uncertified rows -- and the replicated-re-rank variant) --
SURVEY.md section 8e.  The GPU kernels are replaced through ShardedIndex's hooks
by torch restatements of their contracts (include/isb.h) on top of the oracle,
so what is exercised is exactly the code bench.py runs at N > 1 minus the
kernels, which tests/test_gpu_core.py / test_gpu_sharded.py cover on the device."""

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


def _worker(rank, world, port, n_total, k, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from instance_search_b200.search import ShardedIndex, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(7)
    db = oracle.normalize_l2(torch.randn(n_total, 32, generator=g))
    q = oracle.normalize_l2(torch.randn(17, 32, generator=g))

    class HostShard(object):
        def __init__(self, rows, off):
            self.rows, self.off = rows, off

    class CpuSharded(ShardedIndex):
        def _make_local(self, local_db, row_offset):
            return HostShard(local_db, row_offset)

        def _local_search(self, q, k, events=None):
            s, i = oracle.topk_search(q, self.local.rows, k)
            return s, i + self.local.off

        def _merge(self, cs, ci):
            # same contract as isb_topk_merge: best first, ties -> lower index, idx < 0 last
            R, Q, kk = cs.shape
            s = cs.permute(1, 0, 2).reshape(Q, R * kk).clone()
            i = ci.permute(1, 0, 2).reshape(Q, R * kk)
            s[i < 0] = float("-inf")
            key = i.clone()
            key[i < 0] = torch.iinfo(torch.int64).max
            o1 = key.argsort(dim=1, stable=True)
            s1, i1 = s.gather(1, o1), i.gather(1, o1)
            o2 = s1.argsort(dim=1, descending=True, stable=True)
            return s1.gather(1, o2)[:, :kk], i1.gather(1, o2)[:, :kk]

        # ---- candidate exchange hooks (contracts of isb_topk_candidates / _global_threshold /
        # _rerank_owned / _merge_certified)
        def _local_candidates(self, q, k, kc, events=None):
            sim = oracle.similarity(q, self.local.rows)
            screen = sim.bfloat16().float()          # a noisy screen, like the bf16 GEMM
            n = min(kc, sim.size(1))
            v, c = screen.topk(n, dim=1)
            cs = torch.full((q.size(0), kc), float("-inf"))
            cc = torch.full((q.size(0), kc), -1, dtype=torch.int32)
            cs[:, :n], cc[:, :n] = v, c.int()
            return cs, cc

        def _global_threshold(self, all_screen, kth_rank):
            R, Q, kc = all_screen.shape
            assert kth_rank >= kc
            flat = all_screen.permute(1, 0, 2).reshape(Q, R * kc)
            srt = flat.sort(dim=1, descending=True).values
            kth = srt[:, kth_rank - 1] if kth_rank <= R * kc else torch.full((Q,), float("-inf"))
            enough = (flat > float("-inf")).sum(1) >= kth_rank
            # reduced lists: too few entries overall but some shard's list is full -> +inf (nothing
            # can be concluded; the merge cannot certify)
            full = (all_screen[:, :, kc - 1] > float("-inf")).any(0) & (kth_rank > kc)
            short = torch.where(full, torch.full_like(kth, float("inf")), torch.full_like(kth, float("-inf")))
            return torch.where(enough, kth, short)

        def _rerank_owned(self, q, k, cand_screen, cand_col, thr):
            sim = oracle.similarity(q, self.local.rows)
            Q, kc = cand_screen.shape
            own = (cand_col >= 0) & (cand_screen >= thr[:, None])
            exact = sim.gather(1, cand_col.clamp(min=0).long())
            exact = torch.where(own, exact, torch.full_like(exact, float("-inf")))
            gidx = torch.where(own, cand_col.long() + self.local.off, torch.full_like(cand_col.long(), -1))
            key = torch.where(own, gidx, torch.full_like(gidx, torch.iinfo(torch.int64).max))
            o1 = key.argsort(dim=1, stable=True)
            e1, g1 = exact.gather(1, o1), gidx.gather(1, o1)
            o2 = e1.argsort(dim=1, descending=True, stable=True)
            e2, g2 = e1.gather(1, o2), g1.gather(1, o2)
            s = torch.full((Q, k), float("-inf"))
            i = torch.full((Q, k), -1, dtype=torch.int64)
            n = min(k, kc)
            s[:, :n], i[:, :n] = e2[:, :n], g2[:, :n]
            d = torch.where(own, cand_screen - sim.gather(1, cand_col.clamp(min=0).long()), torch.zeros_like(exact))
            stat = torch.stack([(d * d).sum(1), own.sum(1).float()], 1)
            # a full list that lies entirely above the threshold may have been cut above it
            cut = ((cand_col >= 0) & (cand_screen > thr[:, None])).sum(1) == kc
            stat[cut, 0] = float("inf")
            # packed row (include/isb.h): k scores | k LOCAL rows | 2 stat words, 32 bits each
            col = torch.where(i >= 0, i - self.local.off, i).int()
            return torch.cat([s.view(torch.int32), col, stat.view(torch.int32)], 1)

        def _merge_certified(self, packed_all, thr, k):
            R, Q, pw = packed_all.shape
            assert pw == 2 * k + 2 and thr.shape == (Q,) and packed_all.dtype == torch.int32
            cs = packed_all[:, :, :k].contiguous().view(torch.float32)
            col = packed_all[:, :, k:2 * k].long()
            offs = torch.tensor([lo for lo, _ in shard_bounds(self.n_total, self.world_size)])
            ci = torch.where(col >= 0, col + offs[:, None, None], col)
            stat = packed_all[:, :, 2 * k:].contiguous().view(torch.float32)
            # every candidate >= thr was re-ranked by exactly one shard
            total = stat[:, :, 1].sum(0)
            kc = min(k + 28, 128)
            assert bool(((total >= min(kc, n_total)) | (thr == float("-inf")) | (thr == float("inf"))).all())
            s, i = self._merge(cs, ci)
            # pretend the certificate rejects every 5th row (listed in a rank-dependent order,
            # like the device's atomicAdd): the resolve path must reproduce them exactly
            rows = torch.arange(3, cs.size(1), 5, dtype=torch.int32)
            if self.rank % 2:
                rows = rows.flip(0)
            unc = torch.zeros(cs.size(1), dtype=torch.int32)
            unc[:rows.numel()] = rows
            s[rows.long()] = 123.0           # garbage the resolve path has to overwrite
            i[rows.long()] = -7
            return s, i, unc, torch.tensor([rows.numel()], dtype=torch.int32)

    lo, hi = shard_bounds(n_total, world)[rank]
    index = CpuSharded(db[lo:hi], n_total, rank, world)
    want_s, want_i = oracle.topk_search(q, db, k)
    ok = True
    for exchange in (True, False):
        s, i = index.search(q, k, exchange=exchange)
        ok = ok and torch.equal(i, want_i) and torch.equal(s, want_s)
    ok = ok and index.stats["resolved_locally_exact"] == len(range(3, 17, 5))
    # every rank holds the full merged answer
    flags = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1.0 if ok else 0.0]))
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok and all(float(x) == 1.0 for x in flags) else
                "mismatch %s vs %s" % (i[0].tolist(), want_i[0].tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,k", [
    (2, 1001, 10),     # uneven split
    (2, 12, 10),       # shards (6 rows) smaller than k: padded with invalid entries
    (3, 500, 25),
])
def test_sharded_search_plumbing(tmp_path, world, n_total, k):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, k, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("rank%d" % r)).read_text() == "ok"


def test_shard_bounds():
    from instance_search_b200.search import shard_bounds
    for n, w in [(10, 1), (10, 3), (1000000, 8), (5, 8), (0, 2)]:
        b = shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
