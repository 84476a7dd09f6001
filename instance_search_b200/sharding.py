"""Host-side plumbing shared by the multi-GPU paths: contiguous shard bounds
and the one collective the independent-unit paths need -- an all-gather of per-rank row
blocks of unequal length.  One process per GPU over ``torch.distributed`` (NCCL on the box,
gloo in the CPU tests); nothing here touches a kernel."""

import torch


def shard_bounds(n_rows, world_size):
    """Row range [lo, hi) of every rank: as even as possible, contiguous."""
    base, extra = divmod(n_rows, world_size)
    bounds, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def all_gather_rows(local, n_total, rank, world_size, group=None):
    """local: rows [lo, hi) of an [n_total, ...] tensor (this rank's shard_bounds slice).
    Returns the full [n_total, ...] tensor on every rank (ONE all_gather_into_tensor of
    blocks padded to the largest shard; the padding rows are dropped)."""
    if world_size == 1:
        return local
    import torch.distributed as dist
    bounds = shard_bounds(n_total, world_size)
    lo, hi = bounds[rank]
    if local.size(0) != hi - lo:
        raise ValueError("rank %d: %d rows, expected %d" % (rank, local.size(0), hi - lo))
    per = bounds[0][1] - bounds[0][0]                   # the largest shard
    tail = tuple(local.shape[1:])
    block = local.new_zeros((per,) + tail)
    block[:hi - lo] = local
    full = local.new_empty((world_size * per,) + tail)
    dist.all_gather_into_tensor(full.view(world_size * per, -1), block.view(per, -1), group=group)
    if per * world_size == n_total:
        return full
    keep = torch.cat([torch.arange(r * per, r * per + (b[1] - b[0]), device=local.device)
                      for r, b in enumerate(bounds)])
    return full.index_select(0, keep)
