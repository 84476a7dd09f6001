"""Persisted descriptor database.

The reference has no on-disk format: every run recomputes the embeddings of the
whole reference set (train/siamese_regions.py:26-41, utils/train_siamese.py:48-55)
before it can search it.  This module stores the ``[N, D]`` fp32 descriptors
``get_embeddings`` returns ONCE and lets every rank of a row-sharded search load
just its rows ``[lo, hi)`` (``search.shard_bounds``) straight into HBM.

File layout (little endian), one file per database:

    offset 0    8 bytes   magic  b"ISBDESC1"
           8    uint64    n_rows
          16    uint32    dim
          20    uint32    dtype code (0 = float32)
          24    uint32    flags (reserved, 0)
          28    36 bytes  zero padding (header = 64 bytes)
          64    n_rows * dim float32, row-major

The bf16 screen copy is NOT stored: ``isb_f32_to_bf16`` rebuilds it at HBM speed
(2 ms for 1M x 2048), cheaper than reading it from any disk.  Host logic only --
reading is ``numpy.memmap`` + pinned staging + ``cudaMemcpyAsync``; all arithmetic
stays in libisb.so.
"""

import os
import struct

import numpy as np
import torch

from ._lib import IsbError

MAGIC = b"ISBDESC1"
HEADER_BYTES = 64
_HEADER = struct.Struct("<8sQIII")      # magic, n_rows, dim, dtype, flags


def write_descriptors(path, emb, chunk_rows=65536):
    """Write ``emb`` ([N, D] float32, CPU or CUDA) as a descriptor file; device tensors
    are streamed to the host ``chunk_rows`` rows at a time.  Returns the byte size."""
    if emb.dim() != 2 or emb.dtype != torch.float32:
        raise IsbError("write_descriptors: expected a [N, D] float32 tensor")
    n, d = emb.shape
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_HEADER.pack(MAGIC, n, d, 0, 0).ljust(HEADER_BYTES, b"\0"))
        for s in range(0, n, chunk_rows):
            block = emb[s:s + chunk_rows].detach().contiguous().cpu()
            f.write(block.numpy().tobytes())
    os.replace(tmp, path)       # a reader never sees a half-written file
    return HEADER_BYTES + n * d * 4


class DescriptorFile(object):
    """Read access to a descriptor file (memory-mapped; nothing is read until asked for)."""

    def __init__(self, path):
        self.path = path
        size = os.path.getsize(path)
        if size < HEADER_BYTES:
            raise IsbError("%s: too short to be a descriptor file" % path)
        with open(path, "rb") as f:
            magic, n, d, dtype, flags = _HEADER.unpack(f.read(_HEADER.size))
        if magic != MAGIC:
            raise IsbError("%s: bad magic %r (expected %r)" % (path, magic, MAGIC))
        if dtype != 0:
            raise IsbError("%s: unsupported dtype code %d" % (path, dtype))
        if d == 0 or size != HEADER_BYTES + n * d * 4:
            raise IsbError("%s: header says %d x %d float32 but the file has %d bytes" % (path, n, d, size))
        self.n_rows, self.dim = int(n), int(d)
        self._map = None

    def _rows(self):
        if self._map is None:
            self._map = np.memmap(self.path, dtype="<f4", mode="r", offset=HEADER_BYTES,
                                  shape=(self.n_rows, self.dim))
        return self._map

    def read_rows(self, lo, hi):
        """Rows [lo, hi) as a CPU float32 tensor (a copy)."""
        if not (0 <= lo <= hi <= self.n_rows):
            raise IsbError("read_rows: [%d, %d) outside [0, %d)" % (lo, hi, self.n_rows))
        return torch.from_numpy(np.array(self._rows()[lo:hi], dtype=np.float32, copy=True))

    def load_rows(self, lo, hi, device, chunk_rows=131072):
        """Rows [lo, hi) on ``device``: read in chunks through two pinned staging buffers so the
        disk read of chunk i + 1 overlaps the host->device copy of chunk i."""
        if not (0 <= lo <= hi <= self.n_rows):
            raise IsbError("load_rows: [%d, %d) outside [0, %d)" % (lo, hi, self.n_rows))
        device = torch.device(device)
        out = torch.empty((hi - lo, self.dim), dtype=torch.float32, device=device)
        if device.type != "cuda":
            out.copy_(self.read_rows(lo, hi))
            return out
        stage = [torch.empty((min(chunk_rows, max(hi - lo, 1)), self.dim), dtype=torch.float32).pin_memory()
                 for _ in range(2)]
        events = [None, None]
        rows = self._rows()
        # copies, events and the final wait all on the TARGET device's current stream (the caller's
        # current device may be another one)
        with torch.cuda.device(device):
            for i, s in enumerate(range(lo, hi, chunk_rows)):
                e = min(s + chunk_rows, hi)
                buf = stage[i & 1]
                if events[i & 1] is not None:
                    events[i & 1].synchronize()          # the previous copy out of this buffer is done
                np.copyto(buf[:e - s].numpy(), rows[s:e])
                out[s - lo:e - lo].copy_(buf[:e - s], non_blocking=True)
                events[i & 1] = torch.cuda.Event()
                events[i & 1].record()
            torch.cuda.current_stream().synchronize()
        return out


def load_index(path, rank=0, world_size=1, device=None, group=None):
    """Open a descriptor file as a searchable index: ``DescriptorIndex`` of the whole file
    (world_size 1) or the ``ShardedIndex`` holding this rank's rows [lo, hi)."""
    from .search import DescriptorIndex, ShardedIndex, shard_bounds
    f = DescriptorFile(path)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    lo, hi = shard_bounds(f.n_rows, world_size)[rank]
    rows = f.load_rows(lo, hi, device)
    if world_size == 1:
        return DescriptorIndex(rows)
    return ShardedIndex(rows, f.n_rows, rank, world_size, group)
