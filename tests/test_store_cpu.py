"""Descriptor file format (instance_search_b200/store.py): round trips, shard reads and
header validation on the host; the GPU loader is covered by tests/test_gpu_sharded.py."""

import struct

import pytest
import torch

from instance_search_b200 import store
from instance_search_b200._lib import IsbError
from instance_search_b200.search import shard_bounds


def _emb(n, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return x / x.norm(dim=1, keepdim=True)


@pytest.mark.parametrize("n,d,chunk", [(1000, 64, 65536), (777, 33, 100), (1, 8, 1), (0, 16, 10)])
def test_round_trip_and_shards(tmp_path, n, d, chunk):
    emb = _emb(n, d) if n else torch.zeros(0, d)
    path = str(tmp_path / "db.isbd")
    size = store.write_descriptors(path, emb, chunk_rows=chunk)
    assert size == store.HEADER_BYTES + n * d * 4
    f = store.DescriptorFile(path)
    assert (f.n_rows, f.dim) == (n, d)
    assert torch.equal(f.read_rows(0, n), emb)                     # bit-exact
    for world in (1, 2, 3, 8):
        parts = [f.read_rows(lo, hi) for lo, hi in shard_bounds(n, world)]
        assert torch.equal(torch.cat(parts), emb)
    assert torch.equal(f.load_rows(0, n, "cpu"), emb)


def test_header_is_little_endian_and_documented(tmp_path):
    path = str(tmp_path / "db.isbd")
    store.write_descriptors(path, _emb(5, 4))
    raw = open(path, "rb").read()
    assert raw[:8] == b"ISBDESC1"
    assert struct.unpack("<QIII", raw[8:28]) == (5, 4, 0, 0)
    assert raw[28:64] == b"\0" * 36 and len(raw) == 64 + 5 * 4 * 4


def test_rejects_bad_files(tmp_path):
    path = str(tmp_path / "db.isbd")
    store.write_descriptors(path, _emb(10, 8))
    raw = bytearray(open(path, "rb").read())
    bad_magic = str(tmp_path / "a")
    open(bad_magic, "wb").write(b"NOTADESC" + bytes(raw[8:]))
    with pytest.raises(IsbError, match="bad magic"):
        store.DescriptorFile(bad_magic)
    truncated = str(tmp_path / "b")
    open(truncated, "wb").write(bytes(raw[:-4]))
    with pytest.raises(IsbError, match="header says"):
        store.DescriptorFile(truncated)
    short = str(tmp_path / "c")
    open(short, "wb").write(b"ISBD")
    with pytest.raises(IsbError, match="too short"):
        store.DescriptorFile(short)
    with pytest.raises(IsbError):
        store.DescriptorFile(path).read_rows(5, 11)
    with pytest.raises(IsbError):
        store.write_descriptors(path, _emb(4, 4).double())
