"""CPU model of the plain GEMM's work decomposition with chunked accumulation (csrc/isb_gemm.cu
PlainSched + gemm_nt_impl, restated in Python): over all segments and passes every
(row block, n-tile, k-block) triple is accumulated exactly once; no pass is longer than the
chunk limit; the first pass of every (output tile, split) is the one that stores (the later
ones add); split operands keep kb = 3 * kk + term whatever the chunk boundaries; the CTA-pair
variant covers the same triples with 256-row blocks -- no GPU needed."""

import pytest

BM, BN, BK, CHUNK = 128, 256, 64, 64      # kBM, kBN, kBK, kAccChunkKb


def host_plan(M, N, K, splits, terms, pair):
    m_blocks, n_tiles = (M + BM - 1) // BM, (N + BN - 1) // BN
    k_blocks = terms * ((K + BK - 1) // BK)
    splits = max(1, min(splits, k_blocks))
    kb_split = (k_blocks + splits - 1) // splits
    n_chunks = (kb_split + CHUNK - 1) // CHUNK
    chunk_kb = (kb_split + n_chunks - 1) // n_chunks
    use_pair = pair and m_blocks >= 2
    sched_m = (m_blocks + 1) // 2 if use_pair else m_blocks
    return dict(m_blocks=m_blocks, sched_m=sched_m, n_tiles=n_tiles, k_blocks=k_blocks, splits=splits,
                chunk_kb=chunk_kb, pair=use_pair)


def segment(p, s):
    t, m_block = divmod(s, p["sched_m"])
    nt, sp = divmod(t, p["splits"])
    kb_begin = sp * p["k_blocks"] // p["splits"]
    kb_end = (sp + 1) * p["k_blocks"] // p["splits"]
    passes = (kb_end - kb_begin + p["chunk_kb"] - 1) // p["chunk_kb"]
    return dict(m_block=m_block, n_tile=nt, split=sp, kb_begin=kb_begin, kb_end=kb_end, passes=passes)


def kb_range(p, seg, c):
    kb0 = seg["kb_begin"] + c * p["chunk_kb"]
    return kb0, min(kb0 + p["chunk_kb"], seg["kb_end"])


@pytest.mark.parametrize("M,N,K,splits,terms", [
    (256, 2048, 100352, 9, 3),      # the whitening projection
    (32, 2048, 100352, 18, 3),
    (16384, 16384, 2048, 1, 3),     # all-pairs similarities
    (4096, 464, 2048, 2, 3),        # candidates' re-score
    (130, 258, 6400, 7, 1), (1, 1, 8, 1, 1), (200, 300, 136, 1, 3), (2048, 100352, 256, 1, 3),
])
@pytest.mark.parametrize("pair", [False, True])
def test_every_k_block_of_every_tile_is_accumulated_exactly_once(M, N, K, splits, terms, pair):
    p = host_plan(M, N, K, splits, terms, pair)
    n_seg = p["sched_m"] * p["n_tiles"] * p["splits"]
    rows_per_block = 2 if p["pair"] else 1
    seen = {}
    first_store = set()
    for s in range(n_seg):
        seg = segment(p, s)
        assert seg["passes"] >= 1 and seg["kb_end"] > seg["kb_begin"]
        covered = []
        for c in range(seg["passes"]):
            kb0, kb1 = kb_range(p, seg, c)
            assert 0 < kb1 - kb0 <= CHUNK                      # chain length bounded
            covered.extend(range(kb0, kb1))
            for r in range(rows_per_block):
                mb = rows_per_block * seg["m_block"] + r
                if mb >= p["m_blocks"]:
                    continue                                   # the pair's second CTA has no rows (TMA zero fill)
                key = (mb, seg["n_tile"], seg["split"])
                if c == 0:
                    assert key not in first_store              # exactly one storing pass per (tile, split)
                    first_store.add(key)
                for kb in range(kb0, kb1):
                    seen[(mb, seg["n_tile"], kb)] = seen.get((mb, seg["n_tile"], kb), 0) + 1
        assert covered == list(range(seg["kb_begin"], seg["kb_end"]))   # passes tile the split's range in order
    assert len(seen) == p["m_blocks"] * p["n_tiles"] * p["k_blocks"] and set(seen.values()) == {1}
    assert len(first_store) == p["m_blocks"] * p["n_tiles"] * p["splits"]
    if terms == 3:
        # kb = 3 * kk + term: all three terms of every k-range kk are covered
        kks = (K + BK - 1) // BK
        assert {kb // 3 for (_, _, kb) in seen} == set(range(kks)) and {kb % 3 for (_, _, kb) in seen} == {0, 1, 2}
