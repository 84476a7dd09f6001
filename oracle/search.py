"""Oracle (test infrastructure): cosine scoring and top-k of the reference.

The reference never calls ``topk`` on similarities; it materialises
``sim = torch.mm(Q, DB.t())`` (test/siamese_regions_test.py:76,
utils/train_siamese.py:70) and consumes it with ``max`` / ``kthvalue`` / a full
descending ``sort`` (utils/metrics.py:11,13,33).  "top-k" here is therefore
defined as the first k entries of that descending sort.
"""

import torch


def similarity(q, db):
    """reference: test/siamese_regions_test.py:76  sim = torch.mm(test, ref.t())"""
    return torch.mm(q, db.t())


def topk_search(q, db, k, chunk=1024):
    """Top-k of every row of mm(q, db.t()), descending.

    reference: test/siamese_regions_test.py:76 + utils/metrics.py:33
    (``sim[i].sort(dim=0, descending=True)``; the first k of that ranking).
    Queries are processed in chunks so the Q x N fp32 matrix never has to exist
    at once -- chunking rows of a matmul does not change any value.
    Returns (scores [Q, k] fp32, idx [Q, k] int64).
    """
    k = min(k, db.size(0))
    scores = torch.empty(q.size(0), k, dtype=torch.float32)
    idx = torch.empty(q.size(0), k, dtype=torch.int64)
    for s in range(0, q.size(0), chunk):
        sim = torch.mm(q[s:s + chunk], db.t())
        v, i = sim.sort(dim=1, descending=True)
        scores[s:s + chunk] = v[:, :k]
        idx[s:s + chunk] = i[:, :k]
    return scores, idx


def topk_search_f64(q, db, k, chunk=256):
    """Same ranking computed with fp64 dot products (the adjudicator).

    Two fp32 summation orders legitimately disagree on the order of two
    database rows whose scores differ by less than fp32 accumulation noise
    ("index-exact vs fp32 noise").  Parity tests use this
    to decide which of two disagreeing fp32 rankings is the true one.
    Returns (scores fp64 [Q, k], idx int64 [Q, k]); ties broken by lower index.
    """
    k = min(k, db.size(0))
    db64 = db.double()
    scores = torch.empty(q.size(0), k, dtype=torch.float64)
    idx = torch.empty(q.size(0), k, dtype=torch.int64)
    for s in range(0, q.size(0), chunk):
        sim = torch.mm(q[s:s + chunk].double(), db64.t())
        v, i = sim.sort(dim=1, descending=True, stable=True)
        scores[s:s + chunk] = v[:, :k]
        idx[s:s + chunk] = i[:, :k]
    return scores, idx
