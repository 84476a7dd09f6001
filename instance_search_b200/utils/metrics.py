"""Drop-in for the reference's ``utils/metrics.py`` (P@1, Oxford AP, mAP from a
similarity matrix) with the row reductions on the GPU.

The reference takes ``sim.max(1)`` / ``kthvalue`` and, for AP, a full descending
``sort`` of every row followed by a Python walk over ALL N ranks
(utils/metrics.py:33-43).  Only the ranks of the query's positives change the AP
sum (every other step adds exactly 0.0: recall is unchanged), so the GPU computes
just those rank positions (``isb_row_ranks``: one streaming pass per row, no
sort) and the trapezoid sum is then accumulated on the host in the reference's
own order and float64 arithmetic -- the result is bit-identical to the
reference's for the same ``sim``.

``sim`` must be a CUDA float32 tensor [n_test, n_ref]; data sets are the
reference's lists of ``(tensor, label, name)`` triples.
"""

import torch

from .. import ops


def _label_table(test_set, ref_set):
    ids = {}
    ref = [ids.setdefault(lab, len(ids)) for _, lab, _ in ref_set]
    test = [ids.get(lab, -1) for _, lab, _ in test_set]
    return test, ref


def precision1(sim, test_set, ref_set, kth=1):
    """reference: utils/metrics.py:8-19.  Returns (precision, correct, total,
    max_sim [n_test, 1], max_label list) like the reference (torch-0.1 ``max(1)``
    kept the reduced dimension)."""
    total = sim.size(0)
    val, idx = ops.row_kth_largest(sim, max(1, kth))
    idx_host = idx.tolist()
    max_label = [ref_set[i][1] for i in idx_host]
    correct = sum(test_label == max_label[j] for j, (_, test_label, _) in enumerate(test_set))
    return float(correct) / total, correct, total, val.view(-1, 1), max_label


def _positive_ranks(sim, rows, test_ids, ref_ids):
    """For each listed query row: sorted rank positions of its positives."""
    by_label = {}
    for j, l in enumerate(ref_ids):
        by_label.setdefault(l, []).append(j)
    lists = [by_label.get(test_ids[i], []) for i in rows]
    P = max(1, max(len(l) for l in lists)) if lists else 1
    cols = torch.full((len(rows), P), -1, dtype=torch.int32)
    for r, l in enumerate(lists):
        if l:
            cols[r, :len(l)] = torch.tensor(l, dtype=torch.int32)
    sub = sim if len(rows) == sim.size(0) else sim[torch.tensor(rows, device=sim.device)]
    rank = ops.row_ranks(sub.contiguous(), cols.to(sim.device)).cpu()
    return [sorted(rank[r, :len(l)].tolist()) for r, l in enumerate(lists)]


def _ap_from_ranks(ranks, n_pos, kth):
    """The loop of utils/metrics.py:31-44 restricted to the steps that change ``ap``."""
    n_pos -= (kth - 1)
    if n_pos <= 0:
        return None
    old_recall, ap = 0.0, 0.0
    intersect_size = 0
    for n in ranks:                      # n: 0-based position in the descending ranking
        if n + 1 < kth:
            continue
        j = n - (kth - 1)                # steps taken before this one
        old_precision = 1.0 if j == 0 else intersect_size / (j - 1 + 1.0)
        intersect_size += 1
        recall = intersect_size / float(n_pos)
        precision = intersect_size / (j + 1.0)
        ap += (recall - old_recall) * ((old_precision + precision) / 2.0)
        old_recall = recall
    return ap


def avg_precision(sim, i, test_set, ref_set, kth=1):
    """reference: utils/metrics.py:25-45"""
    test_ids, ref_ids = _label_table(test_set, ref_set)
    ranks = _positive_ranks(sim, [i], test_ids, ref_ids)[0]
    return _ap_from_ranks(ranks, len(ranks), kth)


def mean_avg_precision(sim, test_set, ref_set, kth=1):
    """reference: utils/metrics.py:48-55"""
    test_ids, ref_ids = _label_table(test_set, ref_set)
    rows = list(range(sim.size(0)))
    all_ranks = _positive_ranks(sim, rows, test_ids, ref_ids)
    aps = []
    for ranks in all_ranks:
        ap = _ap_from_ranks(ranks, len(ranks), kth)
        if ap is not None:
            aps.append(ap)
    return sum(aps) / float(len(aps))
