"""Oracle (test infrastructure): P@1 / AP / mAP from a similarity matrix.

Follows /root/reference/utils/metrics.py with torch-0.1 reduction semantics
(``max(1)`` / ``kthvalue(k, 1)`` kept the reduced dimension, hence
``max_idx[i, 0]`` at :17).  Data sets are lists of ``(tensor, label, name)``
triples as built at test/siamese_regions_test.py:60-69.
"""


def precision1(sim, test_set, ref_set, kth=1):
    """reference: utils/metrics.py:8-19"""
    total = sim.size(0)
    if kth <= 1:
        max_sim, max_idx = sim.max(1, keepdim=True)
    else:
        max_sim, max_idx = sim.kthvalue(sim.size(1) - kth + 1, 1, keepdim=True)
    max_label = []
    for i in range(sim.size(0)):
        max_label.append(ref_set[int(max_idx[i, 0])][1])
    correct = sum(test_label == max_label[j]
                  for j, (_, test_label, _) in enumerate(test_set))
    return float(correct) / total, correct, total, max_sim, max_label


def avg_precision(sim, i, test_set, ref_set, kth=1):
    """Oxford-buildings AP of query i. reference: utils/metrics.py:25-45"""
    test_label = test_set[i][1]
    n_pos = sum(test_label == ref_label for _, ref_label, _ in ref_set)
    n_pos -= (kth - 1)
    if n_pos <= 0:
        return None
    old_recall, old_precision, ap = 0.0, 1.0, 0.0
    intersect_size, j = 0, 0
    _, ranked_list = sim[i].sort(dim=0, descending=True)
    for n, k in enumerate(ranked_list.tolist()):
        if n + 1 < kth:
            continue
        if ref_set[k][1] == test_label:
            intersect_size += 1
        recall = intersect_size / float(n_pos)
        precision = intersect_size / (j + 1.0)
        ap += (recall - old_recall) * ((old_precision + precision) / 2.0)
        old_recall, old_precision = recall, precision
        j += 1
    return ap


def mean_avg_precision(sim, test_set, ref_set, kth=1):
    """reference: utils/metrics.py:48-55"""
    aps = []
    for i in range(sim.size(0)):
        ap = avg_precision(sim, i, test_set, ref_set, kth)
        if ap is not None:
            aps.append(ap)
    return sum(aps) / float(len(aps))
