"""Oracle (test infrastructure): all-pairs similarities and negative selection.

``select_negative`` restates train/siamese_regions.py:106-126.  That module cannot be
imported (its ``P`` reads ``data/CLICIDE_448_train_ms.txt`` at import,
train/siamese_regions_p.py:51), so ``oracle/make_goldens.py`` takes those lines from the
reference's source file as text, executes them on the couples of the ``mining_tiny``
fixture in both modes and asserts that this restatement returns the same indices
(the only line not executed is ``im3 = train_set[k[0]][0]``, :127 -- a data-set look-up
that indexes a 0-dim tensor).  ``get_lab_indicators`` / ``embeddings_device_dim`` are
pinned against utils/train_siamese.py imported as-is.
"""

import torch


def get_lab_indicators(dataset):
    """{label: uint8 mask [N] of items with that label}.

    reference: utils/train_siamese.py:14-25 (duplicate utils/dataset.py:65-76)
    """
    n = len(dataset)
    indicators = {}
    for _, lab1, _ in dataset:
        if lab1 in indicators:
            continue
        indicator = torch.zeros(n, dtype=torch.uint8)
        for i2, (_, lab2, _) in enumerate(dataset):
            if lab1 == lab2:
                indicator[i2] = 1
        indicators[lab1] = indicator
    return indicators


def embeddings_device_dim(cuda_device, feature_dim, embeddings_cuda_size,
                          net_feature_size, n, sim_matrix=False):
    """Placement rule. reference: utils/train_siamese.py:30-43"""
    device = cuda_device
    out_size = feature_dim
    if net_feature_size is not None and out_size <= 0:
        out_size = net_feature_size
    if n * out_size * 4 > embeddings_cuda_size:
        device = -1
    if sim_matrix and n * n * 4 > embeddings_cuda_size:
        device = -1
    return device, out_size


def all_pairs_similarities(embeddings):
    """reference: utils/train_siamese.py:53  similarities = mm(E, E.t())"""
    return torch.mm(embeddings, embeddings.t())


def select_negative_row(row, lab_indicator, i2, semi_hard):
    """The body of select_negative on ONE row of the similarity matrix (row = S[i1, :], so
    large data sets can be checked without materialising S: mm(E[i1:i1+1], E.t()) is that row).
    reference: train/siamese_regions.py:106-126."""
    ind_exl = lab_indicator
    sim_pos = row[i2]
    if semi_hard:
        ind_exl = ind_exl | row.ge(sim_pos).to(torch.uint8)
    if int(ind_exl.sum()) >= row.size(0):
        return -1
    sims = row.clone()
    sims[ind_exl.bool()] = -2
    _, k = sims.max(0)
    return int(k)


def select_negative(similarities, lab_indicator, i1, i2, semi_hard):
    """Negative for the positive couple (i1, i2).

    reference: train/siamese_regions.py:106-129 (same code
    train/siamese_descriptor.py:108-131).  ``semi_hard`` is the reference's
    ``epoch < P.train_epoch_switch``.  Returns the index of the chosen
    negative, or -1 when every item is excluded (the reference then falls back
    to ``choose_rand_neg``, utils/dataset.py:57-61 -- host RNG, out of scope).
    """
    return select_negative_row(similarities[i1], lab_indicator, i2, semi_hard)


def select_negatives(similarities, label_ids, couples, semi_hard):
    """Batched form: one negative per couple. label_ids: int tensor [N]."""
    out = torch.empty(len(couples), dtype=torch.int64)
    for n, (i1, i2) in enumerate(couples):
        ind = (label_ids == label_ids[i1]).to(torch.uint8)
        out[n] = select_negative(similarities, ind, int(i1), int(i2), semi_hard)
    return out
