"""Oracle (test infrastructure): database-side augmentation (DBA).

Follows /root/reference/test/instance_avg.py:7-33.
"""

import torch

from .mining import get_lab_indicators


def instance_avg(embeddings, dataset, k=-1):
    """Replace each embedding by a weighted sum of its k nearest same-label
    neighbours, renormalised with eps OUTSIDE the norm (:32).

    reference: test/instance_avg.py:7-33
    """
    sim = torch.mm(embeddings, embeddings.t())
    lab_ind = get_lab_indicators(dataset)
    new_embeddings = embeddings.clone()
    for i, (_, lab, _) in enumerate(dataset):
        num_neighbors = int(lab_ind[lab].sum()) - 1
        if k >= 0 and k < num_neighbors:
            num_neighbors = k
        if num_neighbors <= 0:
            new_embeddings[i] = embeddings[i]
            continue
        sim[i, i] = -2
        sim[i][(1 - lab_ind[lab]).bool()] = -2
        _, best_neighbors = torch.sort(sim[i], dim=0, descending=True)
        agg_embedding = embeddings[i].clone()
        for j in range(num_neighbors):
            weight = (num_neighbors - j) / float(num_neighbors + 1)
            agg_embedding += embeddings[best_neighbors[j]] * weight
        new_embeddings[i] = agg_embedding / (agg_embedding.norm() + 1e-10)
    return new_embeddings
