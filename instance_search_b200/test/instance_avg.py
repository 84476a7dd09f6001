"""Drop-in for the reference's ``test/instance_avg.py`` (database-side augmentation)."""

from .. import mining, ops


def instance_avg(device, embeddings, dataset, labels=None, k=-1):
    """reference: test/instance_avg.py:7-33 (same signature; ``labels`` is unused
    there too).  Returns (new_embeddings, dataset)."""
    ids, _ = mining.label_ids(dataset)
    out = ops.instance_avg(embeddings.cuda(), ids.cuda(), k)
    return (out if device >= 0 else out.cpu()), dataset
