"""Drop-in for the evaluation section of the reference's ``test/siamese_regions_test.py``
(lines 72-88 of ``main``): embed the test and reference sets, score every test image against
every reference image, report P@1 and mAP, optionally again after database-side feature
augmentation (``--dba``).

The reference's ``main`` also parses the command line, lists and decodes the image folders and
builds the network from a checkpoint -- host-side set-up that is out of scope here;
``evaluate`` takes what that set-up produces: a network in the reference's form, and the two
data sets as lists of ``(image tensor [3, h, w], label, name)`` triples.
"""

from .. import ops
from ..model.nn_utils import set_net_train
from ..train.siamese_regions import get_embeddings
from ..utils import metrics
from .instance_avg import instance_avg


def similarities(test_embeddings, ref_embeddings):
    """``sim = torch.mm(test_embeddings, ref_embeddings.t())`` (test/siamese_regions_test.py:76,85)
    as an fp32-grade split-operand tcgen05 GEMM on the GPU."""
    t, r = test_embeddings.cuda(), ref_embeddings.cuda()
    return ops.gemm_nt_split(ops.to_bf16(t, 0), ops.to_bf16(t, 1), ops.to_bf16(r, 0), ops.to_bf16(r, 1))


def evaluate(net, test_set, ref_set, device=0, labels=None, dba=0, batch_size=32, verbose=True):
    """reference: test/siamese_regions_test.py:72-88.  ``dba``: 0 = no augmentation, k > 0 = the k
    nearest same-instance neighbours, k < 0 = all of them (test/instance_avg.py).  Returns
    ``{"plain": (prec1, correct, total, mAP)}`` plus ``"dba"`` when requested, and prints the
    reference's two report lines."""
    set_net_train(net, False)                                                       # :72
    out_size = getattr(net, "feature_size", None)
    test_embeddings = get_embeddings(net, test_set, device, out_size, batch_size)   # :73
    ref_embeddings = get_embeddings(net, ref_set, device, out_size, batch_size)     # :74
    sim = similarities(test_embeddings, ref_embeddings)                             # :76
    prec1, c, t, _, _ = metrics.precision1(sim, test_set, ref_set)                  # :77
    mAP = metrics.mean_avg_precision(sim, test_set, ref_set)                        # :78
    if verbose:
        print('Descriptor (TEST): {0} / {1} - acc: {2:.4f} - mAP:{3:.4f}'.format(c, t, prec1, mAP))
    result = {"plain": (prec1, c, t, mAP)}
    if dba == 0:                                                                    # :80-81
        return result
    dba_embeddings, dba_set = instance_avg(device, ref_embeddings, ref_set, labels, dba)   # :83-84
    sim = similarities(test_embeddings, dba_embeddings)                             # :85
    prec1, c, t, _, _ = metrics.precision1(sim, test_set, dba_set)                  # :86
    mAP = metrics.mean_avg_precision(sim, test_set, dba_set)                        # :87
    if verbose:
        print('Descriptor (TEST DBA k={4}): {0} / {1} - acc: {2:.4f} - mAP:{3:.4f}'.format(c, t, prec1, mAP, dba))
    result["dba"] = (prec1, c, t, mAP)
    return result
