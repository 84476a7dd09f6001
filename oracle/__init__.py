"""CPU oracle for the Instance-Search retrieval hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a torch-CPU fp32 restatement of the reference's algorithm for the path
SURVEY.md section 8 names (region-descriptor aggregation, cosine top-k search,
all-pairs similarities + negative selection, and the metrics that consume
them).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or the timed CPU baseline.  Nothing under ``instance_search_b200/``
imports it; the product path raises when the CUDA library is missing.

Pinning status
--------------
The reference ships no unit tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so there is nothing of the reference's own to pin
against.  Instead the restatement is pinned against the *reference code itself
run in the build container*: ``oracle/make_goldens.py`` imports
``/root/reference/model/siamese.py``, ``utils/metrics.py`` and
``test/instance_avg.py`` unmodified (with a two-class shim for the legacy
autograd Functions torch >= 1.5 refuses to run), feeds them seeded inputs,
asserts this package returns bit-identical tensors, and writes the
input/output vectors to ``tests/golden/*.npz``.  ``train/siamese_regions.py``
cannot be imported (it reads a missing data file at import): the lines of its
negative-selection block are executed from the reference's source text instead
(``reference_negative_selection`` in ``oracle/make_goldens.py``) and
``oracle.select_negative`` is asserted against them.

All arithmetic the reference delegates to ``torch`` is delegated to the
torch 2.11 CPU build of this image here as well -- that IS north_star's
"reference's fp32 torch path" (the reference pins no torch version).
"""

from .custom_modules import normalize_l2, shift, triplet_loss  # noqa: F401
from .siamese import (  # noqa: F401
    region_descriptor_forward_single,
    region_descriptor_forward,
    descriptor_forward,
    classif_regions_embedding,
)
from .search import similarity, topk_search, topk_search_f64  # noqa: F401
from .metrics import precision1, avg_precision, mean_avg_precision  # noqa: F401
from .mining import (  # noqa: F401
    get_lab_indicators,
    embeddings_device_dim,
    select_negative,
    select_negative_row,
    select_negatives,
)
from .instance_avg import instance_avg  # noqa: F401
