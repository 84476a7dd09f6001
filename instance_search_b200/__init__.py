"""instance_search_b200 -- B200 (sm_100a) implementation of the retrieval hot
path of maxgreat/Instance-Search (region-descriptor aggregation, cosine top-k
search, all-pairs similarities + negative selection) behind the reference's
own Python operator surface.  See DESIGN.md / INTEGRATION.md.
"""

from . import _lib, ops  # noqa: F401
from . import torch_ops  # noqa: F401  (registers torch.ops.isb.*)
from ._lib import IsbError  # noqa: F401
