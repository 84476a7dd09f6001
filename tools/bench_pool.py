"""Times the region pooling kernel alone (isb_region_select's exact_mode = -1 probe) for a few
ring depths and blocks-per-CTA limits: B x C x H x W fp32 maps, L2 flushed before every launch.  Development probe."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import _lib, regions  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
B, C, ncls = 256, 2048, 464
g = torch.Generator(device=dev).manual_seed(5)
hw = regions.HeadWeights(torch.randn(ncls, C, device=dev, generator=g) / C ** 0.5, torch.zeros(ncls, device=dev),
                         torch.zeros(C * 49, device=dev), torch.zeros(8, C * 49, device=dev), torch.zeros(8, device=dev),
                         terms=1)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for size in (14, 32):
    x = torch.relu(torch.randn(B, C, size, size, device=dev, generator=g))
    for stages, gmax in ((3, 0), (3, 4), (3, 8), (3, 16), (3, 32), (2, 0)):   # 0: the plan's own choice
        _lib.set_option("pool_stages", stages)
        _lib.set_option("pool_g", gmax if gmax else None)
        regions._PROBE_CACHE.clear()   # the workspace layout follows the plan
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rd, wr = regions.region_pool_probe(x, hw, 8, (7, 7))
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(json.dumps({"map": size, "stages": stages, "blocks_per_cta": gmax, "ms": ms, "gbs": (rd + wr) / ms / 1e6,
                          "frac_of_measured_hbm": (rd + wr) / ms / 1e6 / peaks["hbm_gbs"]}), flush=True)
    del x
