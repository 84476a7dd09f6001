"""Development probe: kernel time of the tensor-core pooling kernel under its ablation
switches (ISB_TC_DEBUG; bit 8 = stop after the pool kernel), measured with the torch
profiler (CUPTI kernel durations, no host overhead)."""
import os, sys, subprocess
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from instance_search_b200 import regions
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
B, C, H = 256, 2048, int(os.environ.get("TC_H", "14"))
hw = regions.HeadWeights(torch.randn(464, C, device=dev, generator=g) / C ** 0.5, 0.01 * torch.randn(464, device=dev, generator=g),
                         0.01 * torch.randn(C * 49, device=dev, generator=g), torch.randn(64, C * 49, device=dev, generator=g) / (C * 49) ** 0.5,
                         0.01 * torch.randn(64, device=dev, generator=g))
x = torch.relu(torch.randn(B, C, H, H, device=dev, generator=g))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for it in range(3):
    regions.region_select(x, hw, 8, (7, 7))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(5):
        flush.zero_()
        regions.region_select(x, hw, 8, (7, 7))
    torch.cuda.synchronize()
for ev in prof.key_averages():
    if "region_pool" in ev.key:
        print("ISB_TC_DEBUG=%s  %s  avg %.1f us (n=%d)" % (os.environ.get("ISB_TC_DEBUG"), ev.key[:40], ev.device_time_total / ev.count, ev.count))
'''
for flags in [int(a) for a in sys.argv[1:]] or (8, 9, 10, 12, 11, 14, 15, 31):
    env = dict(os.environ, ISB_TC_DEBUG=str(flags))
    subprocess.run([sys.executable, "-c", code], env=env)
