"""Development probe: isb_gemm_nt vs torch.matmul (cuBLAS) on bf16 operands."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instance_search_b200 import ops

def bench(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

for (M, N, K) in [(8192, 8192, 8192), (10000, 100000, 2048), (16384, 16384, 2048)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    fl = 2.0 * M * N * K
    t_isb = bench(lambda: ops.gemm_nt(a, b))
    t_cub = bench(lambda: a @ b.t())
    print(json.dumps({"M": M, "N": N, "K": K, "isb_ms": t_isb, "isb_tflops": fl / t_isb / 1e9,
                      "cublas_ms": t_cub, "cublas_tflops": fl / t_cub / 1e9}))
    del a, b
