import sys, torch
sys.path.insert(0, '/root/repo')
from instance_search_b200 import ops, _lib
def run(M,N,K,splits,pair,split_ops=False):
    _lib.set_option("gemm_pair", pair)
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M,K,generator=g).cuda(); b = torch.randn(N,K,generator=g).cuda()
    a16,b16 = ops.to_bf16(a), ops.to_bf16(b)
    try:
        if split_ops:
            y = ops.gemm_nt_split(a16, ops.to_bf16(a,1), b16, ops.to_bf16(b,1), splits=splits, k=K)
        else:
            y = ops.gemm_nt(a16,b16,splits=splits,k=K)
        torch.cuda.synchronize()
        want = a16[:,:K].double()@b16[:,:K].double().t()
        print("OK", M,N,K,splits,"pair",pair, float((y.double()-want).abs().max()/want.abs().max()), flush=True)
    except Exception as e:
        print("FAIL", M,N,K,splits,"pair",pair, repr(e)[:200], flush=True)
        sys.exit(1)
case = sys.argv[1]
M,N,K,splits,pair = [int(v) for v in case.split(",")]
run(M,N,K,splits,pair)
