// C[M,N] fp32 = A[M,K] . B[N,K]^T (+ bias): the tcgen05 mainloop with a plain
// fp32 store epilogue and deterministic split-K.
//
// replaces torch.mm(E, E.t()) (utils/train_siamese.py:53, test/instance_avg.py:12),
// nn.Linear(100352, D) (model/siamese.py:180) and the 1x1-conv window classifier
// (model/siamese.py:188) of the reference.
#include "isb_host.cuh"
#include "isb_gemm_core.cuh"

namespace isb {

struct StoreEpiParams {
  float* out;            // C (splits == 1) or the partial buffer [splits][M][ldo]
  long long ldo;         // leading dimension of out
  long long split_stride;  // elements between consecutive split partials
  const float* bias;     // added only when splits == 1 (else in the reduce kernel)
  int M, N;
  int vec_ok;            // out rows are 16-byte aligned
};

struct StoreEpilogue {
  using Params = StoreEpiParams;
  const Params& p;
  const int row_in_tile;
  __device__ StoreEpilogue(const Params& p_, int r) : p(p_), row_in_tile(r) {}
  __device__ __forceinline__ void begin_segment(const Segment&) {}
  __device__ __forceinline__ void end_segment(const Segment&) {}

  __device__ __forceinline__ void tile(const Segment& seg, int nt, uint32_t tmem_acc,
                                       uint64_t* tmem_empty_bar) {
    const int row = seg.m_block * kBM + row_in_tile;
    float* orow = p.out + static_cast<long long>(seg.aux) * p.split_stride +
                  static_cast<long long>(row) * p.ldo;
    const int col0 = nt * kBN;
    uint32_t v[2][32];
    ptx::tmem_ld_32x32b_x32(tmem_acc, v[0]);
#pragma unroll
    for (int c = 0; c < kBN / 32; ++c) {
      ptx::tmem_ld_wait();
      if (c + 1 < kBN / 32) {
        ptx::tmem_ld_32x32b_x32(tmem_acc + (c + 1) * 32, v[(c + 1) & 1]);
      } else {
        ptx::tc_fence_before();
        ptx::mbar_arrive(tmem_empty_bar);
      }
      const int cb = col0 + c * 32;
      if (row < p.M && cb < p.N) {
        const uint32_t(&x)[32] = v[c & 1];
        if (p.vec_ok && cb + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(x[j]), __uint_as_float(x[j + 1]),
                                   __uint_as_float(x[j + 2]), __uint_as_float(x[j + 3]));
            if (p.bias != nullptr) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cb + j));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            *reinterpret_cast<float4*>(orow + cb + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cb + j < p.N) {
              float o = __uint_as_float(x[j]);
              if (p.bias != nullptr) o += __ldg(p.bias + cb + j);
              orow[cb + j] = o;
            }
          }
        }
      }
      __syncwarp();
    }
  }
};

// One n-tile per segment.  m-block varies fastest so the CTAs running
// concurrently share a B tile; then the k-split, then the n-tile.
struct PlainSched {
  int m_blocks, n_tiles, k_blocks, splits;
  __device__ __forceinline__ void gate(const Segment&, int, int, int) const {}
  __device__ __forceinline__ void leave(const Segment&) const {}
  __device__ __forceinline__ int num_segments() const { return m_blocks * n_tiles * splits; }
  __device__ __forceinline__ Segment segment(int s) const {
    Segment seg;
    const int t = s / m_blocks;
    seg.m_block = s - t * m_blocks;
    const int nt = t / splits;
    const int sp = t - nt * splits;
    seg.nt_begin = nt;
    seg.nt_end = nt + 1;
    seg.kb_begin = static_cast<int>(static_cast<long long>(sp) * k_blocks / splits);
    seg.kb_end = static_cast<int>(static_cast<long long>(sp + 1) * k_blocks / splits);
    seg.aux = sp;
    return seg;
  }
};

// C[m][n] = sum_s partial[s][m][n] (fixed order) + bias[n]
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, long long split_stride,
                                     long long ldp, int splits, int M, int N,
                                     const float* __restrict__ bias, float* __restrict__ C, long long ldc) {
  const long long total = static_cast<long long>(M) * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / N);
    const int n = static_cast<int>(i - static_cast<long long>(m) * N);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[s * split_stride + m * ldp + n];
    if (bias != nullptr) acc += bias[n];
    C[m * ldc + n] = acc;
  }
}

}  // namespace isb

using namespace isb;

extern "C" size_t isb_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K, int splits) {
  (void)K;
  if (splits <= 1 || M <= 0 || N <= 0) return 0;
  const size_t ldp = align_up(static_cast<size_t>(N), 4);
  return align_up(static_cast<size_t>(splits) * M * ldp * 4, 1024);
}

static int gemm_nt_impl(const char* fn, const uint16_t* A, const uint16_t* A_lo, int64_t lda,
                        const uint16_t* B, const uint16_t* B_lo, int64_t ldb, int64_t M, int64_t N,
                        int64_t K, const float* bias, float* C, int64_t ldc, int splits, void* workspace,
                        size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(A && B && C, "%s: null pointer", fn);
  ISB_CHECK_ARG((A_lo == nullptr) == (B_lo == nullptr), "%s: A_lo and B_lo go together", fn);
  ISB_CHECK_ARG(M > 0 && N > 0 && K > 0, "%s: empty problem", fn);
  ISB_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 29), "%s: dimension too large", fn);
  ISB_CHECK_ARG(lda >= K && ldb >= K && lda % 8 == 0 && ldb % 8 == 0,
                "%s: lda/ldb must be >= K and multiples of 8", fn);
  ISB_CHECK_ARG(ldc >= N, "%s: ldc < N", fn);
  int rc = isb_check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool split_ops = A_lo != nullptr;
  const int m_blocks = static_cast<int>((M + kBM - 1) / kBM);
  const int n_tiles = static_cast<int>((N + kBN - 1) / kBN);
  const int kb_term = static_cast<int>((K + kBK - 1) / kBK);
  const int k_blocks = split_ops ? 3 * kb_term : kb_term;
  if (splits < 1) splits = 1;
  if (splits > k_blocks) splits = k_blocks;
  const size_t need = isb_gemm_nt_workspace_bytes(M, N, K, splits);
  if (splits > 1 && (workspace == nullptr || workspace_bytes < need)) {
    set_error("%s: workspace too small (need %zu bytes, got %zu)", fn, need, workspace_bytes);
    return ISB_ERR_WORKSPACE;
  }
  CUtensorMap ta, tb, ta_lo, tb_lo;
  rc = make_tmap_bf16_k64(&ta, A, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_k64(&tb, B, N, K, ldb, kBN);
  if (rc) return rc;
  ta_lo = ta;
  tb_lo = tb;
  if (split_ops) {
    rc = make_tmap_bf16_k64(&ta_lo, A_lo, M, K, lda, kBM);
    if (rc) return rc;
    rc = make_tmap_bf16_k64(&tb_lo, B_lo, N, K, ldb, kBN);
    if (rc) return rc;
  }

  PlainSched sched{m_blocks, n_tiles, k_blocks, splits};
  StoreEpiParams ep;
  ep.M = static_cast<int>(M);
  ep.N = static_cast<int>(N);
  if (splits == 1) {
    ep.out = C;
    ep.ldo = ldc;
    ep.split_stride = 0;
    ep.bias = bias;
  } else {
    ep.out = static_cast<float*>(workspace);
    ep.ldo = static_cast<long long>(align_up(static_cast<size_t>(N), 4));
    ep.split_stride = static_cast<long long>(M) * ep.ldo;
    ep.bias = nullptr;
  }
  ep.vec_ok = ((reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && ep.ldo % 4 == 0 &&
               (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0)) ? 1 : 0;

  const long long segs = static_cast<long long>(m_blocks) * n_tiles * splits;
  const int sms = device_sm_count();
  const int grid = static_cast<int>(segs < sms ? segs : sms);
  auto kern = gemm_tc_kernel<PlainSched, StoreEpilogue>;
  ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  kern<<<grid, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, split_ops ? kb_term : kSingleTerm,
                                                   sched, ep);
  ISB_CUDA(cudaGetLastError());
  if (splits > 1) {
    const long long total = static_cast<long long>(M) * N;
    const long long blocks = (total + 255) / 256;
    splitk_reduce_kernel<<<static_cast<int>(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(
        ep.out, ep.split_stride, ep.ldo, splits, ep.M, ep.N, bias, C, ldc);
    ISB_CUDA(cudaGetLastError());
  }
  return ISB_OK;
}

extern "C" int isb_gemm_nt(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, int64_t M,
                           int64_t N, int64_t K, const float* bias, float* C, int64_t ldc, int splits,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return gemm_nt_impl("isb_gemm_nt", A, nullptr, lda, B, nullptr, ldb, M, N, K, bias, C, ldc, splits,
                      workspace, workspace_bytes, stream);
}

extern "C" int isb_gemm_nt_split(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda,
                                 const uint16_t* B_hi, const uint16_t* B_lo, int64_t ldb, int64_t M,
                                 int64_t N, int64_t K, const float* bias, float* C, int64_t ldc,
                                 int splits, void* workspace, size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(A_lo && B_lo, "isb_gemm_nt_split: null pointer");
  return gemm_nt_impl("isb_gemm_nt_split", A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, bias, C, ldc,
                      splits, workspace, workspace_bytes, stream);
}
