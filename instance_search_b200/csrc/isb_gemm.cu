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

  // pass nt of a segment = chunk nt of the split's k-range (PlainSched): the first pass stores
  // the tile (+ bias), the later ones add to it -- every output element is owned by exactly one
  // thread of one CTA, so the read-modify-write is plain, ordered and deterministic (the slab
  // stays in L2 between passes).  The running sums of a 32-column chunk are fetched BEFORE the
  // wait for the accumulator chunk, so the L2 round trip hides behind the TMEM load.
  __device__ __forceinline__ void tile(const Segment& seg, int nt, uint32_t tmem_acc,
                                       uint64_t* tmem_empty_bar) {
    const int row = seg.m_block * kBM + row_in_tile;
    float* orow = p.out + static_cast<long long>(seg.aux) * p.split_stride +
                  static_cast<long long>(row) * p.ldo;
    const int col0 = seg.n_tile * kBN;
    const bool first = nt == seg.nt_begin;
    // fast path: the tile's columns and ALL 32 rows of this warp in bounds, 16-byte rows -- the
    // predicate must be warp-uniform (the general path ends in a __syncwarp)
    const bool fast = p.vec_ok && (row | 31) < p.M && col0 + kBN <= p.N;
    uint32_t v0[32], v1[32];
    float4 o0[8], o1[8];
    ptx::tmem_ld_32x32b_x32(tmem_acc, v0);
    if (fast && !first) fetch(o0, orow + col0);
#pragma unroll 1
    for (int it = 0; it < kBN / 64; ++it) {
      if (fast && !first) fetch(o1, orow + col0 + it * 64 + 32);
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(tmem_acc + it * 64 + 32, v1);
      if (fast) emit_fast(v0, o0, orow + col0 + it * 64, col0 + it * 64, first);
      else emit(v0, orow, row, col0 + it * 64, first);
      if (fast && !first && it + 1 < kBN / 64) fetch(o0, orow + col0 + it * 64 + 64);
      ptx::tmem_ld_wait();
      if (it + 1 < kBN / 64) {
        ptx::tmem_ld_32x32b_x32(tmem_acc + it * 64 + 64, v0);
      } else {
        // the whole accumulator is in registers: hand the TMEM buffer back
        ptx::tc_fence_before();
        ptx::mbar_arrive(tmem_empty_bar);
      }
      if (fast) emit_fast(v1, o1, orow + col0 + it * 64 + 32, col0 + it * 64 + 32, first);
      else emit(v1, orow, row, col0 + it * 64 + 32, first);
    }
  }

  __device__ __forceinline__ void fetch(float4 (&o)[8], const float* src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = s4[j];
  }

  // 32 columns, fully in bounds: first pass -> acc (+ bias); later passes -> running sum + acc
  __device__ __forceinline__ void emit_fast(const uint32_t (&x)[32], const float4 (&old)[8], float* dst, int cb,
                                            bool first) {
    float4* o4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 o = make_float4(__uint_as_float(x[4 * j]), __uint_as_float(x[4 * j + 1]),
                             __uint_as_float(x[4 * j + 2]), __uint_as_float(x[4 * j + 3]));
      if (first) {
        if (p.bias != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cb) + j);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
      } else {
        o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w;
      }
      o4[j] = o;
    }
    __syncwarp();   // same convergence point as the general path: the next tcgen05.ld is .aligned
  }

  // general path (ragged tile / unaligned rows)
  __device__ __forceinline__ void emit(const uint32_t (&x)[32], float* orow, int row, int cb, bool first) {
    if (row < p.M && cb < p.N) {
      if (p.vec_ok && cb + 32 <= p.N) {
        float4* o4 = reinterpret_cast<float4*>(orow + cb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(__uint_as_float(x[4 * j]), __uint_as_float(x[4 * j + 1]),
                                 __uint_as_float(x[4 * j + 2]), __uint_as_float(x[4 * j + 3]));
          if (first) {
            if (p.bias != nullptr) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cb) + j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
          } else {
            const float4 b = o4[j];
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          o4[j] = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (cb + j < p.N) {
            float o = __uint_as_float(x[j]);
            if (first) {
              if (p.bias != nullptr) o += __ldg(p.bias + cb + j);
            } else {
              o += orow[cb + j];
            }
            orow[cb + j] = o;
          }
        }
      }
    }
    __syncwarp();
  }
};

// One n-tile per segment.  m-block varies fastest so the CTAs running
// concurrently share a B tile; then the k-split, then the n-tile.
// The k-range of a segment is accumulated in passes of at most chunk_kb k-blocks (see the
// scheduler interface in isb_gemm_core.cuh): pass c of the segment covers k-blocks
// [kb_begin + c * chunk_kb, ...) and the epilogue sums the passes in fp32.
constexpr int kAccChunkKb = 64;   // 256 tcgen05.mma accumulation steps per chunk (error 6.6e-6 vs 5.1e-6 at 32,
                                  // 4.6e-6 = the 16-bit representation floor; tools/gemm_precision.py)

struct PlainSched {
  int m_blocks, n_tiles, k_blocks, splits, chunk_kb;
  __device__ __forceinline__ void gate(const Segment&, int, int, int) const {}
  __device__ __forceinline__ void leave(const Segment&) const {}
  __device__ __forceinline__ int num_segments() const { return m_blocks * n_tiles * splits; }
  __device__ __forceinline__ Segment segment(int s) const {
    Segment seg;
    const int t = s / m_blocks;
    seg.m_block = s - t * m_blocks;
    const int nt = t / splits;
    const int sp = t - nt * splits;
    seg.kb_begin = static_cast<int>(static_cast<long long>(sp) * k_blocks / splits);
    seg.kb_end = static_cast<int>(static_cast<long long>(sp + 1) * k_blocks / splits);
    seg.nt_begin = 0;                                              // passes = chunks of the k-range
    seg.nt_end = (seg.kb_end - seg.kb_begin + chunk_kb - 1) / chunk_kb;
    seg.aux = sp;
    seg.n_tile = nt;
    return seg;
  }
  __device__ __forceinline__ int b_tile(const Segment& seg, int) const { return seg.n_tile; }
  __device__ __forceinline__ void kb_range(const Segment& seg, int pass, int& kb0, int& kb1) const {
    kb0 = seg.kb_begin + pass * chunk_kb;
    kb1 = min(kb0 + chunk_kb, seg.kb_end);
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
  // CTA-pair kernel (one tcgen05.mma.cta_group::2 of M = 256 per pair, every B tile staged once
  // per 256 output rows instead of once per 128: the single-CTA kernel is shared-memory-port
  // bound) whenever there are at least two row blocks; ISB_OPT_GEMM_PAIR = 0 keeps
  // the single-CTA kernel.
  const int sms = device_sm_count();
  const bool pair = m_blocks >= 2 && sms >= 2 && option(ISB_OPT_GEMM_PAIR, 1) != 0;
  const int b_box = pair ? kPairBRows : kBN;
  CUtensorMap ta, tb, ta_lo, tb_lo;
  rc = make_tmap_bf16_k64(&ta, A, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_k64(&tb, B, N, K, ldb, b_box);
  if (rc) return rc;
  ta_lo = ta;
  tb_lo = tb;
  if (split_ops) {
    rc = make_tmap_bf16_k64(&ta_lo, A_lo, M, K, lda, kBM);
    if (rc) return rc;
    rc = make_tmap_bf16_k64(&tb_lo, B_lo, N, K, ldb, b_box);
    if (rc) return rc;
  }

  // accumulation chunks: equal parts of at most kAccChunkKb k-blocks of every split's range
  const int kb_split = (k_blocks + splits - 1) / splits;
  const int n_chunks = (kb_split + kAccChunkKb - 1) / kAccChunkKb;
  const int chunk_kb = (kb_split + n_chunks - 1) / n_chunks;
  // the scheduler counts row blocks in the kernel's unit: 128 rows, or 256 for the pair kernel
  PlainSched sched{pair ? (m_blocks + 1) / 2 : m_blocks, n_tiles, k_blocks, splits, chunk_kb};
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

  const long long segs = static_cast<long long>(sched.m_blocks) * n_tiles * splits;
  if (pair) {
    const int pairs = static_cast<int>(segs < sms / 2 ? segs : sms / 2);
    auto kern = gemm_tc_pair_kernel<PlainSched, StoreEpilogue>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    kern<<<2 * pairs, kGemmThreads, kPairSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, split_ops ? kb_term : kSingleTerm,
                                                         sched, ep);
  } else {
    const int grid = static_cast<int>(segs < sms ? segs : sms);
    auto kern = gemm_tc_kernel<PlainSched, StoreEpilogue>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    kern<<<grid, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, split_ops ? kb_term : kSingleTerm,
                                                     sched, ep);
  }
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
