// Ranking reductions over a similarity matrix that the reference's metrics take
// on the host (utils/metrics.py): the k-th best column of every row
// (precision1, :11-13) and the rank positions of selected columns in the
// descending sort of their row (avg_precision, :33-43) -- the only thing the AP
// loop needs from the full sort.  Bandwidth-bound: each row is streamed once
// (kth passes for the k-th best), 128-bit loads when the row is aligned.
#include "isb_host.cuh"

namespace isb {

constexpr int kMetThreads = 256;

// total order of a row's columns: best first = larger score, ties -> lower column
__device__ __forceinline__ bool col_before(float sa, int ca, float sb, int cb) {
  return (sa > sb) || (sa == sb && ca < cb);
}

__device__ __forceinline__ void block_best_f32(float& s, int& c, float* red_s, int* red_c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const int oc = __shfl_xor_sync(0xffffffffu, c, o);
    if (col_before(os, oc, s, c)) { s = os; c = oc; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red_* may still be read from the previous call
  if (lane == 0) { red_s[warp] = s; red_c[warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kMetThreads / 32; ++w)
      if (col_before(red_s[w], red_c[w], red_s[0], red_c[0])) { red_s[0] = red_s[w]; red_c[0] = red_c[w]; }
  }
  __syncthreads();
  s = red_s[0];
  c = red_c[0];
}

// One CTA per row.  Pass p finds the best column strictly after the (p-1)-th
// winner in the total order; kth is small (the reference uses 1 or 2).
__global__ void __launch_bounds__(kMetThreads)
row_kth_largest_kernel(const float* __restrict__ sim, int N, int64_t ld, int kth,
                       float* __restrict__ val, int64_t* __restrict__ idx) {
  __shared__ float red_s[kMetThreads / 32];
  __shared__ int red_c[kMetThreads / 32];
  const float* row = sim + static_cast<int64_t>(blockIdx.x) * ld;
  float bound_s = INFINITY;
  int bound_c = -1;  // (+inf, -1) is before every column
  float best_s = -INFINITY;
  int best_c = 0x7FFFFFFF;
  for (int p = 0; p < kth; ++p) {
    float s = -INFINITY;
    int c = 0x7FFFFFFF;
    for (int j = threadIdx.x; j < N; j += kMetThreads) {
      const float v = __ldg(row + j);
      // NaN never wins; candidates are the columns strictly after the previous winner
      if (col_before(bound_s, bound_c, v, j) && col_before(v, j, s, c)) { s = v; c = j; }
    }
    block_best_f32(s, c, red_s, red_c);
    best_s = s; best_c = c;
    bound_s = s; bound_c = c;
  }
  if (threadIdx.x == 0) {
    val[blockIdx.x] = best_s;
    idx[blockIdx.x] = (best_c == 0x7FFFFFFF) ? -1 : static_cast<int64_t>(best_c);
  }
}

// One CTA per row; P listed columns (slots with col < 0 are unused).
//   1. thresholds (score, col, slot) of the listed columns -> shared memory,
//      bitonic-sorted best first;
//   2. stream the row: binary-search the first threshold the element is before,
//      count it in that bucket (elements worse than every threshold -- almost
//      all -- touch nothing);
//   3. rank of the j-th best threshold = inclusive prefix sum of the buckets.
__global__ void __launch_bounds__(kMetThreads)
row_ranks_kernel(const float* __restrict__ sim, int N, int64_t ld, const int* __restrict__ cols, int P,
                 int P2, int* __restrict__ rank) {
  extern __shared__ __align__(16) uint8_t met_smem[];
  float* ts = reinterpret_cast<float*>(met_smem);   // [P2] threshold scores
  int* tc = reinterpret_cast<int*>(ts + P2);        // [P2] threshold columns
  int* tslot = tc + P2;                             // [P2] slot in cols[]
  int* hist = tslot + P2;                           // [P2]
  const float* row = sim + static_cast<int64_t>(blockIdx.x) * ld;
  const int* rcols = cols + static_cast<int64_t>(blockIdx.x) * P;
  int* rrank = rank + static_cast<int64_t>(blockIdx.x) * P;
  for (int i = threadIdx.x; i < P2; i += kMetThreads) {
    int c = (i < P) ? rcols[i] : -1;
    if (c >= N) c = -1;
    tc[i] = (c >= 0) ? c : 0x7FFFFFFF;
    ts[i] = (c >= 0) ? __ldg(row + c) : -INFINITY;  // unused slots sort last
    tslot[i] = i;
    hist[i] = 0;
  }
  __syncthreads();
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < P2; t += kMetThreads) {
        const int partner = t ^ stride;
        if (partner > t) {
          const bool up = (t & size) == 0;
          const bool a_first = col_before(ts[t], tc[t], ts[partner], tc[partner]);
          if (a_first != up) {
            const float fs = ts[t]; ts[t] = ts[partner]; ts[partner] = fs;
            const int fc = tc[t]; tc[t] = tc[partner]; tc[partner] = fc;
            const int fl = tslot[t]; tslot[t] = tslot[partner]; tslot[partner] = fl;
          }
        }
      }
      __syncthreads();
    }
  }
  // number of valid thresholds (they are sorted first)
  __shared__ int n_valid;
  if (threadIdx.x == 0) {
    int n = 0;
    while (n < P2 && tc[n] != 0x7FFFFFFF) ++n;
    n_valid = n;
  }
  __syncthreads();
  const int nv = n_valid;
  if (nv > 0) {
    const float worst_s = ts[nv - 1];
    const int worst_c = tc[nv - 1];
    for (int j = threadIdx.x; j < N; j += kMetThreads) {
      const float v = __ldg(row + j);
      if (!col_before(v, j, worst_s, worst_c)) continue;  // before no threshold
      int lo = 0, hi = nv - 1;                             // first t with (v, j) before t
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (col_before(v, j, ts[mid], tc[mid])) hi = mid; else lo = mid + 1;
      }
      atomicAdd(&hist[lo], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < nv; ++i) {
      acc += hist[i];
      hist[i] = acc;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P2; i += kMetThreads) {
    const int slot = tslot[i];
    if (slot < P) rrank[slot] = (i < nv) ? hist[i] : -1;
  }
}

}  // namespace isb

using namespace isb;

extern "C" int isb_row_kth_largest(const float* sim, int64_t Q, int64_t N, int64_t ld, int kth, float* val,
                                   int64_t* idx, void* stream) {
  ISB_CHECK_ARG(Q >= 0 && N > 0 && ld >= N && N < (1ll << 31), "isb_row_kth_largest: bad shape");
  ISB_CHECK_ARG(kth >= 1 && kth <= 64 && kth <= N, "isb_row_kth_largest: need 1 <= kth <= min(64, N) (kth=%d)", kth);
  if (Q == 0) return ISB_OK;
  ISB_CHECK_ARG(sim && val && idx, "isb_row_kth_largest: null pointer");
  row_kth_largest_kernel<<<static_cast<unsigned>(Q), kMetThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      sim, static_cast<int>(N), ld, kth, val, idx);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_row_ranks(const float* sim, int64_t Q, int64_t N, int64_t ld, const int32_t* cols, int P,
                             int32_t* rank, void* stream) {
  ISB_CHECK_ARG(Q >= 0 && N > 0 && ld >= N && N < (1ll << 31), "isb_row_ranks: bad shape");
  ISB_CHECK_ARG(P >= 1 && P <= 8192, "isb_row_ranks: need 1 <= P <= 8192 listed columns per row (P=%d)", P);
  if (Q == 0) return ISB_OK;
  ISB_CHECK_ARG(sim && cols && rank, "isb_row_ranks: null pointer");
  int P2 = 2;
  while (P2 < P) P2 <<= 1;
  const size_t smem = static_cast<size_t>(P2) * 16;
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(row_ranks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  row_ranks_kernel<<<static_cast<unsigned>(Q), kMetThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      sim, static_cast<int>(N), ld, cols, P, P2, rank);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}
