// Database-side augmentation: replaces every embedding by a weighted sum of its
// nearest neighbours WITHIN its own instance (label), renormalised.
//   reference: test/instance_avg.py:7-33  (caller test/siamese_regions_test.py:83-87)
// The reference materialises sim = mm(E, E^T) and masks everything that is not
// the same label with -2 before sorting each row; only the same-label entries
// of a row can ever be used.  Here one CTA per row computes exactly those
// similarities (exact dots, fp64 accumulation) -- no N x N matrix -- sorts them,
// and forms the weighted sum in the reference's operation order (fp32 multiply
// then add, j = 0 .. nn-1), then divides by (||agg||_2 + 1e-10)  (eps OUTSIDE
// the norm here, unlike NormalizeL2).
#include "isb_host.cuh"

namespace isb {

constexpr int kDbaThreads = 256;
constexpr int kDbaMaxMembers = 2048;  // same-label items per row held in shared memory

__global__ void __launch_bounds__(kDbaThreads)
instance_avg_kernel(const float* __restrict__ emb, const int* __restrict__ label, int N, int D, int k,
                    float* __restrict__ out, int* __restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t dba_smem[];
  float* qs = reinterpret_cast<float*>(dba_smem);      // [D] this row
  __shared__ double m_sim[kDbaMaxMembers];
  __shared__ int m_col[kDbaMaxMembers];
  __shared__ int m_cnt;
  __shared__ float part[kDbaThreads / 32];
  __shared__ float s_norm;
  const int i = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* er = emb + static_cast<size_t>(i) * D;
  float* orow = out + static_cast<size_t>(i) * D;
  if (tid == 0) m_cnt = 0;
  for (int d = tid; d < D; d += kDbaThreads) qs[d] = __ldg(er + d);
  __syncthreads();
  // ---- members: same label, not the row itself (:18-23)
  const int lab = label[i];
  for (int j0 = 0; j0 < N; j0 += kDbaThreads) {
    const int j = j0 + tid;
    const bool is_m = j < N && j != i && __ldg(label + j) == lab;
    // keep members in index order inside each 256-wide slab: ballot + prefix
    const uint32_t b = __ballot_sync(0xffffffffu, is_m);
    __shared__ int slab[kDbaThreads / 32];
    if (lane == 0) slab[warp] = __popc(b);
    __syncthreads();
    int base = m_cnt;
    for (int w = 0; w < warp; ++w) base += slab[w];
    if (is_m) {
      const int pos = base + __popc(b & ((1u << lane) - 1u));
      if (pos < kDbaMaxMembers) m_col[pos] = j;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kDbaThreads / 32; ++w) t += slab[w];
      m_cnt += t;
    }
    __syncthreads();
  }
  int nm = m_cnt;
  if (nm > kDbaMaxMembers) {
    if (tid == 0) atomicAdd(overflow, 1);
    nm = kDbaMaxMembers;
  }
  // ---- exact similarities of the members (one warp each)
  for (int m = warp; m < nm; m += kDbaThreads / 32) {
    const float* r = emb + static_cast<size_t>(m_col[m]) * D;
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) acc = fma(static_cast<double>(qs[d]), static_cast<double>(__ldg(r + d)), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) m_sim[m] = acc;
  }
  int n2 = 2;
  while (n2 < nm) n2 <<= 1;
  for (int m = nm + tid; m < n2 && m < kDbaMaxMembers; m += kDbaThreads) { m_sim[m] = -INFINITY; m_col[m] = 0x7FFFFFFF; }
  __syncthreads();
  // ---- descending sort (ties -> lower index)
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < n2; t += kDbaThreads) {
        const int partner = t ^ stride;
        if (partner > t) {
          const bool up = (t & size) == 0;
          const double sa = m_sim[t], sb = m_sim[partner];
          const int ca = m_col[t], cb = m_col[partner];
          const bool a_first = (sa > sb) || (sa == sb && ca < cb);
          if (a_first != up) { m_sim[t] = sb; m_sim[partner] = sa; m_col[t] = cb; m_col[partner] = ca; }
        }
      }
      __syncthreads();
    }
  }
  // ---- weighted sum (:19-31)
  int nn = nm;
  if (k >= 0 && k < nn) nn = k;
  if (nn <= 0) {
    for (int d = tid; d < D; d += kDbaThreads) orow[d] = qs[d];
    return;
  }
  float sq = 0.f;
  for (int d = tid; d < D; d += kDbaThreads) {
    float agg = qs[d];
    for (int j = 0; j < nn; ++j) {
      const float w = static_cast<float>(static_cast<double>(nn - j) / static_cast<double>(nn + 1));
      agg = __fadd_rn(agg, __fmul_rn(__ldg(emb + static_cast<size_t>(m_col[j]) * D + d), w));
    }
    qs[d] = agg;
    sq = fmaf(agg, agg, sq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) part[warp] = sq;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < kDbaThreads / 32; ++w) t += part[w];
    s_norm = sqrtf(t) + 1e-10f;
  }
  __syncthreads();
  const float norm = s_norm;
  for (int d = tid; d < D; d += kDbaThreads) orow[d] = qs[d] / norm;
}

}  // namespace isb

using namespace isb;

extern "C" int isb_instance_avg(const float* emb, const int32_t* label, int64_t N, int64_t D, int k,
                                float* out, int32_t* overflow, void* stream) {
  ISB_CHECK_ARG(N >= 0 && D > 0 && N < (1ll << 31), "isb_instance_avg: bad shape");
  if (N == 0) return ISB_OK;
  ISB_CHECK_ARG(emb && label && out && overflow, "isb_instance_avg: null pointer");
  ISB_CHECK_ARG(emb != out, "isb_instance_avg: out must not alias emb (rows read their neighbours)");
  const size_t smem = static_cast<size_t>(D) * 4;
  ISB_CHECK_ARG(smem <= 160 * 1024, "isb_instance_avg: D too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ISB_CUDA(cudaMemsetAsync(overflow, 0, 4, st));
  if (smem > 16 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(instance_avg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  instance_avg_kernel<<<static_cast<unsigned>(N), kDbaThreads, smem, st>>>(emb, label, (int)N, (int)D, k, out, overflow);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}
