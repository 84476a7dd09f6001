// Backward passes of the row operators and the triplet loss of
// model/custom_modules.py -- what train/siamese_*.py needs from this path
// besides the forward kernels (isb_elementwise.cu).  All bandwidth-bound.
//   isb_l2norm_rows_backward  <- NormalizeL2Fun.backward   :59-67
//   isb_col_sums              <- ShiftFun.backward          :20-25 (grad_param)
//   isb_triplet_loss_forward  <- TripletLossFun.forward     :153-171
//   isb_triplet_loss_backward <- TripletLossFun.backward    :173-203
#include "isb_host.cuh"

namespace isb {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two values (256 threads); result valid in every thread
__device__ __forceinline__ void block_sum2(float& a, float& b, float* sa, float* sb) {
  a = wsum(a);
  b = wsum(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  float ta = 0.f, tb = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { ta += sa[w]; tb += sb[w]; }
  a = ta;
  b = tb;
}

// gx = (norm2 * g - x * <x, g>) / (norm2 * norm),  norm2 = sum x^2 + eps.  One CTA per row.
__global__ void __launch_bounds__(256)
l2norm_rows_backward_kernel(const float* __restrict__ x, const float* __restrict__ g, int64_t M,
                            int64_t F, float eps, float* __restrict__ gx) {
  __shared__ float sa[8], sb[8];
  for (int64_t row = blockIdx.x; row < M; row += gridDim.x) {
    const float* xr = x + row * F;
    const float* gr = g + row * F;
    float* or_ = gx + row * F;
    float n2 = 0.f, cross = 0.f;
    for (int64_t i = threadIdx.x; i < F; i += 256) {
      const float xv = __ldg(xr + i), gv = __ldg(gr + i);
      n2 = fmaf(xv, xv, n2);
      cross = fmaf(xv, gv, cross);
    }
    block_sum2(n2, cross, sa, sb);
    n2 += eps;
    const float denom = n2 * sqrtf(n2);
    for (int64_t i = threadIdx.x; i < F; i += 256)
      or_[i] = (n2 * __ldg(gr + i) - __ldg(xr + i) * cross) / denom;
    __syncthreads();
  }
}

// out[j] = sum_m g[m, j]  (rows added in order: deterministic).  One thread per column.
__global__ void col_sums_kernel(const float* __restrict__ g, int64_t M, int64_t F, float* __restrict__ out) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (j >= F) return;
  float acc = 0.f;
  for (int64_t m = 0; m < M; ++m) acc += __ldg(g + m * F + j);
  out[j] = acc;
}

// row_loss[i] = clamp(a.n - a.p + margin)                      (normalized)
//             = clamp((|a-p|^2 - |a-n|^2 + 2 margin) / 2)      (otherwise);  one warp per row
__global__ void __launch_bounds__(256)
triplet_rows_kernel(const float* __restrict__ a, const float* __restrict__ p, const float* __restrict__ n,
                    int64_t B, int64_t D, float margin, int normalized, float* __restrict__ row_loss,
                    uint8_t* __restrict__ clamp) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* ar = a + row * D;
  const float* pr = p + row * D;
  const float* nr = n + row * D;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = lane; i < D; i += 32) {
    const float av = __ldg(ar + i), pv = __ldg(pr + i), nv = __ldg(nr + i);
    if (normalized) {
      s1 = fmaf(av, nv, s1);
      s2 = fmaf(av, pv, s2);
    } else {
      s1 = fmaf(av - pv, av - pv, s1);
      s2 = fmaf(av - nv, av - nv, s2);
    }
  }
  s1 = wsum(s1);
  s2 = wsum(s2);
  if (lane == 0) {
    float loss = normalized ? (s1 - s2 + margin) : ((s1 - s2 + margin * 2.f) / 2.f);
    const bool c = loss <= 0.f;  // torch.le(loss, 0)
    clamp[row] = c ? 1 : 0;
    row_loss[row] = c ? 0.f : loss;
  }
}

// loss[0] = sum_i row_loss[i] (/ B), rows added in a fixed order
__global__ void __launch_bounds__(256)
sum_rows_kernel(const float* __restrict__ row_loss, int64_t B, float scale, float* __restrict__ loss) {
  __shared__ float part[256];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += 256) acc += row_loss[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = part[0] * scale;
}

__global__ void triplet_backward_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                        const float* __restrict__ n, int64_t B, int64_t D,
                                        const uint8_t* __restrict__ clamp, const float* __restrict__ gout,
                                        float scale, int normalized, float* __restrict__ ga,
                                        float* __restrict__ gp, float* __restrict__ gn) {
  const int64_t total = B * D;
  const float g = gout[0] * scale;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / D;
    float va = 0.f, vp = 0.f, vn = 0.f;
    if (!clamp[row]) {
      const float av = a[i], pv = p[i], nv = n[i];
      va = (nv - pv) * g;
      vp = (normalized ? -av : (pv - av)) * g;
      vn = (normalized ? av : (av - nv)) * g;
    }
    ga[i] = va; gp[i] = vp; gn[i] = vn;
  }
}

}  // namespace isb

using namespace isb;

extern "C" int isb_l2norm_rows_backward(const float* x, const float* grad_out, int64_t M, int64_t F, float eps,
                                        float* grad_in, void* stream) {
  ISB_CHECK_ARG(M >= 0 && F >= 0, "isb_l2norm_rows_backward: negative shape");
  if (M == 0 || F == 0) return ISB_OK;
  ISB_CHECK_ARG(x && grad_out && grad_in, "isb_l2norm_rows_backward: null pointer");
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  l2norm_rows_backward_kernel<<<static_cast<int>(M < cap ? M : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, grad_out, M, F, eps, grad_in);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_col_sums(const float* g, int64_t M, int64_t F, float* out, void* stream) {
  ISB_CHECK_ARG(M >= 0 && F >= 0, "isb_col_sums: negative shape");
  if (F == 0) return ISB_OK;
  ISB_CHECK_ARG(g && out, "isb_col_sums: null pointer");
  col_sums_kernel<<<static_cast<unsigned>((F + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, M, F, out);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_triplet_loss_forward(const float* anchor, const float* pos, const float* neg, int64_t B,
                                        int64_t D, float margin, int size_average, int normalized,
                                        float* loss, float* row_loss, uint8_t* clamp, void* stream) {
  ISB_CHECK_ARG(B > 0 && D > 0, "isb_triplet_loss_forward: empty batch");
  ISB_CHECK_ARG(anchor && pos && neg && loss && row_loss && clamp, "isb_triplet_loss_forward: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  triplet_rows_kernel<<<static_cast<unsigned>((B * 32 + 255) / 256), 256, 0, st>>>(
      anchor, pos, neg, B, D, margin, normalized, row_loss, clamp);
  ISB_CUDA(cudaGetLastError());
  sum_rows_kernel<<<1, 256, 0, st>>>(row_loss, B, size_average ? 1.f / static_cast<float>(B) : 1.f, loss);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_triplet_loss_backward(const float* anchor, const float* pos, const float* neg, int64_t B,
                                         int64_t D, const uint8_t* clamp, const float* grad_out,
                                         int size_average, int normalized, float* grad_anchor,
                                         float* grad_pos, float* grad_neg, void* stream) {
  ISB_CHECK_ARG(B > 0 && D > 0, "isb_triplet_loss_backward: empty batch");
  ISB_CHECK_ARG(anchor && pos && neg && clamp && grad_out && grad_anchor && grad_pos && grad_neg,
                "isb_triplet_loss_backward: null pointer");
  const int64_t total = B * D;
  const int64_t blocks = (total + 255) / 256;
  triplet_backward_kernel<<<static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(
      anchor, pos, neg, B, D, clamp, grad_out, size_average ? 1.f / static_cast<float>(B) : 1.f, normalized,
      grad_anchor, grad_pos, grad_neg);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}
