// Bandwidth-bound row operators of the descriptor head.
//   isb_l2norm_rows  <- NormalizeL2Fun.forward   model/custom_modules.py:52-57
//   isb_shift_rows   <- ShiftFun.forward         model/custom_modules.py:16-18
#include "isb_host.cuh"

namespace isb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per row (rows of thousands to 100k floats): pass 1 streams the row with
// 128-bit loads and reduces sum(x^2) (warp shuffles + one smem hop), pass 2
// re-reads it (L2-resident: a row is at most ~400 KB) and writes x / sqrt(s+eps).
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
l2norm_rows_cta_kernel(const float* __restrict__ x, int64_t M, int64_t F, float eps,
                       float* __restrict__ y) {
  __shared__ float warp_part[kThreads / 32];
  __shared__ float s_norm;
  for (int64_t row = blockIdx.x; row < M; row += gridDim.x) {
    const float* xr = x + row * F;
    float* yr = y + row * F;
    const bool vec = ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(yr)) & 15) == 0;
    const int64_t F4 = vec ? (F / 4) : 0;
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < F4; i += kThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
      acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    for (int64_t i = F4 * 4 + threadIdx.x; i < F; i += kThreads) {
      const float v = __ldg(xr + i);
      acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = (threadIdx.x < kThreads / 32) ? warp_part[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) s_norm = sqrtf(t + eps);
    }
    __syncthreads();
    const float norm = s_norm;
    for (int64_t i = threadIdx.x; i < F4; i += kThreads) {
      float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);
      v.x = v.x / norm; v.y = v.y / norm; v.z = v.z / norm; v.w = v.w / norm;
      reinterpret_cast<float4*>(yr)[i] = v;
    }
    for (int64_t i = F4 * 4 + threadIdx.x; i < F; i += kThreads) yr[i] = __ldg(xr + i) / norm;
    __syncthreads();
  }
}

// One warp per row (short rows, e.g. final D-dimensional descriptors).
__global__ void __launch_bounds__(256)
l2norm_rows_warp_kernel(const float* __restrict__ x, int64_t M, int64_t F, float eps,
                        float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < M; row += nwarps) {
    const float* xr = x + row * F;
    float* yr = y + row * F;
    float acc = 0.f;
    for (int64_t i = lane; i < F; i += 32) {
      const float v = __ldg(xr + i);
      acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    const float norm = sqrtf(acc + eps);
    for (int64_t i = lane; i < F; i += 32) yr[i] = __ldg(xr + i) / norm;
  }
}

__global__ void shift_rows_kernel(const float* __restrict__ x, const float* __restrict__ param,
                                  int64_t M, int64_t F, float* __restrict__ y) {
  const int64_t total = M * F;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = x[i] + __ldg(param + (i % F));
}

}  // namespace isb

using namespace isb;

extern "C" int isb_l2norm_rows(const float* x, int64_t M, int64_t F, float eps, float* y, void* stream) {
  ISB_CHECK_ARG(M >= 0 && F >= 0, "isb_l2norm_rows: negative shape");
  if (M == 0 || F == 0) return ISB_OK;
  ISB_CHECK_ARG(x && y, "isb_l2norm_rows: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = device_sm_count();
  if (F >= 2048) {
    const int64_t grid = M < static_cast<int64_t>(sms) * 8 ? M : static_cast<int64_t>(sms) * 8;
    l2norm_rows_cta_kernel<256><<<static_cast<int>(grid), 256, 0, st>>>(x, M, F, eps, y);
  } else {
    const int64_t blocks = (M + 7) / 8;
    const int64_t grid = blocks < static_cast<int64_t>(sms) * 8 ? blocks : static_cast<int64_t>(sms) * 8;
    l2norm_rows_warp_kernel<<<static_cast<int>(grid), 256, 0, st>>>(x, M, F, eps, y);
  }
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_shift_rows(const float* x, const float* param, int64_t M, int64_t F, float* y,
                              void* stream) {
  ISB_CHECK_ARG(M >= 0 && F >= 0, "isb_shift_rows: negative shape");
  if (M == 0 || F == 0) return ISB_OK;
  ISB_CHECK_ARG(x && param && y, "isb_shift_rows: null pointer");
  const int64_t total = M * F;
  const int64_t blocks = (total + 255) / 256;
  const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
  shift_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, param, M, F, y);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}
