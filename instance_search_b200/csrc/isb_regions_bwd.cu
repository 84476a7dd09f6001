// Backward of the fused region-descriptor head (training path of RegionDescriptorNet,
// reference: model/siamese.py:199-222 through model/custom_modules.py:20-25,59-67).
//
// Forward, per image b with selected windows i < n_b at (r_i, c_i):
//   crop_i = x[b, :, r_i:r_i+fh, c_i:c_i+fw] flattened  [Kin = C*fh*fw]
//   u_b    = sum_i crop_i / sqrt(|crop_i|^2 + eps) + n_b * shift
//   y_b    = W u_b + n_b * bias ;  desc_b = y_b / sqrt(|y_b|^2 + eps)
//   cls_out[b, :, i] = Wc mean_i + bc,  mean_i[c] = mean of x[b, c, window_i]
// The dense parts of the backward (g_u = g_y W, dW = g_y^T u, dWc, g_mean) are tcgen05 GEMMs
// (isb_gemm_nt_split); the two kernels here are the bandwidth-bound glue around them:
//   isb_region_crop_stats     per selected window: |crop|^2, <crop, g_u>, the window mean
//   isb_region_scatter_grad   g_x: every pixel gathers the contributions of the windows that
//                             cover it (no atomics: deterministic, overlapping windows add up)
#include "isb_host.cuh"

namespace isb {

constexpr int kBwdThreads = 256;

// one CTA per (window slot i, image b); one warp per channel, lanes over the window's pixels
__global__ void __launch_bounds__(kBwdThreads)
region_crop_stats_kernel(const float* __restrict__ x, int C, int H, int W, int fh, int fw, int k,
                         const int64_t* __restrict__ idx, const int* __restrict__ nsel,
                         const float* __restrict__ g_u, long long ldg, float* __restrict__ n2_out,
                         float* __restrict__ dot_out, float* __restrict__ means) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double red[2][kBwdThreads / 32];
  const bool live = i < nsel[b];
  const int Wo = W - fw + 1;
  const int area = fh * fw;
  double n2 = 0.0, dot = 0.0;
  if (live) {
    const int flat = static_cast<int>(idx[static_cast<long long>(b) * k + i]);
    const int r0 = flat / Wo, c0 = flat - r0 * Wo;
    const float* xb = x + static_cast<long long>(b) * C * H * W;
    const float* gb = g_u != nullptr ? g_u + static_cast<long long>(b) * ldg : nullptr;
    for (int c = warp; c < C; c += kBwdThreads / 32) {
      const float* plane = xb + static_cast<long long>(c) * H * W;
      float s = 0.f;
      for (int p = lane; p < area; p += 32) {
        const int pr = p / fw, pc = p - pr * fw;
        const float v = __ldg(plane + (r0 + pr) * W + c0 + pc);
        s += v;
        n2 = fma(static_cast<double>(v), static_cast<double>(v), n2);
        if (gb != nullptr) dot = fma(static_cast<double>(v), static_cast<double>(__ldg(gb + c * area + p)), dot);
      }
      if (means != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) means[(static_cast<long long>(b) * k + i) * C + c] = s / static_cast<float>(area);
      }
    }
  } else if (means != nullptr) {
    for (int c = threadIdx.x; c < C; c += kBwdThreads) means[(static_cast<long long>(b) * k + i) * C + c] = 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  if (lane == 0) { red[0][warp] = n2; red[1][warp] = dot; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int w = 0; w < kBwdThreads / 32; ++w) { a += red[0][w]; d += red[1][w]; }
    n2_out[static_cast<long long>(b) * k + i] = static_cast<float>(a);
    dot_out[static_cast<long long>(b) * k + i] = static_cast<float>(d);
  }
}

constexpr int kBwdMaxK = 32;

// one thread per element of g_x[b]; the image's windows sit in shared memory
__global__ void __launch_bounds__(kBwdThreads)
region_scatter_grad_kernel(const float* __restrict__ x, int C, int H, int W, int fh, int fw, int k,
                           const int64_t* __restrict__ idx, const int* __restrict__ nsel,
                           const float* __restrict__ g_u, long long ldg, const float* __restrict__ n2,
                           const float* __restrict__ dot, const float* __restrict__ g_mean, float eps,
                           float* __restrict__ g_x) {
  const int b = blockIdx.y;
  __shared__ int s_r0[kBwdMaxK], s_c0[kBwdMaxK];
  __shared__ float s_inv[kBwdMaxK], s_coef[kBwdMaxK];
  const int n = min(nsel[b], k);
  const int Wo = W - fw + 1;
  if (threadIdx.x < n) {
    const int flat = static_cast<int>(idx[static_cast<long long>(b) * k + threadIdx.x]);
    s_r0[threadIdx.x] = flat / Wo;
    s_c0[threadIdx.x] = flat % Wo;
    // d/dx [x / sqrt(|x|^2 + eps)] . g  =  g / n - x <x, g> / n^3,   n = sqrt(|x|^2 + eps)
    const double nn = static_cast<double>(n2[static_cast<long long>(b) * k + threadIdx.x]) + static_cast<double>(eps);
    const double inv = 1.0 / sqrt(nn);
    s_inv[threadIdx.x] = static_cast<float>(inv);
    s_coef[threadIdx.x] = static_cast<float>(static_cast<double>(dot[static_cast<long long>(b) * k + threadIdx.x]) * inv / nn);
  }
  __syncthreads();
  const int HW = H * W, area = fh * fw;
  const float inv_area = 1.f / static_cast<float>(area);
  const long long per_image = static_cast<long long>(C) * HW;
  for (long long e = blockIdx.x * static_cast<long long>(kBwdThreads) + threadIdx.x; e < per_image;
       e += static_cast<long long>(gridDim.x) * kBwdThreads) {
    const int c = static_cast<int>(e / HW);
    const int hw = static_cast<int>(e - static_cast<long long>(c) * HW);
    const int h = hw / W, w = hw - h * W;
    const float xv = x[static_cast<long long>(b) * per_image + e];
    float acc = 0.f;
    for (int i = 0; i < n; ++i) {
      const int pr = h - s_r0[i], pc = w - s_c0[i];
      if (pr < 0 || pr >= fh || pc < 0 || pc >= fw) continue;
      if (g_u != nullptr)
        acc += __ldg(g_u + static_cast<long long>(b) * ldg + c * area + pr * fw + pc) * s_inv[i] - xv * s_coef[i];
      if (g_mean != nullptr) acc += __ldg(g_mean + (static_cast<long long>(b) * k + i) * C + c) * inv_area;
    }
    g_x[static_cast<long long>(b) * per_image + e] = acc;
  }
}

}  // namespace isb

using namespace isb;

static int check_bwd_args(const char* fn, const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh,
                          int fw, int k, const int64_t* idx, const int32_t* nsel) {
  ISB_CHECK_ARG(x && idx && nsel, "%s: null pointer", fn);
  ISB_CHECK_ARG(B > 0 && C > 0 && fh > 0 && fw > 0 && H >= fh && W >= fw && k >= 1 && k <= kBwdMaxK,
                "%s: bad shape (B=%lld C=%lld H=%lld W=%lld window %dx%d k=%d)", fn, (long long)B, (long long)C,
                (long long)H, (long long)W, fh, fw, k);
  ISB_CHECK_ARG(C * H * W < (1ll << 31) && C * fh * fw < (1ll << 31) && B < 65536, "%s: too large", fn);
  return isb_check_device();
}

extern "C" int isb_region_crop_stats(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh, int fw,
                                     int k, const int64_t* idx, const int32_t* nsel, const float* g_u,
                                     int64_t ldg, float* crop_norm2, float* crop_dot, float* win_mean,
                                     void* stream) {
  int rc = check_bwd_args("isb_region_crop_stats", x, B, C, H, W, fh, fw, k, idx, nsel);
  if (rc) return rc;
  ISB_CHECK_ARG(crop_norm2 && crop_dot, "isb_region_crop_stats: null output");
  ISB_CHECK_ARG(g_u == nullptr || ldg >= C * fh * fw, "isb_region_crop_stats: ldg < C*fh*fw");
  region_crop_stats_kernel<<<dim3(static_cast<unsigned>(k), static_cast<unsigned>(B)), kBwdThreads, 0,
                             static_cast<cudaStream_t>(stream)>>>(
      x, (int)C, (int)H, (int)W, fh, fw, k, idx, nsel, g_u, ldg, crop_norm2, crop_dot, win_mean);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_region_scatter_grad(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh,
                                       int fw, int k, const int64_t* idx, const int32_t* nsel, const float* g_u,
                                       int64_t ldg, const float* crop_norm2, const float* crop_dot,
                                       const float* g_mean, float eps, float* g_x, void* stream) {
  int rc = check_bwd_args("isb_region_scatter_grad", x, B, C, H, W, fh, fw, k, idx, nsel);
  if (rc) return rc;
  ISB_CHECK_ARG(g_x && crop_norm2 && crop_dot, "isb_region_scatter_grad: null pointer");
  ISB_CHECK_ARG(g_u == nullptr || ldg >= C * fh * fw, "isb_region_scatter_grad: ldg < C*fh*fw");
  const long long per_image = C * H * W;
  long long bx = (per_image + kBwdThreads - 1) / kBwdThreads;
  if (bx > 148 * 8) bx = 148 * 8;
  region_scatter_grad_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(B)), kBwdThreads, 0,
                               static_cast<cudaStream_t>(stream)>>>(
      x, (int)C, (int)H, (int)W, fh, fw, k, idx, nsel, g_u, ldg, crop_norm2, crop_dot, g_mean, eps, g_x);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}
