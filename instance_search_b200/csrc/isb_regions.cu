// Region-descriptor aggregation of RegionDescriptorNet.forward_single
// (model/siamese.py:185-223) for a whole batch of trunk feature maps.
//
//   1. region_pool_kernel       x[B,C,H,W] fp32 -> window means P[B*H'W', C] bf16
//                               (AvgPool2d(fh x fw, stride 1), :164-166,187) and
//                               per-pixel channel energy partials (for the crop norms)
//   2. gemm_tc_kernel<RowMax>   P . Wc^T + bc  (the 1x1-conv classifier, :188) with
//                               the class-max (:191) fused into the TMEM epilogue
//   3. region_select_kernel     top-(k+margin) windows by the bf16 screen, exact
//                               (fp64-accumulated) re-score of those, top-k (:194),
//                               cls_out (:216), ||crop||_2 of the chosen windows
//   4. region_gather_kernel     u[b] = sum_i crop_i / ||crop_i|| + nsel * shift
//                               (NormalizeL2 + Shift, :218-219, summed BEFORE the
//                               projection -- the Linear is linear) -> bf16 hi (+ lo)
//   5. isb_gemm_nt[_split]      u . W^T   (nn.Linear(100352, D), :180)
//   6. descriptor_finalize      desc = l2norm(y + nsel * bias)   (:220-222)
#include "isb_host.cuh"
#include "isb_gemm_core.cuh"

namespace isb {

__device__ __forceinline__ uint16_t bf16_rn(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
__device__ __forceinline__ float bf16_f(uint16_t h) { return __uint_as_float(static_cast<uint32_t>(h) << 16); }

// ------------------------------------------------------------------ 1. pooling
// One CTA = one image x CB channels.  The CB planes are contiguous in NCHW, so
// the load is one coalesced stream; the 2-D window sum is separable (fw-wide
// row sums, then fh-tall column sums) out of shared memory; the output is
// written channel-contiguous (the K-major operand layout the classifier GEMM
// wants), CB consecutive bf16 per window.  Plane strides are odd so that lanes
// that differ in channel hit different banks.
template <int CB>
__global__ void __launch_bounds__(256)
region_pool_kernel(const float* __restrict__ x, int C, int H, int W, int fh, int fw,
                   uint16_t* __restrict__ P, int ldp, float* __restrict__ e_part) {
  extern __shared__ float pool_smem[];
  const int HW = H * W, Wo = W - fw + 1, Ho = H - fh + 1;
  const int HWp = HW | 1, HRp = (H * Wo) | 1;
  float* plane = pool_smem;               // [CB][HWp]
  float* R = pool_smem + CB * HWp;        // [CB][HRp]
  const int b = blockIdx.y, blk = blockIdx.x, c0 = blk * CB;
  const int nblk = gridDim.x;
  const float* src = x + (static_cast<size_t>(b) * C + c0) * HW;
  const int cvalid = min(CB, C - c0);
  const int tid = threadIdx.x;

  const int total = cvalid * HW;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = tid * 4; i < total; i += 256 * 4) {
      if (i + 4 <= total) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int cb = (i + t) / HW, p = (i + t) - cb * HW;
          plane[cb * HWp + p] = vv[t];
        }
      } else {
        for (int t = i; t < total; ++t) {
          const int cb = t / HW, p = t - cb * HW;
          plane[cb * HWp + p] = __ldg(src + t);
        }
      }
    }
  } else {
    for (int i = tid; i < total; i += 256) {
      const int cb = i / HW, p = i - cb * HW;
      plane[cb * HWp + p] = __ldg(src + i);
    }
  }
  for (int i = total + tid; i < CB * HW; i += 256) {  // channels past C (ragged last block)
    const int cb = i / HW, p = i - cb * HW;
    plane[cb * HWp + p] = 0.f;
  }
  __syncthreads();

  // per-pixel energy of these CB channels (summed over blocks, in block order, later)
  for (int p = tid; p < HW; p += 256) {
    float s = 0.f;
#pragma unroll 8
    for (int cb = 0; cb < CB; ++cb) {
      const float v = plane[cb * HWp + p];
      s = fmaf(v, v, s);
    }
    e_part[(static_cast<size_t>(b) * nblk + blk) * HW + p] = s;
  }
  // horizontal fw-sums
  for (int i = tid; i < CB * H * Wo; i += 256) {
    const int cb = i % CB, t = i / CB;
    const int h = t / Wo, w = t - h * Wo;
    const float* r = plane + cb * HWp + h * W + w;
    float s = 0.f;
    for (int dx = 0; dx < fw; ++dx) s += r[dx];
    R[cb * HRp + h * Wo + w] = s;
  }
  __syncthreads();
  // vertical fh-sums, mean, bf16, channel-contiguous store
  const float inv_area = 1.f / static_cast<float>(fh * fw);
  const size_t row0 = static_cast<size_t>(b) * Ho * Wo;
  for (int i = tid; i < CB * Ho * Wo; i += 256) {
    const int cb = i % CB, win = i / CB;
    const int h = win / Wo, w = win - h * Wo;
    const float* r = R + cb * HRp + h * Wo + w;
    float s = 0.f;
    for (int dy = 0; dy < fh; ++dy) s += r[dy * Wo];
    if (cb < cvalid) P[(row0 + win) * ldp + c0 + cb] = bf16_rn(s * inv_area);
  }
}

// ------------------------------------------------------------------ 2. class-max epilogue
struct RowMaxEpiParams {
  float* row_max;      // [M]
  const float* bias;   // [N]
  int M, N;
};

struct RowMaxEpilogue {
  using Params = RowMaxEpiParams;
  const Params& p;
  const int row_in_tile;
  float m;
  __device__ RowMaxEpilogue(const Params& p_, int r) : p(p_), row_in_tile(r), m(0.f) {}
  __device__ __forceinline__ void begin_segment(const Segment&) { m = -INFINITY; }
  __device__ __forceinline__ void tile(const Segment& seg, int nt, uint32_t tmem_acc,
                                       uint64_t* tmem_empty_bar) {
    const int col0 = nt * kBN;
    uint32_t v0[32], v1[32];
    ptx::tmem_ld_32x32b_x32(tmem_acc, v0);
#pragma unroll 1
    for (int it = 0; it < kBN / 64; ++it) {
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(tmem_acc + it * 64 + 32, v1);
      scan(v0, col0 + it * 64);
      ptx::tmem_ld_wait();
      if (it + 1 < kBN / 64) {
        ptx::tmem_ld_32x32b_x32(tmem_acc + it * 64 + 64, v0);
      } else {
        ptx::tc_fence_before();
        ptx::mbar_arrive(tmem_empty_bar);
      }
      scan(v1, col0 + it * 64 + 32);
    }
  }
  __device__ __forceinline__ void scan(const uint32_t (&v)[32], int cb) {
    if (cb >= p.N) return;  // warp-uniform
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = cb + j;
      const float b = (col < p.N) ? __ldg(p.bias + col) : -INFINITY;
      m = fmaxf(m, __uint_as_float(v[j]) + b);
    }
  }
  __device__ __forceinline__ void end_segment(const Segment& seg) {
    const int row = seg.m_block * kBM + row_in_tile;
    if (row < p.M) p.row_max[row] = m;
  }
};

// one segment = one m-block x ALL n-tiles (the max runs over every class)
struct RowSched {
  int m_blocks, n_tiles, k_blocks;
  __device__ __forceinline__ int num_segments() const { return m_blocks; }
  __device__ __forceinline__ Segment segment(int s) const {
    Segment seg;
    seg.m_block = s; seg.nt_begin = 0; seg.nt_end = n_tiles;
    seg.kb_begin = 0; seg.kb_end = k_blocks; seg.aux = 0;
    return seg;
  }
};

// ------------------------------------------------------------------ 3. selection
constexpr int kSelThreads = 256;
constexpr int kSelMaxCand = 32;   // k + margin
constexpr int kSelChunk = 8;      // candidates re-scored per pass over the classifier weights

__device__ __forceinline__ void block_argmax(float v, int i, float* s_val, int* s_idx, float& out_v,
                                             int& out_i) {
  // larger value first; ties -> lower index
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = v; s_idx[warp] = i; }
  __syncthreads();
  if (warp == 0) {
    v = (lane < kSelThreads / 32) ? s_val[lane] : -INFINITY;
    i = (lane < kSelThreads / 32) ? s_idx[lane] : 0x7FFFFFFF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (lane == 0) { s_val[0] = v; s_idx[0] = i; }
  }
  __syncthreads();
  out_v = s_val[0];
  out_i = s_idx[0];
  __syncthreads();
}

__global__ void __launch_bounds__(kSelThreads)
region_select_kernel(const float* __restrict__ x, int C, int H, int W, int fh, int fw,
                     const float* __restrict__ cls_w, const float* __restrict__ cls_b, int ncls,
                     const float* __restrict__ screen,   // [B*HoWo] class-max by the bf16 GEMM
                     const float* __restrict__ e_part, int nblk, int k, int ncand, float eps,
                     int64_t* __restrict__ idx, int* __restrict__ nsel_out,
                     float* __restrict__ cls_out, float* __restrict__ win_norm) {
  extern __shared__ __align__(16) uint8_t sel_smem_raw[];
  const int Ho = H - fh + 1, Wo = W - fw + 1, nwin = Ho * Wo, HW = H * W;
  float* sc = reinterpret_cast<float*>(sel_smem_raw);          // [nwin]
  float* pooled = sc + ((nwin + 3) & ~3);                      // [kSelChunk][C]
  float* logits = pooled + kSelChunk * C;                      // [ncand][ncls]
  __shared__ float s_val[kSelThreads / 32];
  __shared__ int s_idx[kSelThreads / 32];
  __shared__ int cand[kSelMaxCand];
  __shared__ float cand_max[kSelMaxCand];
  __shared__ int order[kSelMaxCand];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = x + static_cast<size_t>(b) * C * HW;

  for (int i = tid; i < nwin; i += kSelThreads) sc[i] = screen[static_cast<size_t>(b) * nwin + i];
  __syncthreads();
  // ---- candidates: the ncand best windows of the screen
  for (int c = 0; c < ncand; ++c) {
    float v = -INFINITY;
    int vi = 0x7FFFFFFF;
    for (int i = tid; i < nwin; i += kSelThreads) {
      const float s = sc[i];
      if (s > v || (s == v && i < vi)) { v = s; vi = i; }
    }
    float bv; int bi;
    block_argmax(v, vi, s_val, s_idx, bv, bi);
    if (tid == 0) { cand[c] = bi; sc[bi] = -INFINITY; }
    __syncthreads();
  }
  // ---- exact logits of the candidates (fp64 accumulation of fp32 products)
  const double inv_area = 1.0 / static_cast<double>(fh * fw);
  for (int c0 = 0; c0 < ncand; c0 += kSelChunk) {
    const int nc = min(kSelChunk, ncand - c0);
    for (int i = tid; i < nc * C; i += kSelThreads) {
      const int ci = i / C, ch = i - ci * C;
      const int win = cand[c0 + ci];
      const int h = win / Wo, w = win - h * Wo;
      const float* pl = xb + static_cast<size_t>(ch) * HW + h * W + w;
      double s = 0.0;
      for (int dy = 0; dy < fh; ++dy)
        for (int dx = 0; dx < fw; ++dx) s += static_cast<double>(__ldg(pl + dy * W + dx));
      pooled[ci * C + ch] = static_cast<float>(s * inv_area);
    }
    __syncthreads();
    for (int j = warp; j < ncls; j += kSelThreads / 32) {
      const float* wr = cls_w + static_cast<size_t>(j) * C;
      double acc[kSelChunk];
#pragma unroll
      for (int t = 0; t < kSelChunk; ++t) acc[t] = 0.0;
      for (int ch = lane; ch < C; ch += 32) {
        const double wv = static_cast<double>(__ldg(wr + ch));
#pragma unroll
        for (int t = 0; t < kSelChunk; ++t)
          if (t < nc) acc[t] = fma(wv, static_cast<double>(pooled[t * C + ch]), acc[t]);
      }
#pragma unroll
      for (int t = 0; t < kSelChunk; ++t) {
        double a = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && t < nc)
          logits[(c0 + t) * ncls + j] = static_cast<float>(a + static_cast<double>(__ldg(cls_b + j)));
      }
    }
    __syncthreads();
  }
  // ---- exact class-max of every candidate
  for (int c = warp; c < ncand; c += kSelThreads / 32) {
    float m = -INFINITY;
    for (int j = lane; j < ncls; j += 32) m = fmaxf(m, logits[c * ncls + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) cand_max[c] = m;
  }
  __syncthreads();
  const int nsel = min(nwin, k);
  if (tid == 0) {
    // rank the (<= 32) candidates: value desc, window index asc
    for (int c = 0; c < ncand; ++c) order[c] = c;
    for (int i = 1; i < ncand; ++i) {
      const int o = order[i];
      int j = i - 1;
      while (j >= 0 && (cand_max[order[j]] < cand_max[o] ||
                        (cand_max[order[j]] == cand_max[o] && cand[order[j]] > cand[o]))) {
        order[j + 1] = order[j];
        --j;
      }
      order[j + 1] = o;
    }
    nsel_out[b] = nsel;
  }
  __syncthreads();
  for (int i = tid; i < k; i += kSelThreads)
    idx[static_cast<size_t>(b) * k + i] = (i < nsel) ? static_cast<int64_t>(cand[order[i]]) : -1;
  // cls_out[b, cls, i]  (zero beyond nsel, model/siamese.py:207-208)
  for (int t = tid; t < ncls * k; t += kSelThreads) {
    const int j = t / k, i = t - j * k;
    cls_out[(static_cast<size_t>(b) * ncls + j) * k + i] = (i < nsel) ? logits[order[i] * ncls + j] : 0.f;
  }
  // ||crop||: sqrt(sum over the window of the per-pixel energy + eps)
  for (int i = warp; i < k; i += kSelThreads / 32) {
    if (i < nsel) {
      const int win = cand[order[i]];
      const int h = win / Wo, w = win - h * Wo;
      double s = 0.0;
      for (int t = lane; t < fh * fw * nblk; t += 32) {
        const int blk = t / (fh * fw), r = t - blk * (fh * fw);
        const int dy = r / fw, dx = r - dy * fw;
        s += static_cast<double>(e_part[(static_cast<size_t>(b) * nblk + blk) * HW + (h + dy) * W + w + dx]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) win_norm[static_cast<size_t>(b) * k + i] = sqrtf(static_cast<float>(s) + eps);
    } else if (lane == 0) {
      win_norm[static_cast<size_t>(b) * k + i] = 1.f;
    }
  }
}

// ------------------------------------------------------------------ 4. gather
// One thread = 8 consecutive elements of the C*fh*fw region vector (one 16-byte
// bf16 store per term).  u[e] = sum_i x[b, c, h_i+dy, w_i+dx] / norm_i + nsel*shift[e].
__global__ void __launch_bounds__(256)
region_gather_kernel(const float* __restrict__ x, int C, int H, int W, int fh, int fw, int k,
                     const int64_t* __restrict__ idx, const int* __restrict__ nsel_in,
                     const float* __restrict__ win_norm, const float* __restrict__ shift,
                     uint16_t* __restrict__ U_hi, uint16_t* __restrict__ U_lo, int64_t ldu) {
  __shared__ int s_off[kSelMaxCand];
  __shared__ float s_norm[kSelMaxCand];
  const int b = blockIdx.y;
  const int Wo = W - fw + 1, HW = H * W, area = fh * fw;
  const int Kin = C * area;
  const int nsel = nsel_in[b];
  if (threadIdx.x < nsel) {
    const int win = static_cast<int>(idx[static_cast<size_t>(b) * k + threadIdx.x]);
    const int h = win / Wo, w = win - h * Wo;
    s_off[threadIdx.x] = h * W + w;
    s_norm[threadIdx.x] = win_norm[static_cast<size_t>(b) * k + threadIdx.x];
  }
  __syncthreads();
  const int KinP = (Kin + 7) & ~7;  // rows are zero padded to 8 columns
  const int e0 = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (e0 >= KinP) return;
  const float* xb = x + static_cast<size_t>(b) * C * HW;
  uint16_t hi[8], lo[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int e = e0 + t;
    float u = 0.f;
    if (e < Kin) {
      const int c = e / area, r = e - c * area;
      const int dy = r / fw, dx = r - dy * fw;
      const float* pl = xb + static_cast<size_t>(c) * HW + dy * W + dx;
      for (int i = 0; i < nsel; ++i) u += __ldg(pl + s_off[i]) / s_norm[i];
      u += static_cast<float>(nsel) * __ldg(shift + e);
    }
    hi[t] = bf16_rn(u);
    lo[t] = bf16_rn(u - bf16_f(hi[t]));
  }
  uint4 vh, vl;
  vh.x = hi[0] | (uint32_t(hi[1]) << 16); vh.y = hi[2] | (uint32_t(hi[3]) << 16);
  vh.z = hi[4] | (uint32_t(hi[5]) << 16); vh.w = hi[6] | (uint32_t(hi[7]) << 16);
  vl.x = lo[0] | (uint32_t(lo[1]) << 16); vl.y = lo[2] | (uint32_t(lo[3]) << 16);
  vl.z = lo[4] | (uint32_t(lo[5]) << 16); vl.w = lo[6] | (uint32_t(lo[7]) << 16);
  *reinterpret_cast<uint4*>(U_hi + static_cast<size_t>(b) * ldu + e0) = vh;
  if (U_lo != nullptr)  // split operand for the fp32-grade projection (isb_gemm_nt_split)
    *reinterpret_cast<uint4*>(U_lo + static_cast<size_t>(b) * ldu + e0) = vl;
}

// ------------------------------------------------------------------ 6. finalize
__global__ void __launch_bounds__(256)
descriptor_finalize_kernel(const float* __restrict__ y, const float* __restrict__ bias,
                           const int* __restrict__ nsel, int D, float eps, float* __restrict__ desc) {
  __shared__ float part[8];
  __shared__ float s_norm;
  const int b = blockIdx.x;
  const float scale = (nsel != nullptr) ? static_cast<float>(nsel[b]) : 1.f;
  float acc = 0.f;
  for (int j = threadIdx.x; j < D; j += 256) {
    const float v = y[static_cast<size_t>(b) * D + j] + (bias != nullptr ? scale * __ldg(bias + j) : 0.f);
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    s_norm = sqrtf(t + eps);
  }
  __syncthreads();
  const float norm = s_norm;
  for (int j = threadIdx.x; j < D; j += 256) {
    const float v = y[static_cast<size_t>(b) * D + j] + (bias != nullptr ? scale * __ldg(bias + j) : 0.f);
    desc[static_cast<size_t>(b) * D + j] = v / norm;
  }
}

struct RegionPlan {
  int Ho, Wo, nwin, CB, nblk, ncand;
  int64_t ldp;
  size_t off_P, off_epart, off_screen, total;
  size_t pool_smem, sel_smem;
};

static RegionPlan make_region_plan(int64_t B, int C, int H, int W, int ncls, int fh, int fw, int k,
                                   int margin) {
  RegionPlan p;
  p.Ho = H - fh + 1; p.Wo = W - fw + 1; p.nwin = p.Ho * p.Wo;
  p.CB = (H * W <= 400) ? 32 : 16;
  p.nblk = (C + p.CB - 1) / p.CB;
  p.ncand = k + margin;
  if (p.ncand > kSelMaxCand) p.ncand = kSelMaxCand;
  if (p.ncand > p.nwin) p.ncand = p.nwin;
  p.ldp = static_cast<int64_t>(align_up(static_cast<size_t>(C), 8));
  size_t off = 0;
  p.off_P = off;      off = align_up(off + static_cast<size_t>(B) * p.nwin * p.ldp * 2, 1024);
  p.off_epart = off;  off = align_up(off + static_cast<size_t>(B) * p.nblk * H * W * 4, 1024);
  p.off_screen = off; off = align_up(off + static_cast<size_t>(B) * p.nwin * 4, 1024);
  p.total = off;
  p.pool_smem = static_cast<size_t>(p.CB) * (((H * W) | 1) + ((H * p.Wo) | 1)) * 4;
  p.sel_smem = (static_cast<size_t>((p.nwin + 3) & ~3) + static_cast<size_t>(kSelChunk) * C +
                static_cast<size_t>(p.ncand) * ncls) * 4;
  return p;
}

}  // namespace isb

using namespace isb;

extern "C" size_t isb_region_select_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W,
                                                    int64_t ncls, int fh, int fw, int k, int margin) {
  if (B <= 0 || C <= 0 || H < fh || W < fw || fh <= 0 || fw <= 0) return 0;
  return make_region_plan(B, (int)C, (int)H, (int)W, (int)ncls, fh, fw, k, margin).total + 1024;
}

extern "C" int isb_region_select(const float* x, int64_t B, int64_t C, int64_t H, int64_t W,
                                 const float* cls_w, const uint16_t* cls_w_bf16, int64_t ld_w,
                                 const float* cls_b, int64_t ncls, int fh, int fw, int k, int margin,
                                 int64_t* idx, int32_t* nsel, float* cls_out, float* win_norm,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(x && cls_w && cls_w_bf16 && cls_b && idx && nsel && cls_out && win_norm,
                "isb_region_select: null pointer");
  ISB_CHECK_ARG(B > 0 && C > 0 && ncls > 0 && fh > 0 && fw > 0 && H >= fh && W >= fw,
                "isb_region_select: bad shape (B=%lld C=%lld H=%lld W=%lld window %dx%d)", (long long)B,
                (long long)C, (long long)H, (long long)W, fh, fw);
  ISB_CHECK_ARG(k >= 1 && k <= kSelMaxCand && margin >= 0, "isb_region_select: need 1 <= k <= %d", kSelMaxCand);
  ISB_CHECK_ARG(ld_w >= C && ld_w % 8 == 0, "isb_region_select: bad ld_w");
  ISB_CHECK_ARG(B * (H - fh + 1) * (W - fw + 1) < (1ll << 31), "isb_region_select: too many windows");
  int rc = isb_check_device();
  if (rc) return rc;
  const RegionPlan p = make_region_plan(B, (int)C, (int)H, (int)W, (int)ncls, fh, fw, k, margin);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  if (workspace == nullptr || ws + p.total > static_cast<uint8_t*>(workspace) + workspace_bytes) {
    set_error("isb_region_select: workspace too small (need %zu bytes, got %zu)", p.total + 1024, workspace_bytes);
    return ISB_ERR_WORKSPACE;
  }
  ISB_CHECK_ARG(p.pool_smem <= 200 * 1024 && p.sel_smem <= 200 * 1024,
                "isb_region_select: feature map too large for the shared-memory tiles (H*W=%lld)",
                (long long)(H * W));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint16_t* P = reinterpret_cast<uint16_t*>(ws + p.off_P);
  float* e_part = reinterpret_cast<float*>(ws + p.off_epart);
  float* screen = reinterpret_cast<float*>(ws + p.off_screen);
  const int64_t M = B * p.nwin;

  if (p.ldp != C) ISB_CUDA(cudaMemsetAsync(P, 0, static_cast<size_t>(M) * p.ldp * 2, st));
  dim3 pgrid(p.nblk, static_cast<unsigned>(B));
  if (p.CB == 32) {
    ISB_CUDA(cudaFuncSetAttribute(region_pool_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.pool_smem));
    region_pool_kernel<32><<<pgrid, 256, p.pool_smem, st>>>(x, (int)C, (int)H, (int)W, fh, fw, P, (int)p.ldp, e_part);
  } else {
    ISB_CUDA(cudaFuncSetAttribute(region_pool_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.pool_smem));
    region_pool_kernel<16><<<pgrid, 256, p.pool_smem, st>>>(x, (int)C, (int)H, (int)W, fh, fw, P, (int)p.ldp, e_part);
  }
  ISB_CUDA(cudaGetLastError());

  // window classifier + class-max:  screen[m] = max_j (P[m,:] . Wc[j,:] + bc[j])
  CUtensorMap ta, tb;
  rc = make_tmap_bf16_k64(&ta, P, M, C, p.ldp, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_k64(&tb, cls_w_bf16, ncls, C, ld_w, kBN);
  if (rc) return rc;
  RowSched sched{static_cast<int>((M + kBM - 1) / kBM), static_cast<int>((ncls + kBN - 1) / kBN),
                 static_cast<int>((C + kBK - 1) / kBK)};
  RowMaxEpiParams ep{screen, cls_b, static_cast<int>(M), static_cast<int>(ncls)};
  auto kern = gemm_tc_kernel<RowSched, RowMaxEpilogue>;
  ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  const int sms = device_sm_count();
  kern<<<sched.m_blocks < sms ? sched.m_blocks : sms, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta, tb, kSingleTerm, sched, ep);
  ISB_CUDA(cudaGetLastError());

  ISB_CUDA(cudaFuncSetAttribute(region_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.sel_smem));
  region_select_kernel<<<static_cast<unsigned>(B), kSelThreads, p.sel_smem, st>>>(
      x, (int)C, (int)H, (int)W, fh, fw, cls_w, cls_b, (int)ncls, screen, e_part, p.nblk, k, p.ncand,
      1e-10f, idx, nsel, cls_out, win_norm);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_region_gather(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh,
                                 int fw, int k, const int64_t* idx, const int32_t* nsel,
                                 const float* win_norm, const float* shift, uint16_t* U_hi,
                                 uint16_t* U_lo, int64_t ldu, void* stream) {
  ISB_CHECK_ARG(x && idx && nsel && win_norm && shift && U_hi, "isb_region_gather: null pointer");
  ISB_CHECK_ARG(B > 0 && C > 0 && fh > 0 && fw > 0 && H >= fh && W >= fw, "isb_region_gather: bad shape");
  ISB_CHECK_ARG(k >= 1 && k <= kSelMaxCand, "isb_region_gather: need 1 <= k <= %d", kSelMaxCand);
  const int64_t Kin = C * fh * fw;
  const int64_t KinP = (Kin + 7) / 8 * 8;
  ISB_CHECK_ARG(Kin < (1ll << 30), "isb_region_gather: C*fh*fw too large");
  ISB_CHECK_ARG(ldu >= KinP && ldu % 8 == 0 && (reinterpret_cast<uintptr_t>(U_hi) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(U_lo) & 15) == 0, "isb_region_gather: bad ldu / alignment");
  dim3 grid(static_cast<unsigned>((KinP / 8 + 255) / 256), static_cast<unsigned>(B));
  region_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, (int)C, (int)H, (int)W, fh, fw, k, idx, nsel, win_norm, shift, U_hi, U_lo, ldu);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_descriptor_finalize(const float* y, int64_t B, int64_t D, const float* bias,
                                       const int32_t* nsel, float eps, float* desc, void* stream) {
  ISB_CHECK_ARG(y && desc, "isb_descriptor_finalize: null pointer");
  ISB_CHECK_ARG(B >= 0 && D > 0, "isb_descriptor_finalize: bad shape");
  if (B == 0) return ISB_OK;
  descriptor_finalize_kernel<<<static_cast<unsigned>(B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, bias, nsel, (int)D, eps, desc);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}
