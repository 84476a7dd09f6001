// Region-descriptor aggregation of RegionDescriptorNet.forward_single
// (model/siamese.py:185-223) for a whole batch of trunk feature maps.
//
//   1. region_pool_kernel       x[B,C,H,W] fp32 -> window means P[B*H'W', C] as bf16 hi + lo
//                               (AvgPool2d(fh x fw, stride 1), :164-166,187) and
//                               per-pixel channel energy partials (for the crop norms)
//   2. gemm_tc_kernel<RowMax>   P_hi . Wc_hi^T + bc  (the 1x1-conv classifier, :188) with
//                               the class-max (:191) fused into the TMEM epilogue: the screen
//   3. region_candidates_kernel top-(k+margin) windows of every image by the screen; their
//                               pooled rows gathered as split operands
//      isb_gemm_nt_split        fp32-grade logits of the candidates (three tcgen05 products)
//      region_finalize_select   class-max, final top-k (:194), cls_out (:216), ||crop||_2 of
//                               the chosen windows, completeness certificate of the screen
//   4. region_gather_kernel     u[b] = sum_i crop_i / ||crop_i|| + nsel * shift
//                               (NormalizeL2 + Shift, :218-219, summed BEFORE the
//                               projection -- the Linear is linear) -> bf16 hi (+ lo)
//   5. isb_gemm_nt[_split]      u . W^T   (nn.Linear(100352, D), :180)
//   6. descriptor_finalize      desc = l2norm(y + nsel * bias)   (:220-222)
#include <cstdlib>
#include "isb_host.cuh"
#include "isb_gemm_core.cuh"
#include "isb_topk.cuh"

namespace isb {

__device__ __forceinline__ uint16_t bf16_rn(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
__device__ __forceinline__ float bf16_f(uint16_t h) { return __uint_as_float(static_cast<uint32_t>(h) << 16); }

// ------------------------------------------------------------------ staging
// Channel planes of one image are contiguous in NCHW, so a block of planes is ONE
// contiguous byte range: it is staged into shared memory with a single 1-D bulk
// copy through the TMA engine (cp.async.bulk, SASS UBLKCP) completing on an
// mbarrier, double-buffered against the arithmetic on the previous block.  When
// the range is not 16-byte aligned / sized (odd H*W with an odd channel count) the
// same bytes are fetched with plain coalesced loads instead.
constexpr int kPoolThreads = 512;
constexpr int kMaxWinPerThread = 6;   // windows per thread in the horizontal pass (nwin <= 6 * 512)
constexpr int kEParts = 4;   // energy partial planes per CTA (summed in a fixed order)
constexpr int kPoolMaxStages = 6;
constexpr int kPoolDefaultStages = 3;

struct PoolParams {
  const float* x;
  int C, H, W, fh, fw;
  int CB;        // channels per block (multiple of 4)
  int G;         // channel blocks per CTA (one energy partial per CTA)
  int stages;    // plane buffers of the fast kernel's prefetch ring
  int nblk;      // ceil(C / CB)
  int ngroups;   // ceil(nblk / G)
  uint16_t* P_hi;   // [B * nwin, ldp]  bf16(mean)
  uint16_t* P_lo;   // [B * nwin, ldp]  bf16(mean - hi)
  int ldp;
  float* e_part;    // [B][ngroups][H*W]
};

__device__ __forceinline__ void stage_planes(float* dst, const float* src, int n_floats, uint64_t* bar,
                                             bool bulk) {
  // called by every thread; returns after the copy has been ISSUED (bulk) or DONE (plain)
  if (bulk) {
    if (threadIdx.x == 0) {
      ptx::fence_proxy_async();  // earlier generic-proxy writes to dst are ordered before the async write
      // one copy instruction moves at most 2^20 - 16 bytes: split larger ranges
      const uint32_t bytes = static_cast<uint32_t>(n_floats) * 4u;
      ptx::mbar_arrive_expect_tx(bar, bytes);
      uint32_t off = 0;
      while (off < bytes) {
        const uint32_t chunk = (bytes - off) > 65536u ? 65536u : (bytes - off);
        ptx::bulk_load_1d(reinterpret_cast<uint8_t*>(dst) + off, reinterpret_cast<const uint8_t*>(src) + off,
                          chunk, bar);
        off += chunk;
      }
    }
  } else {
    for (int i = threadIdx.x; i < n_floats; i += blockDim.x) dst[i] = __ldg(src + i);
  }
}

// ------------------------------------------------------------------ 1. pooling
// One CTA = one image x G blocks of CB channels.  Per block, out of shared memory:
//   A  per-pixel energy  sum_c x^2                       (for the crop norms)
//   B  vertical fh-sums, in place (one thread per column, fh-register ring)
//   C  horizontal fw-sums -> window mean -> bf16 hi / lo  into a [window][channel]
//      staging tile (odd word stride: conflict-free for lanes that differ in window)
//   D  staging tile -> P_hi / P_lo rows, CB consecutive channels per window
//      (the K-major operand layout of the classifier GEMM), 4-byte coalesced stores
template <int FH>
__global__ void __launch_bounds__(kPoolThreads, 1)
region_pool_generic_kernel(const PoolParams p) {
  extern __shared__ __align__(128) uint8_t pool_smem_raw[];
  const int HW = p.H * p.W, Wo = p.W - p.fw + 1, Ho = p.H - p.fh + 1, nwin = Ho * Wo;
  const int CB = p.CB;
  const int plane_floats = CB * HW;
  float* X0 = reinterpret_cast<float*>(pool_smem_raw);
  float* X1 = X0 + plane_floats;
  float* E = X1 + plane_floats;                                   // [kEParts][HW]
  uint32_t* O = reinterpret_cast<uint32_t*>(E + kEParts * HW);    // [nwin][CB + 1] words
  uint64_t* bars = reinterpret_cast<uint64_t*>(O + static_cast<size_t>(nwin) * (CB + 1) + ((nwin * (CB + 1)) & 1));
  uint16_t* O16 = reinterpret_cast<uint16_t*>(O);
  const int ostride16 = 2 * (CB + 1);

  const int b = blockIdx.y, grp = blockIdx.x, tid = threadIdx.x;
  const int blk0 = grp * p.G;
  const int blk1 = min(blk0 + p.G, p.nblk);
  const float* xb = p.x + static_cast<size_t>(b) * p.C * HW;

  for (int i = tid; i < kEParts * HW; i += kPoolThreads) E[i] = 0.f;
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();

  auto block_src = [&](int blk, int& cvalid) -> const float* {
    const int c0 = blk * CB;
    cvalid = min(CB, p.C - c0);
    return xb + static_cast<size_t>(c0) * HW;
  };
  auto can_bulk = [&](const float* src, int n) -> bool {
    return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n & 3) == 0);
  };

  int cv;
  const float* src = block_src(blk0, cv);
  bool bulk_cur = can_bulk(src, cv * HW);
  stage_planes(X0, src, cv * HW, &bars[0], bulk_cur);
  uint32_t ph0 = 0u, ph1 = 0u;   // mbarrier phase parity of each buffer

  const float inv_area = 1.f / static_cast<float>(p.fh * p.fw);
  const int eparts = min(kEParts, max(1, kPoolThreads / HW));
  // pass C thread grid: TW threads along the windows (a multiple of 32), TC channel slices
  const int TW = min(kPoolThreads, (nwin + 31) & ~31);
  const int TC = kPoolThreads / TW;
  const int tw = tid % TW, tc = (tid / TW < TC) ? tid / TW : CB;   // threads beyond TW * TC idle in pass C
  const int JW = (nwin + TW - 1) / TW;
  int win_in[kMaxWinPerThread];
#pragma unroll
  for (int j = 0; j < kMaxWinPerThread; ++j) {
    const int win = tw + j * TW;
    win_in[j] = (j < JW && win < nwin) ? (win / Wo) * p.W + (win % Wo) : -1;
  }
  int lgCB = 0;
  while ((1 << lgCB) < CB) ++lgCB;

  for (int blk = blk0; blk < blk1; ++blk) {
    const int s = (blk - blk0) & 1;
    float* X = s ? X1 : X0;
    const int c0 = blk * CB;
    const int cvalid = min(CB, p.C - c0);
    // prefetch the next block into the other buffer (free since the barrier that ended iteration blk-1)
    bool bulk_next = false;
    if (blk + 1 < blk1) {
      int cvn;
      const float* nsrc = block_src(blk + 1, cvn);
      bulk_next = can_bulk(nsrc, cvn * HW);
      if (bulk_next) stage_planes(s ? X0 : X1, nsrc, cvn * HW, &bars[s ^ 1], true);
    }
    if (bulk_cur) {
      ptx::mbar_wait(&bars[s], s ? ph1 : ph0);
      if (s) ph1 ^= 1u; else ph0 ^= 1u;
    }
    for (int i = cvalid * HW + tid; i < plane_floats; i += kPoolThreads) X[i] = 0.f;   // ragged last block
    __syncthreads();

    // ---- A: per-pixel energy of these channels
    {
      const int cb_per = (cvalid + eparts - 1) / eparts;
      for (int i = tid; i < eparts * HW; i += kPoolThreads) {
        const int part = i / HW, px = i - part * HW;
        const int cb_lo = part * cb_per, cb_hi = min(cvalid, cb_lo + cb_per);
        float acc = 0.f;
#pragma unroll 4
        for (int cb = cb_lo; cb < cb_hi; ++cb) {
          const float v = X[cb * HW + px];
          acc = fmaf(v, v, acc);
        }
        E[part * HW + px] += acc;
      }
    }
    __syncthreads();
    // ---- B: vertical sums, in place: X[cb][ho][w] = sum_dy x[cb][ho + dy][w]
    for (int i = tid; i < CB * p.W; i += kPoolThreads) {
      const int cb = i / p.W, w = i - cb * p.W;
      float* col = X + cb * HW + w;
      if (FH == 7) {
        float r[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) r[j] = col[j * p.W];
        col[0] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + r[6]);
        for (int ho0 = 1; ho0 < Ho; ho0 += 7) {
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            const int ho = ho0 + j;   // (ho - 1) % 7 == j: slot j leaves the window, x[ho + 6] enters
            if (ho < Ho) {
              r[j] = col[(ho + 6) * p.W];
              col[ho * p.W] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + r[6]);
            }
          }
        }
      } else {
        for (int ho = 0; ho < Ho; ++ho) {   // rows < ho are already overwritten, rows >= ho are raw
          float sum = 0.f;
          for (int dy = 0; dy < p.fh; ++dy) sum += col[(ho + dy) * p.W];
          col[ho * p.W] = sum;
        }
      }
    }
    __syncthreads();
    // ---- C: horizontal sums -> mean -> bf16 hi / lo into the staging tile.  Threads are a
    // (window, channel-slice) grid fixed for the whole kernel: no index arithmetic in the loop.
    for (int cb = tc; cb < CB; cb += TC) {
      const float* xp = X + cb * HW;
#pragma unroll
      for (int j = 0; j < kMaxWinPerThread; ++j) {
        if (j < JW && win_in[j] >= 0) {
          const float* r = xp + win_in[j];
          float sum;
          if (FH == 7) {   // fw == 7 (checked on the host for this instantiation)
            sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + r[6]);
          } else {
            sum = 0.f;
            for (int dx = 0; dx < p.fw; ++dx) sum += r[dx];
          }
          const float mean = sum * inv_area;
          const uint16_t hi = bf16_rn(mean);
          uint16_t* o = O16 + (tw + j * TW) * ostride16 + cb;
          o[0] = hi;
          o[CB] = bf16_rn(mean - bf16_f(hi));
        }
      }
    }
    __syncthreads();
    // ---- D: rows of the staging tile -> global (CB/2 words of hi, CB/2 words of lo per window)
    {
      const size_t row0 = static_cast<size_t>(b) * nwin;
      const int half = CB >> 1;
      const int valid_words = (cvalid + 1) >> 1;
      const int wd = tid & (CB - 1);             // CB is a power of two
      const bool lo = wd >= half;
      const int k2 = lo ? wd - half : wd;
      if (k2 < valid_words) {
        uint16_t* base = (lo ? p.P_lo : p.P_hi) + row0 * p.ldp + c0;
        for (int win = tid >> lgCB; win < nwin; win += kPoolThreads >> lgCB)
          reinterpret_cast<uint32_t*>(base + static_cast<size_t>(win) * p.ldp)[k2] = O[win * (CB + 1) + wd];
      }
    }
    // the non-bulk path fetches the next block only now (buffer s^1 is idle, nothing to overlap with)
    if (blk + 1 < blk1 && !bulk_next) {
      int cvn;
      const float* nsrc = block_src(blk + 1, cvn);
      stage_planes(s ? X0 : X1, nsrc, cvn * HW, &bars[s ^ 1], false);
    }
    bulk_cur = bulk_next;
    __syncthreads();
  }
  float* eo = p.e_part + (static_cast<size_t>(b) * p.ngroups + grp) * HW;
  for (int px = tid; px < HW; px += kPoolThreads) {
    float t = 0.f;
    for (int part = 0; part < eparts; ++part) t += E[part * HW + px];
    eo[px] = t;
  }
}

// ------------------------------------------------------------------ 1b. pooling, fast path
// 7 x 7 windows on maps up to 32 x 32 (the reference's ResNet maps: 14 x 14 at 448 px
// input, up to 32 x 32 at 1024 px).  Same staging, far fewer instructions per element:
//   A  per-pixel energy (as above)
//   B  one thread per (channel, column): the column goes to registers, the 7-tap
//      vertical sums slide down it (add the entering row, subtract the leaving one),
//      and are written back IN PLACE transposed, VT[cb][w][ho], with an odd ho-pitch
//      and a per-plane skew that puts plane cb on bank (cb % 32) -- both passes are
//      then bank-conflict free
//   C  one thread per (channel, output row, row segment): 7-tap horizontal sliding
//      sums out of VT, mean -> bf16 hi / lo (cvt.rn.bf16.f32), stored straight to
//      P_hi / P_lo with the 32 lanes of a warp on 32 consecutive channels (full
//      32-byte sectors; no staging tile, no fourth pass)
constexpr int kFastThreads = 1024;   // maps up to 32 x 32
constexpr int kFastThreadsSmall = 256;   // maps up to 16 x 16: two or three CTAs per SM overlap their passes

__device__ __forceinline__ uint16_t cvt_bf16(float f) {
  uint16_t h;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(f));
  return h;
}

struct FastGeom {
  int Hop;    // odd pitch of the ho index in VT
  int Q;      // output rows per warp-channel group: 32 / min(32, CB)
  int nseg;   // row segments per output row
  int segw;   // outputs per segment
};

// thread geometry of the row pass for a (H x W map, CB channels, NT threads) block; nseg == 0: does not fit
__host__ __device__ constexpr FastGeom make_fast_geom(int H, int W, int cb, int nt) {
  const int Ho = H - 6, Wo = W - 6;
  const int Q = 32 / (cb < 32 ? cb : 32);
  const int cpw = 32 / Q;
  const int cgroups = (cb + cpw - 1) / cpw;
  const int hoblks = (Ho + Q - 1) / Q;
  const int warps_per_seg = cgroups * hoblks;
  int nseg = warps_per_seg > nt / 32 ? 0 : (nt / 32) / warps_per_seg;
  if (nseg > (Wo + 1) / 2) nseg = (Wo + 1) / 2;
  return FastGeom{Ho | 1, Q, nseg, nseg > 0 ? (Wo + nseg - 1) / nseg : 0};
}

__device__ __forceinline__ int vt_base(int cb, int HW, int Q) {
  const int target = (cb % (32 / Q)) * Q;                // bank of VT[cb][0][0]
  return cb * HW + ((target - cb * HW) & 31);
}

// HC, WC, CBC != 0: map size and channel block known at compile time (the reference's 14 x 14
// and 32 x 32 maps) -- every index, pitch and loop bound below folds to a constant, which cuts
// the instruction count per plane ~3x (the kernel is issue-bound otherwise: ncu, 83 % issue
// slots busy at 3.1 TB/s).  0: taken from the launch parameters.
template <int HMAX, int NT, int MINB, int HC, int WC, int CBC>
__global__ void __launch_bounds__(NT, MINB)
region_pool_fast_kernel(const PoolParams p, const FastGeom g_rt) {
  extern __shared__ __align__(128) uint8_t pool_smem_raw[];
  constexpr int FH = 7;
  constexpr bool kStatic = HC != 0;
  const int H = kStatic ? HC : p.H, W = kStatic ? WC : p.W;
  const int CB = kStatic ? CBC : p.CB;
  const FastGeom g = kStatic ? make_fast_geom(HC ? HC : 7, WC ? WC : 7, CBC ? CBC : 16, NT) : g_rt;
  const int HW = H * W, Wo = W - 6, Ho = H - 6, nwin = Ho * Wo;
  const int plane_floats = CB * HW;
  const int NST = p.stages;                                       // ring of NST plane buffers (2 .. kPoolMaxStages)
  float* X0 = reinterpret_cast<float*>(pool_smem_raw);
  float* E = X0 + static_cast<size_t>(NST) * plane_floats;        // [kEParts][HW]
  uint64_t* bars = reinterpret_cast<uint64_t*>(E + kEParts * HW + ((kEParts * HW) & 1));

  const int b = blockIdx.y, grp = blockIdx.x, tid = threadIdx.x;
  const int blk0 = grp * p.G;
  const int blk1 = min(blk0 + p.G, p.nblk);
  const float* xb = p.x + static_cast<size_t>(b) * p.C * HW;

  for (int i = tid; i < kEParts * HW; i += NT) E[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();

  // ---- fixed thread roles
  // pass A: four pixels (one float4) x a quarter of the channels per thread when H*W % 4 == 0
  const bool quadA = (HW & 3) == 0 && NT >= (HW >> 2);
  const int HW4 = HW >> 2;
  const int eparts4 = min(kEParts, NT / max(HW4, 1));
  const int partA = tid / max(HW4, 1), qA = tid - partA * HW4;
  // pass B: column (cb_b, w_b)
  const bool colv = tid < CB * W;
  const int cb_b = tid / W, w_b = tid - cb_b * W;
  const int col_in = cb_b * HW + w_b;
  const int col_out = vt_base(colv ? cb_b : 0, HW, g.Q) + w_b * g.Hop;
  // pass C: lane -> (channel, q), warp -> (channel group, ho block, segment)
  const int lane = tid & 31, warp = tid >> 5;
  const int cpw = 32 / g.Q;                       // channels per warp
  const int cgroups = (CB + cpw - 1) / cpw;
  const int hoblks = (Ho + g.Q - 1) / g.Q;
  const int cg = warp % cgroups;
  const int hb = (warp / cgroups) % hoblks;
  const int seg = warp / (cgroups * hoblks);
  const int cb_c = cg * cpw + lane / g.Q;
  const int ho_c = hb * g.Q + (lane % g.Q);
  const int wo0 = seg * g.segw;
  const int wo1 = min(Wo, wo0 + g.segw);
  const bool rowv = seg < g.nseg && cb_c < CB && ho_c < Ho && wo0 < wo1;
  const int row_in = vt_base(cb_c < CB ? cb_c : 0, HW, g.Q) + ho_c;
  const float inv_area = 1.f / 49.f;
  const int eparts = min(kEParts, max(1, NT / HW));
  // rows of P this thread writes: window (ho_c, wo) of image b, channel c0 + cb_c
  const size_t prow = (static_cast<size_t>(b) * nwin + ho_c * Wo) * p.ldp + cb_c;
  const ptrdiff_t lo_delta = p.P_lo - p.P_hi;     // same offsets into both planes

  auto block_src = [&](int blk, int& cvalid) -> const float* {
    const int c0 = blk * CB;
    cvalid = min(CB, p.C - c0);
    return xb + static_cast<size_t>(c0) * HW;
  };
  // Bulk copies need 16-byte aligned, 16-byte sized ranges.  Every block of this CTA starts a
  // whole number of CB-plane blocks after the first, so one test covers them all; otherwise
  // (odd H*W with an unlucky channel count) the CTA takes the plain-load path, unpipelined.
  int cv_last;
  block_src(blk1 - 1, cv_last);
  const bool bulk_all = ((reinterpret_cast<uintptr_t>(xb + static_cast<size_t>(blk0) * plane_floats) & 15) == 0) &&
                        ((plane_floats & 3) == 0) && (((cv_last * HW) & 3) == 0);
  const int nloc = blk1 - blk0;
  if (bulk_all) {
    // prologue: NST - 1 blocks in flight before the first wait
    for (int j = 0; j < NST - 1 && j < nloc; ++j) {
      int cvn;
      const float* nsrc = block_src(blk0 + j, cvn);
      stage_planes(X0 + static_cast<size_t>(j) * plane_floats, nsrc, cvn * HW, &bars[j], true);
    }
  }

  int s = 0;            // ring slot of the block being consumed
  uint32_t par = 0u;    // its mbarrier phase parity
  for (int it = 0; it < nloc; ++it) {
    const int blk = blk0 + it;
    float* X = X0 + static_cast<size_t>(s) * plane_floats;
    const int c0 = blk * CB;
    const int cvalid = min(CB, p.C - c0);
    if (bulk_all) {
      // refill the buffer the previous iteration finished with (every thread passed its last
      // __syncthreads), NST - 1 blocks ahead of the one consumed now
      const int ahead = it + NST - 1;
      if (ahead < nloc) {
        int cvn;
        const float* nsrc = block_src(blk0 + ahead, cvn);
        const int sn = (s == 0) ? NST - 1 : s - 1;
        stage_planes(X0 + static_cast<size_t>(sn) * plane_floats, nsrc, cvn * HW, &bars[sn], true);
      }
      ptx::mbar_wait(&bars[s], par);
    } else {
      int cvn;
      const float* nsrc = block_src(blk, cvn);
      stage_planes(X, nsrc, cvn * HW, &bars[s], false);
      __syncthreads();
    }
    // ---- A: per-pixel energy
    if (quadA) {
      if (partA < eparts4) {
        const int cb_per = (cvalid + eparts4 - 1) / eparts4;
        const int cb_lo = partA * cb_per, cb_hi = min(cvalid, cb_lo + cb_per);
        const float4* xp = reinterpret_cast<const float4*>(X) + qA;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kStatic && cvalid == CB && eparts4 == kEParts) {
          constexpr int kPer = kStatic ? (CBC + kEParts - 1) / kEParts : 1;   // eparts4 == kEParts for the static shapes
#pragma unroll
          for (int j = 0; j < kPer; ++j) {
            const float4 v = xp[(cb_lo + j) * HW4];
            a.x = fmaf(v.x, v.x, a.x); a.y = fmaf(v.y, v.y, a.y);
            a.z = fmaf(v.z, v.z, a.z); a.w = fmaf(v.w, v.w, a.w);
          }
        } else {
          for (int cb = cb_lo; cb < cb_hi; ++cb) {
            const float4 v = xp[cb * HW4];
            a.x = fmaf(v.x, v.x, a.x); a.y = fmaf(v.y, v.y, a.y);
            a.z = fmaf(v.z, v.z, a.z); a.w = fmaf(v.w, v.w, a.w);
          }
        }
        float4* ep = reinterpret_cast<float4*>(E) + partA * HW4 + qA;
        float4 e = *ep;
        e.x += a.x; e.y += a.y; e.z += a.z; e.w += a.w;
        *ep = e;
      }
    } else {
      const int cb_per = (cvalid + eparts - 1) / eparts;
      for (int i = tid; i < eparts * HW; i += NT) {
        const int part = i / HW, px = i - part * HW;
        const int cb_lo = part * cb_per, cb_hi = min(cvalid, cb_lo + cb_per);
        const float* xp = X + px;
        float a0 = 0.f, a1 = 0.f;
        int cb = cb_lo;
        for (; cb + 1 < cb_hi; cb += 2) {
          const float v0 = xp[cb * HW], v1 = xp[(cb + 1) * HW];
          a0 = fmaf(v0, v0, a0);
          a1 = fmaf(v1, v1, a1);
        }
        if (cb < cb_hi) { const float v0 = xp[cb * HW]; a0 = fmaf(v0, v0, a0); }
        E[part * HW + px] += a0 + a1;
      }
    }
    // ---- B: vertical sliding sums down the column (7-register ring), kept in registers
    float o[HMAX - FH + 1];
    if (colv) {
      const float* col = X + col_in;
      float r[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) r[j] = col[j * W];
      float sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + r[6]);
      o[0] = sum;
#pragma unroll
      for (int ho = 1; ho < HMAX - FH + 1; ++ho) {
        if (ho < Ho) {
          const float nv = col[(ho + FH - 1) * W];
          sum += nv - r[(ho - 1) % 7];      // row ho - 1 leaves the window, row ho + 6 enters
          r[(ho - 1) % 7] = nv;
          o[ho] = sum;
        }
      }
    }
    __syncthreads();   // every raw value has been read: the planes may be overwritten
    if (colv) {
#pragma unroll
      for (int ho = 0; ho < HMAX - FH + 1; ++ho)
        if (ho < Ho) X[col_out + ho] = o[ho];
    }
    __syncthreads();
    // ---- C: horizontal sliding sums -> mean -> bf16 hi / lo -> global
    if (rowv && cb_c < cvalid) {
      const float* r = X + row_in + wo0 * g.Hop;
      const int pitch = g.Hop;
      uint16_t* ohi = p.P_hi + prow + c0 + static_cast<size_t>(wo0) * p.ldp;
      float q[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) q[j] = r[j * pitch];
      float sum = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + q[6]);
      {
        const float mean = sum * inv_area;
        const uint16_t hi = cvt_bf16(mean);
        ohi[0] = hi;
        ohi[lo_delta] = cvt_bf16(mean - __uint_as_float(static_cast<uint32_t>(hi) << 16));
      }
      if (kStatic) {
        // segw is a compile-time constant: fully unrolled, offsets are immediates
        constexpr int kSegw = kStatic ? make_fast_geom(HC ? HC : 7, WC ? WC : 7, CBC ? CBC : 16, NT).segw : 1;
        const int nout = wo1 - wo0;
#pragma unroll
        for (int d = 1; d < kSegw; ++d) {
          if (d < nout) {
            const float nv = r[(d + 6) * pitch];
            sum += nv - q[(d - 1) % 7];
            q[(d - 1) % 7] = nv;
            const float mean = sum * inv_area;
            const uint16_t hi = cvt_bf16(mean);
            uint16_t* o2 = ohi + static_cast<size_t>(d) * p.ldp;
            o2[0] = hi;
            o2[lo_delta] = cvt_bf16(mean - __uint_as_float(static_cast<uint32_t>(hi) << 16));
          }
        }
      } else {
        for (int base = wo0 + 1; base < wo1; base += 7) {
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            const int wo = base + j;            // (wo - wo0 - 1) % 7 == j: slot j leaves, column wo + 6 enters
            if (wo < wo1) {
              const float nv = r[(wo - wo0 + 6) * pitch];
              sum += nv - q[j];
              q[j] = nv;
              const float mean = sum * inv_area;
              const uint16_t hi = cvt_bf16(mean);
              uint16_t* o2 = ohi + static_cast<size_t>(wo - wo0) * p.ldp;
              o2[0] = hi;
              o2[lo_delta] = cvt_bf16(mean - __uint_as_float(static_cast<uint32_t>(hi) << 16));
            }
          }
        }
      }
    }
    __syncthreads();
    if (++s == NST) { s = 0; par ^= 1u; }
  }
  float* eo = p.e_part + (static_cast<size_t>(b) * p.ngroups + grp) * HW;
  const int epsum = quadA ? eparts4 : eparts;
  for (int px = tid; px < HW; px += NT) {
    float t = 0.f;
    for (int part = 0; part < epsum; ++part) t += E[part * HW + px];
    eo[px] = t;
  }
}

// ------------------------------------------------------------------ 1c. pooling on the tensor cores
// Maps of up to 256 pixels (14 x 14 at the reference's 448-px input): the window sums are
// one small matrix product per block of 64 channels,
//     S[win, c] = sum_px Box[win, px] * x[c, px],      Box[win, px] = 1 inside the window,
// so the feature map goes HBM -> smem (one bulk copy per 64 channel planes) -> tensor cores
// and no CUDA-core pass ever slides a window.  fp32 inputs are split into bf16 hi + lo
// (x = hi + lo to 2^-16) and both terms accumulate in ONE fp32 TMEM tile; Box is exact in
// bf16.  Persistent CTAs, warp-specialised, three 64 KB buffers that each hold first the raw
// planes and then, converted IN PLACE, the two operand tiles:
//   warp 0       producer    cp.async.bulk of 64 planes into a free buffer
//   warps 6..21  converter   planes -> registers -> K-major 128B-swizzled bf16 hi / lo tiles
//                            (pixels along K); per-pixel channel energy sum_c x^2 on the way
//   warp 1       MMA issuer  tcgen05.mma 128 x 64 x 16: A = Box, B = operand tile, 13 k-steps
//                            per term at 14 x 14, ring of four 64-column accumulators
//   warps 2..5   epilogue    tcgen05.ld (lane = window, 64 channels), mean -> bf16 hi / lo ->
//                            P_hi / P_lo rows (128 contiguous bytes per window and term)
// Box has round8(nwin) real rows; the MMA is issued with M = 128 and the rows beyond read
// whatever follows in shared memory -- their accumulator lanes are never loaded.
constexpr int kTcThreads = 704;
constexpr int kTcCB = 64;       // channels per block = rows of an operand tile = N of the MMA
constexpr int kTcConv = 512;    // converter threads: the conversion is ALU work and needs the issue slots of 16 warps
constexpr int kTcConvWarp0 = 6;
constexpr int kTcBufs = 3;
constexpr int kTcAccBufs = 4;   // accumulator ring
constexpr int kTcMaxItems = 8;  // float4 items a converter thread holds in registers (7 at 14 x 14)
constexpr uint32_t kTcTmemCols = 256;
constexpr uint32_t kTcSlabBytes = kTcCB * 128;   // one 64-pixel slab of an operand tile: 8 KB
constexpr uint32_t kTcBufBytes = 65536;          // raw planes (<= 64 KB), then hi [0,32K) + lo [32K,64K)

struct TcGeom {
  int nslab;     // ceil(H*W / 64)
  int R8;        // rows of Box: nwin rounded up to 8
  int lanes_c;   // converter channel lanes: kTcConv / (H*W/4); energy partial planes per unit
  int n_units;   // B * ngroups
  uint32_t off_buf, off_bars, raw_bytes;
  int dbg;   // development ablation switches (ISB_TC_DEBUG): 1 no MMA, 2 no conversion, 4 no epilogue stores
};

struct TcBars {
  uint64_t raw_full[kTcBufs], op_full[kTcBufs], buf_free[kTcBufs];
  uint64_t acc_full[kTcAccBufs], acc_empty[kTcAccBufs];
  uint32_t tmem_base;
};

// fp32 -> bf16 round-to-nearest (ties away from zero) with integer arithmetic: the bf16 is
// the upper half of the result.  One IADD instead of cvt.rn.bf16x2.f32 (F2FP issues on the
// quarter-rate XU pipe); the tie rule is immaterial here because x - hi is carried by the lo
// term, and |x - hi| <= half an ulp either way.  Finite inputs.
__device__ __forceinline__ uint32_t rn_bf16_hi16(float f) { return __float_as_uint(f) + 0x8000u; }
// two rounded values -> {bf16(b) : bf16(a)}, a at the lower address
__device__ __forceinline__ uint32_t pack_hi16(uint32_t ra, uint32_t rb) { return __byte_perm(ra, rb, 0x7632); }

__device__ __forceinline__ void stg256(void* ptr, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __maxnreg__(88)
region_pool_tc_kernel(const PoolParams p, const TcGeom g) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HW = p.H * p.W, Wo = p.W - p.fw + 1, Ho = p.H - p.fh + 1, nwin = Ho * Wo;
  const int Q4 = HW >> 2;
  uint8_t* box = smem;
  uint8_t* bufs = smem + g.off_buf;
  TcBars* bars = reinterpret_cast<TcBars*>(smem + g.off_bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  // ---- one-time setup: barriers, TMEM, Box (zero elsewhere)
  if (tid == 0) {
    for (int i = 0; i < kTcBufs; ++i) {
      ptx::mbar_init(&bars->raw_full[i], 1);
      ptx::mbar_init(&bars->op_full[i], kTcConv);
      ptx::mbar_init(&bars->buf_free[i], 1);
    }
    for (int i = 0; i < kTcAccBufs; ++i) {
      ptx::mbar_init(&bars->acc_full[i], 1);
      ptx::mbar_init(&bars->acc_empty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTcTmemCols>(&bars->tmem_base);
  for (uint32_t i = tid * 16u; i < g.off_bars; i += kTcThreads * 16u)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < nwin * p.fh * p.fw; i += kTcThreads) {
    const int win = i / (p.fh * p.fw), r = i - win * (p.fh * p.fw);
    const int dy = r / p.fw, dx = r - dy * p.fw;
    const int px = (win / Wo + dy) * p.W + (win % Wo) + dx;
    const int slab = px >> 6, col = px & 63;
    uint8_t* a = box + static_cast<size_t>(slab) * g.R8 * 128 + win * 128 + (((col >> 3) ^ (win & 7)) << 4) +
                 ((col & 7) << 1);
    *reinterpret_cast<uint16_t*>(a) = 0x3F80u;   // bf16 1.0
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int nb_unit = p.G;   // blocks per unit (the last unit of an image may hold fewer)

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
        const int b = u / p.ngroups, grp = u - b * p.ngroups;
        const int blk0 = grp * nb_unit, blk1 = min(blk0 + nb_unit, p.nblk);
        for (int blk = blk0; blk < blk1; ++blk, ++it) {
          const uint32_t q = it % kTcBufs;
          ptx::mbar_wait(&bars->buf_free[q], ((it / kTcBufs) & 1u) ^ 1u);
          const float* src = p.x + (static_cast<size_t>(b) * p.C + static_cast<size_t>(blk) * kTcCB) * HW;
          ptx::mbar_arrive_expect_tx(&bars->raw_full[q], g.raw_bytes);
          uint8_t* dst = bufs + static_cast<size_t>(q) * kTcBufBytes;
          const uint32_t chunk = (g.dbg & 32) ? 65536u : ((g.dbg & 64) ? 4096u : 16384u);
          for (uint32_t off = 0; off < g.raw_bytes; off += chunk) {
            const uint32_t n = min(chunk, g.raw_bytes - off);
            ptx::bulk_load_1d(dst + off, reinterpret_cast<const uint8_t*>(src) + off, n, &bars->raw_full[q]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, kTcCB);
      const uint32_t box_addr = ptx::smem_u32(box), buf_addr = ptx::smem_u32(bufs);
      uint32_t it = 0;
      for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
        const int b = u / p.ngroups, grp = u - b * p.ngroups;
        const int blk0 = grp * nb_unit, blk1 = min(blk0 + nb_unit, p.nblk);
        for (int blk = blk0; blk < blk1; ++blk, ++it) {
          const uint32_t q = it % kTcBufs, d = it % kTcAccBufs;
          ptx::mbar_wait(&bars->op_full[q], (it / kTcBufs) & 1u);
          ptx::mbar_wait(&bars->acc_empty[d], ((it / kTcAccBufs) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t tmem_acc = tmem_base + d * kTcCB;
          uint32_t accumulate = 0u;
          for (int term = 0; term < 2; ++term) {
            for (int slab = 0; slab < g.nslab; ++slab) {
              const int px_left = HW - slab * 64;
              const int ks = px_left >= 64 ? 4 : (px_left + 15) >> 4;
              const uint64_t da = ptx::make_smem_desc_k_sw128(box_addr + static_cast<uint32_t>(slab) * g.R8 * 128u);
              const uint64_t db = ptx::make_smem_desc_k_sw128(buf_addr + q * kTcBufBytes + term * (kTcBufBytes / 2) +
                                                              slab * kTcSlabBytes);
              for (int k = 0; k < ks; ++k) {
                if (!(g.dbg & 1)) ptx::umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc, accumulate);
                accumulate = 1u;
              }
            }
          }
          ptx::umma_commit(&bars->buf_free[q]);
          ptx::umma_commit(&bars->acc_full[d]);
        }
      }
    }
  } else if (warp < kTcConvWarp0) {
    // ------------------------------------------------------------ epilogue
    const int lg = warp & 3;
    const int win = lg * 32 + lane;
    const float inv_area = 1.f / static_cast<float>(p.fh * p.fw);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
      const int b = u / p.ngroups, grp = u - b * p.ngroups;
      const int blk0 = grp * nb_unit, blk1 = min(blk0 + nb_unit, p.nblk);
      for (int blk = blk0; blk < blk1; ++blk, ++it) {
        const uint32_t d = it % kTcAccBufs;
        ptx::mbar_wait(&bars->acc_full[d], (it / kTcAccBufs) & 1u);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + d * kTcCB + (static_cast<uint32_t>(lg * 32) << 16);
        const size_t o = (static_cast<size_t>(b) * nwin + win) * p.ldp + static_cast<size_t>(blk) * kTcCB;
#pragma unroll 1
        for (int hv = 0; hv < 2; ++hv) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr + 32 * hv, v);
          ptx::tmem_ld_wait();
          if (hv == 1) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars->acc_empty[d]);
          }
          if (win < nwin && !(g.dbg & 4)) {
            // 32 channels = 64 bytes per term: two 256-bit stores (whole 32-byte sectors)
            uint16_t* ohi = p.P_hi + o + 32 * hv;
            uint16_t* olo = p.P_lo + o + 32 * hv;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t hw_[8], lw_[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float m0 = __uint_as_float(v[16 * q + 2 * j]) * inv_area;
                const float m1 = __uint_as_float(v[16 * q + 2 * j + 1]) * inv_area;
                const uint32_t r0 = rn_bf16_hi16(m0), r1 = rn_bf16_hi16(m1);
                hw_[j] = pack_hi16(r0, r1);
                lw_[j] = pack_hi16(rn_bf16_hi16(m0 - __uint_as_float(r0 & 0xFFFF0000u)),
                                   rn_bf16_hi16(m1 - __uint_as_float(r1 & 0xFFFF0000u)));
              }
              stg256(ohi + 16 * q, hw_);
              stg256(olo + 16 * q, lw_);
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ converter
    // thread (cl, t): pixel quad t of channels cl, cl + lanes_c, ... of every block
    const int tc = tid - kTcConvWarp0 * 32;
    const int cl = tc / Q4, t = tc - cl * Q4;
    const bool active = cl < g.lanes_c;
    const int n_items = active ? (kTcCB - cl + g.lanes_c - 1) / g.lanes_c : 0;
    // pixel quad t -> slab t / 16, 16-byte chunk (t % 16) / 2, 8-byte half t % 2; the 
    const uint32_t t_off = static_cast<uint32_t>(t >> 4) * kTcSlabBytes + ((t & 1) << 3);
    const uint32_t t_chunk = (t & 15) >> 1;
    const bool tail = active && (t == Q4 - 1) && (HW & 63) != 0;   // this thread also zeroes the K tail
    const uint32_t bufs_addr = ptx::smem_u32(bufs);
    const uint32_t ld_off = (static_cast<uint32_t>(cl) * Q4 + t) * 16u;
    const uint32_t ld_step = static_cast<uint32_t>(g.lanes_c) * Q4 * 16u;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
      const int b = u / p.ngroups, grp = u - b * p.ngroups;
      const int blk0 = grp * nb_unit, blk1 = min(blk0 + nb_unit, p.nblk);
      for (int blk = blk0; blk < blk1; ++blk, ++it) {
        const uint32_t q = it % kTcBufs;
        const uint32_t base = bufs_addr + q * kTcBufBytes;
        ptx::mbar_wait(&bars->raw_full[q], (it / kTcBufs) & 1u);
        float4 x4[kTcMaxItems];
#pragma unroll
        for (int j = 0; j < kTcMaxItems; ++j)
          if (j < n_items && !(g.dbg & 16)) x4[j] = lds128(base + ld_off + j * ld_step);
        ptx::named_bar_sync(1, kTcConv);   // every plane is in registers: the tiles may overwrite them
#pragma unroll
        for (int j = 0; j < kTcMaxItems; ++j) {
          if (j < n_items && !(g.dbg & 2)) {
            const float4 x = x4[j];
            e0 = fmaf(x.x, x.x, e0); e1 = fmaf(x.y, x.y, e1);
            e2 = fmaf(x.z, x.z, e2); e3 = fmaf(x.w, x.w, e3);
            // hi = the upper 16 bits (truncation), lo = bf16(x - hi): x = hi + lo to 2^-16
            const uint32_t u0 = __float_as_uint(x.x), u1 = __float_as_uint(x.y);
            const uint32_t u2 = __float_as_uint(x.z), u3 = __float_as_uint(x.w);
            const uint32_t l0 = rn_bf16_hi16(x.x - __uint_as_float(u0 & 0xFFFF0000u));
            const uint32_t l1 = rn_bf16_hi16(x.y - __uint_as_float(u1 & 0xFFFF0000u));
            const uint32_t l2 = rn_bf16_hi16(x.z - __uint_as_float(u2 & 0xFFFF0000u));
            const uint32_t l3 = rn_bf16_hi16(x.w - __uint_as_float(u3 & 0xFFFF0000u));
            const uint32_t ch = static_cast<uint32_t>(cl + j * g.lanes_c);
            const uint32_t o = base + t_off + ch * 128u + ((t_chunk ^ (ch & 7u)) << 4);
            sts64(o, pack_hi16(u0, u1), pack_hi16(u2, u3));
            sts64(o + kTcBufBytes / 2, pack_hi16(l0, l1), pack_hi16(l2, l3));
            if (tail) {
              // pixels [HW, next multiple of 16) of the last slab are read by the MMA: zero them
              // (the raw planes were lying there)
              const uint32_t rowb = base + static_cast<uint32_t>(t >> 4) * kTcSlabBytes + ch * 128u;
              const int col_end = (((HW & 63) + 15) >> 4) << 4;       // columns the MMA reads in this slab
              for (int col = (HW & 63); col < col_end; col += 4) {  // HW % 4 == 0: whole 8-byte halves
                const uint32_t oz = rowb + ((static_cast<uint32_t>(col >> 3) ^ (ch & 7u)) << 4) + ((col & 4) << 1);
                sts64(oz, 0u, 0u);
                sts64(oz + kTcBufBytes / 2, 0u, 0u);
              }
            }
          }
        }
        ptx::fence_proxy_async();    // generic-proxy tile writes -> visible to the tensor core
        ptx::mbar_arrive(&bars->op_full[q]);
      }
      // unit done: this lane's share of the per-pixel energy (the consumer sums the planes)
      if (active) {
        float* eo = p.e_part + ((static_cast<size_t>(b) * p.ngroups + grp) * g.lanes_c + cl) * HW + 4 * t;
        *reinterpret_cast<float4*>(eo) = make_float4(e0, e1, e2, e3);
        e0 = e1 = e2 = e3 = 0.f;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTcTmemCols>(tmem_base);
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

// The same epilogue, additionally keeping the FOUR best class values of every row and their classes
// (row_top [M][8]: m1 m2 m3 m4 | c1 c2 c3 c4): what region_select_top_kernel needs to re-score a
// window's class-max without the logits of all classes.  Branch-free: the class index (< 512)
// replaces the low 9 mantissa bits of the value (a 2^-14 relative perturbation of a bf16-noise screen
// value, still an ordinary float), and the packed value is inserted into the sorted four by a
// min / max chain -- 7 FMNMX + 1 LOP3 per accumulator value, no divergence (a compare-and-shift
// insertion behind `if (x > m4)` doubled the classifier GEMM's time: its branch is taken by some lane
// of the warp on most columns).  row_max keeps the exact maximum.
constexpr uint32_t kTopIdxBits = 9;                       // ncls <= 512
constexpr uint32_t kTopIdxMask = (1u << kTopIdxBits) - 1u;
constexpr float kTopNone = -3.0e38f;                      // "no class": below any logit

struct RowTopEpiParams {
  float* row_max;      // [M]
  uint32_t* row_top;   // [M][8]
  const float* bias;   // [N]
  int M, N;
};

struct RowTopEpilogue {
  using Params = RowTopEpiParams;
  const Params& p;
  const int row_in_tile;
  float m, m1, m2, m3, m4;
  __device__ RowTopEpilogue(const Params& p_, int r) : p(p_), row_in_tile(r) { m = m1 = m2 = m3 = m4 = 0.f; }
  __device__ __forceinline__ void begin_segment(const Segment&) {
    m = -INFINITY;
    m1 = m2 = m3 = m4 = kTopNone;
  }
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
      const float x = (col < p.N) ? __uint_as_float(v[j]) + __ldg(p.bias + col) : kTopNone;
      m = fmaxf(m, x);
      // fmaxf also maps a NaN logit to "no class": a NaN would otherwise duplicate m1 down the chain
      float t = __uint_as_float((__float_as_uint(fmaxf(x, kTopNone)) & ~kTopIdxMask) | static_cast<uint32_t>(col));
      float n = fmaxf(m1, t); t = fminf(m1, t); m1 = n;
      n = fmaxf(m2, t); t = fminf(m2, t); m2 = n;
      n = fmaxf(m3, t); t = fminf(m3, t); m3 = n;
      m4 = fmaxf(m4, t);
    }
  }
  __device__ __forceinline__ void end_segment(const Segment& seg) {
    const int row = seg.m_block * kBM + row_in_tile;
    if (row < p.M) {
      p.row_max[row] = m;
      const float mm[4] = {m1, m2, m3, m4};
      uint32_t val[4], cls[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool none = mm[i] <= kTopNone * 0.5f;
        val[i] = none ? __float_as_uint(-INFINITY) : (__float_as_uint(mm[i]) & ~kTopIdxMask);
        cls[i] = none ? 0xFFFFFFFFu : (__float_as_uint(mm[i]) & kTopIdxMask);
      }
      uint4* o = reinterpret_cast<uint4*>(p.row_top + static_cast<size_t>(row) * 8);
      o[0] = make_uint4(val[0], val[1], val[2], val[3]);
      o[1] = make_uint4(cls[0], cls[1], cls[2], cls[3]);
    }
  }
};

// one segment = one m-block x ALL n-tiles (the max runs over every class)
struct RowSched {
  ISB_PLAIN_SEGMENT_PASSES
  int m_blocks, n_tiles, k_blocks;
  __device__ __forceinline__ void gate(const Segment&, int, int, int) const {}
  __device__ __forceinline__ void leave(const Segment&) const {}
  __device__ __forceinline__ int num_segments() const { return m_blocks; }
  __device__ __forceinline__ Segment segment(int s) const {
    Segment seg;
    seg.m_block = s; seg.nt_begin = 0; seg.nt_end = n_tiles;
    seg.kb_begin = 0; seg.kb_end = k_blocks; seg.aux = 0; seg.n_tile = 0;
    return seg;
  }
};

// ------------------------------------------------------------------ 3. candidates
constexpr int kSelThreads = 256;
constexpr int kRankSelectMaxWin = 1024;   // maps with more windows select by repeated arg-max
constexpr int kRankSelectSmallWin = 128;  // up to here every window ranks itself against all others
constexpr int kSelMaxCand = 32;   // k + margin

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

// The ncand best windows of one image's screen (value desc, index asc) into cand[] / cand_val[]
// (shared memory, ncand <= kSelMaxCand entries), by all kSelThreads threads of the CTA.
// sc [nwin] holds the screen scores (NaN already mapped to -inf) and is followed by nwin words of
// scratch for the radix keys; on maps of more than kRankSelectMaxWin windows sc is consumed.
__device__ __forceinline__ void select_best_windows(float* sc, int nwin, int ncand, int* cand, float* cand_val,
                                                    float* s_val, int* s_idx) {
  const int tid = threadIdx.x;
  if (nwin <= kRankSelectSmallWin) {
    // rank selection: every window counts the windows that come before it in the order
    // (value desc, index asc); the ncand first write themselves to their slot.  One pass and
    // one barrier instead of ncand block-wide arg-max rounds (27 us per batch at 14 x 14).
    for (int i = tid; i < nwin; i += kSelThreads) {
      const float v = sc[i];
      int rank = 0;
      for (int j = 0; j < nwin; ++j) {
        const float s = sc[j];                      // broadcast read
        rank += (s > v || (s == v && j < i)) ? 1 : 0;
      }
      if (rank < ncand) {
        cand[rank] = i;
        cand_val[rank] = v;
      }
    }
    __syncthreads();
  } else if (nwin <= kRankSelectMaxWin) {
    // larger maps (32 x 32: 676 windows -- the all-pairs ranking above cost 115 us per batch there):
    // the ncand-th largest key by a warp radix select, the windows above it (and the lowest-indexed
    // ties) compacted into a list, and the ranking done inside that list of ncand
    uint32_t* skey = reinterpret_cast<uint32_t*>(sc + ((nwin + 3) & ~3));   // [nwin]
    __shared__ int hist[256];
    __shared__ int list[kSelMaxCand];
    __shared__ uint32_t s_T;
    __shared__ int s_ngt;
    for (int i = tid; i < nwin; i += kSelThreads) skey[i] = f2key(__float_as_uint(sc[i]));
    if (tid == 0) s_ngt = 0;
    __syncthreads();
    if (tid < 32) {
      const uint32_t T = warp_kth_largest(skey, nwin, ncand, hist);
      if (tid == 0) s_T = T;
    }
    __syncthreads();
    const uint32_t T = s_T;
    for (int i = tid; i < nwin; i += kSelThreads)
      if (skey[i] > T) list[atomicAdd(&s_ngt, 1)] = i;      // fewer than ncand by construction
    __syncthreads();
    const int n_gt = s_ngt;
    for (int i = tid; i < nwin; i += kSelThreads) {
      if (skey[i] == T) {
        int before = 0;
        for (int j = 0; j < i; ++j) before += (skey[j] == T) ? 1 : 0;
        if (n_gt + before < ncand) list[n_gt + before] = i;
      }
    }
    __syncthreads();
    if (tid < ncand) {
      const int i = list[tid];
      const float v = sc[i];
      int rank = 0;
      for (int u = 0; u < ncand; ++u) {
        const int j = list[u];
        const float s2 = sc[j];
        rank += (s2 > v || (s2 == v && j < i)) ? 1 : 0;
      }
      cand[rank] = i;
      cand_val[rank] = v;
    }
    __syncthreads();
  } else {
    for (int c = 0; c < ncand; ++c) {
      float v = -INFINITY;
      int vi = 0x7FFFFFFF;
      for (int i = tid; i < nwin; i += kSelThreads) {
        const float s = sc[i];
        if (s > v || (s == v && i < vi)) { v = s; vi = i; }
      }
      float bv; int bi;
      block_argmax(v, vi, s_val, s_idx, bv, bi);
      if (tid == 0) {
        cand[c] = bi;
        cand_val[c] = bv;
        sc[bi] = -INFINITY;
      }
      __syncthreads();
    }
  }
}

// One CTA per image: the ncand best windows of the bf16 screen (value desc, index
// asc), and their pooled rows copied out of P_hi / P_lo into the operand of the
// fp32-grade re-score GEMM: A_hi / A_lo [B * ncand_max, ldp], rows beyond the
// image's candidates zeroed.
__global__ void __launch_bounds__(kSelThreads)
region_candidates_kernel(const float* __restrict__ screen, int nwin, int ncand, int ncand_max,
                         const uint16_t* __restrict__ P_hi, const uint16_t* __restrict__ P_lo, int ldp,
                         int* __restrict__ cand_out, float* __restrict__ cand_screen,
                         uint16_t* __restrict__ A_hi, uint16_t* __restrict__ A_lo) {
  extern __shared__ __align__(16) uint8_t sel_smem_raw[];
  float* sc = reinterpret_cast<float*>(sel_smem_raw);   // [nwin] + [nwin] radix keys
  __shared__ float s_val[kSelThreads / 32];
  __shared__ int s_idx[kSelThreads / 32];
  __shared__ int cand[kSelMaxCand];
  __shared__ float cand_val[kSelMaxCand];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < nwin; i += kSelThreads) {
    const float v = screen[static_cast<size_t>(b) * nwin + i];
    sc[i] = (v == v) ? v : -INFINITY;   // a NaN never outranks anything
  }
  __syncthreads();
  select_best_windows(sc, nwin, ncand, cand, cand_val, s_val, s_idx);
  if (blockIdx.y == 0 && tid < ncand) {
    cand_out[b * ncand_max + tid] = cand[tid];
    cand_screen[b * ncand_max + tid] = cand_val[tid];
  }
  if (tid < ncand_max - ncand) {
    cand_out[b * ncand_max + ncand + tid] = -1;
    cand_screen[b * ncand_max + ncand + tid] = -INFINITY;
  }
  // the candidates' pooled rows -> split operands; the CTAs of an image (gridDim.y of them, each
  // repeats the cheap selection above) copy interleaved 16-byte chunks
  const int vec = ldp / 8;   // 16-byte chunks per row
  for (int i = blockIdx.y * kSelThreads + tid; i < ncand_max * vec; i += kSelThreads * gridDim.y) {
    const int c = i / vec, j = i - c * vec;
    const size_t dst = (static_cast<size_t>(b) * ncand_max + c) * ldp;
    uint4 h = make_uint4(0, 0, 0, 0), l = make_uint4(0, 0, 0, 0);
    if (c < ncand) {
      const size_t srcrow = (static_cast<size_t>(b) * nwin + cand[c]) * ldp;
      h = __ldg(reinterpret_cast<const uint4*>(P_hi + srcrow) + j);
      l = __ldg(reinterpret_cast<const uint4*>(P_lo + srcrow) + j);
    }
    reinterpret_cast<uint4*>(A_hi + dst)[j] = h;
    reinterpret_cast<uint4*>(A_lo + dst)[j] = l;
  }
}

// ------------------------------------------------------------------ 4. final selection
// One CTA per image, from the fp32-grade logits [ncand_max, ldl] of its candidates:
// class-max per candidate (model/siamese.py:191), order (value desc, window asc),
// the k' = min(nwin, k) best (:193-194), cls_out (:216), crop norms, and the
// completeness flag of the screen (see region_select in include/isb.h).
__global__ void __launch_bounds__(kSelThreads)
region_finalize_select_kernel(const float* __restrict__ logits, int ldl, int ncls, const int* __restrict__ cand_in,
                              const float* __restrict__ cand_screen, int nwin, int ncand, int ncand_max,
                              const float* __restrict__ e_part, int ngroups, int H, int W, int fh, int fw,
                              int k, float eps, int64_t* __restrict__ idx, int* __restrict__ nsel_out,
                              float* __restrict__ cls_out, float* __restrict__ win_norm,
                              float* __restrict__ approx_max, float* __restrict__ runner_up,
                              int* __restrict__ n_uncertified) {
  __shared__ int cand[kSelMaxCand];
  __shared__ float cand_max[kSelMaxCand];
  __shared__ int order[kSelMaxCand];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wo = W - fw + 1, HW = H * W;
  const float* lg = logits + static_cast<size_t>(b) * ncand_max * ldl;
  if (tid < ncand) cand[tid] = cand_in[b * ncand_max + tid];
  for (int c = warp; c < ncand; c += kSelThreads / 32) {
    float m = -INFINITY;
    for (int j = lane; j < ncls; j += 32) m = fmaxf(m, lg[static_cast<size_t>(c) * ldl + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) cand_max[c] = m;
  }
  __syncthreads();
  const int nsel = min(nwin, k);
  // order of the <= 32 candidates (class-max desc, window asc): every candidate counts the ones
  // before it and writes itself to that slot
  if (tid < ncand) {
    const float v = cand_max[tid];
    const int w = cand[tid];
    int rank = 0;
    for (int j = 0; j < ncand; ++j) {
      const float s = cand_max[j];
      rank += (s > v || (s == v && cand[j] < w)) ? 1 : 0;
    }
    order[rank] = tid;
  }
  __syncthreads();
  if (tid == 0) {
    nsel_out[b] = nsel;
    if (approx_max != nullptr) {
      for (int i = 0; i < k; ++i) approx_max[static_cast<size_t>(b) * k + i] = (i < nsel) ? cand_max[order[i]] : 0.f;
      runner_up[b] = (ncand > nsel) ? cand_max[order[nsel]] : -INFINITY;
    }
    // Certificate: every window that is not a candidate has a screen value <= t_min (the
    // worst candidate's); with sigma = rms(screen - exact) over the candidates, the k
    // selected windows are provably the top k when  exact_k - t_min > 8 sigma.
    if (nwin > ncand && n_uncertified != nullptr) {
      float t_min = INFINITY, s2 = 0.f;
      for (int c = 0; c < ncand; ++c) {
        const float sv = cand_screen[b * ncand_max + c];
        t_min = fminf(t_min, sv);
        const float d = sv - cand_max[c];
        s2 += d * d;
      }
      const float sigma = sqrtf(s2 / static_cast<float>(ncand));
      const float kth = cand_max[order[nsel - 1]];
      if (!(kth - t_min > 8.f * sigma + 4e-7f * fmaxf(fabsf(kth), fabsf(t_min))))
        n_uncertified[1 + atomicAdd(n_uncertified, 1)] = b;   // count, then the list of images
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += kSelThreads)
    idx[static_cast<size_t>(b) * k + i] = (i < nsel) ? static_cast<int64_t>(cand[order[i]]) : -1;
  // cls_out[b, cls, i]  (zero beyond nsel, model/siamese.py:207-208)
  for (int t = tid; t < ncls * k; t += kSelThreads) {
    const int j = t / k, i = t - j * k;
    cls_out[(static_cast<size_t>(b) * ncls + j) * k + i] =
        (i < nsel) ? lg[static_cast<size_t>(order[i]) * ldl + j] : 0.f;
  }
  // ||crop||: sqrt(sum over the window of the per-pixel energy + eps)
  for (int i = warp; i < k; i += kSelThreads / 32) {
    if (i < nsel) {
      const int win = cand[order[i]];
      const int h = win / Wo, w = win - h * Wo;
      double s = 0.0;
#pragma unroll 8
      for (int t = lane; t < fh * fw * ngroups; t += 32) {   // unrolled: eight loads in flight per lane
        const int g = t / (fh * fw), r = t - g * (fh * fw);
        const int dy = r / fw, dx = r - dy * fw;
        s += static_cast<double>(__ldg(e_part + (static_cast<size_t>(b) * ngroups + g) * HW + (h + dy) * W + w + dx));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) win_norm[static_cast<size_t>(b) * k + i] = sqrtf(static_cast<float>(s) + eps);
    } else if (lane == 0) {
      win_norm[static_cast<size_t>(b) * k + i] = 1.f;
    }
  }
}

// ------------------------------------------------------------------ 4a. candidates + re-score + final selection in one kernel
// What region_candidates_kernel -> isb_gemm_nt_split -> splitk_reduce -> region_finalize_select_kernel
// compute (four launches, 79 us per 256 x 2048 x 14 x 14 batch of which the tensor-core GEMM is 30),
// for the class-max only: the screen GEMM's epilogue (RowTopEpilogue) has kept the four best class
// values of every window and the classes of the best three, so a window's class-max needs the
// fp32-grade logits of THREE classes, not 464 -- 3 * ncand dot products of C per image, done here
// on the CUDA cores from the window's pooled row (hi + lo = the fp32 mean to 2^-17) and the fp32
// weights.  One CTA per image.
//   A class outside the best three has a screen logit <= m4; the window is certified when its
//   re-scored class-max clears m4 by 8 sigma, sigma = max(rms(screen - re-scored) over the image's
//   3 * ncand pairs, half the bf16 noise expected for the rows at hand: 2.34e-3 |p| |w| / sqrt(C)).
//   An image with a window that fails is listed in n_uncertified (second line), like one whose
//   candidate list the completeness certificate rejects.
// cls_out [B, ncls, k] receives the re-scored logits of those classes and -inf elsewhere (isb_region_logits
// prunes on it: only classes within its tolerance of the maximum are scored in true fp32).
constexpr int kTopChunk = 8;   // candidates whose pooled rows sit in shared memory at a time

__global__ void __launch_bounds__(kSelThreads)
region_select_top_kernel(const float* __restrict__ screen, const uint32_t* __restrict__ row_top, int nwin,
                         int ncand, const uint16_t* __restrict__ P_hi, const uint16_t* __restrict__ P_lo, int ldp,
                         int C, const float* __restrict__ cls_w, const float* __restrict__ cls_b, int ncls,
                         const float* __restrict__ e_part, int ngroups, int H, int W, int fh, int fw, int k,
                         float eps, int chunk, int dbg, int64_t* __restrict__ idx, int* __restrict__ nsel_out,
                         float* __restrict__ cls_out, float* __restrict__ win_norm,
                         float* __restrict__ approx_max, float* __restrict__ runner_up,
                         int* __restrict__ n_uncertified) {
  extern __shared__ __align__(16) uint8_t sel_smem_raw[];
  float* sc = reinterpret_cast<float*>(sel_smem_raw);                  // [nwin] + [nwin] radix keys
  float* prow = sc + 2 * ((nwin + 3) & ~3);                            // [chunk][ldp]
  __shared__ float s_val[kSelThreads / 32];
  __shared__ int s_idx[kSelThreads / 32];
  __shared__ int cand[kSelMaxCand];
  __shared__ float cand_val[kSelMaxCand];
  __shared__ float cand_max[kSelMaxCand];
  __shared__ int order[kSelMaxCand];
  __shared__ float t_scr[kSelMaxCand][4];    // m1 .. m4 of the candidate's window (screen)
  __shared__ int t_cls[kSelMaxCand][3];      // classes of m1 .. m3
  __shared__ float t_log[kSelMaxCand][3];    // their re-scored logits
  __shared__ int t_wn2[kSelMaxCand];         // largest squared norm of the weight rows used (float bits)
  __shared__ float t_pn2[kSelMaxCand];       // squared norm of the pooled row
  __shared__ float s_d2;
  __shared__ int s_nv, s_bad;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wo = W - fw + 1, HW = H * W;
  for (int i = tid; i < nwin; i += kSelThreads) {
    const float v = screen[static_cast<size_t>(b) * nwin + i];
    sc[i] = (v == v) ? v : -INFINITY;   // a NaN never outranks anything
  }
  if (tid == 0) { s_d2 = 0.f; s_nv = 0; s_bad = 0; }
  __syncthreads();
  select_best_windows(sc, nwin, ncand, cand, cand_val, s_val, s_idx);
  if (tid < ncand) {
    const uint4* rt = reinterpret_cast<const uint4*>(row_top + (static_cast<size_t>(b) * nwin + cand[tid]) * 8);
    const uint4 a = __ldg(rt), c = __ldg(rt + 1);
    t_scr[tid][0] = __uint_as_float(a.x); t_scr[tid][1] = __uint_as_float(a.y);
    t_scr[tid][2] = __uint_as_float(a.z); t_scr[tid][3] = __uint_as_float(a.w);
    t_cls[tid][0] = static_cast<int>(c.x); t_cls[tid][1] = static_cast<int>(c.y); t_cls[tid][2] = static_cast<int>(c.z);
    t_wn2[tid] = 0;
    t_pn2[tid] = 0.f;
  }
  __syncthreads();
  const int vec = ldp / 8;   // 16-byte chunks of a pooled row (zero padded beyond C)
  const bool w4 = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(cls_w) & 15) == 0;
  for (int c0 = 0; c0 < ((dbg & 1) ? 0 : ncand); c0 += chunk) {
    const int nc = min(chunk, ncand - c0);
    // pooled rows of the chunk: hi + lo -> fp32; four 16-byte loads of each term in flight per thread
    for (int i0 = tid; i0 < nc * vec; i0 += 4 * kSelThreads) {
      uint4 h[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSelThreads;
        if (i < nc * vec) {
          const int c = i / vec, j = i - c * vec;
          const size_t srcrow = (static_cast<size_t>(b) * nwin + cand[c0 + c]) * ldp;
          h[u] = __ldg(reinterpret_cast<const uint4*>(P_hi + srcrow) + j);
          l[u] = __ldg(reinterpret_cast<const uint4*>(P_lo + srcrow) + j);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSelThreads;
        if (i < nc * vec) {
          const int c = i / vec, j = i - c * vec;
          float* d = prow + static_cast<size_t>(c) * ldp + 8 * j;
          const uint32_t hw_[4] = {h[u].x, h[u].y, h[u].z, h[u].w}, lw_[4] = {l[u].x, l[u].y, l[u].z, l[u].w};
          float4 lo4, hi4;
          lo4.x = __uint_as_float(hw_[0] << 16) + __uint_as_float(lw_[0] << 16);
          lo4.y = __uint_as_float(hw_[0] & 0xFFFF0000u) + __uint_as_float(lw_[0] & 0xFFFF0000u);
          lo4.z = __uint_as_float(hw_[1] << 16) + __uint_as_float(lw_[1] << 16);
          lo4.w = __uint_as_float(hw_[1] & 0xFFFF0000u) + __uint_as_float(lw_[1] & 0xFFFF0000u);
          hi4.x = __uint_as_float(hw_[2] << 16) + __uint_as_float(lw_[2] << 16);
          hi4.y = __uint_as_float(hw_[2] & 0xFFFF0000u) + __uint_as_float(lw_[2] & 0xFFFF0000u);
          hi4.z = __uint_as_float(hw_[3] << 16) + __uint_as_float(lw_[3] << 16);
          hi4.w = __uint_as_float(hw_[3] & 0xFFFF0000u) + __uint_as_float(lw_[3] & 0xFFFF0000u);
          reinterpret_cast<float4*>(d)[0] = lo4;
          reinterpret_cast<float4*>(d)[1] = hi4;
        }
      }
    }
    __syncthreads();
    // one warp per candidate: its pooled row against the weight rows of its three classes at once
    // (three independent load streams; the row is read from shared memory once).  Keeping one
    // class's weight row in registers and visiting the candidates that share it was tried and is
    // slower (62 vs 47 us at 14 x 14): on the benchmark's maps an image's 3 * ncand pairs name
    // more than eight distinct classes, so the pooled rows were staged twice or more.
    for (int c = warp; c < nc; c += kSelThreads / 32) {
      const int j0 = t_cls[c0 + c][0], j1 = t_cls[c0 + c][1], j2 = t_cls[c0 + c][2];
      const float* pr = prow + static_cast<size_t>(c) * ldp;
      // a missing class (fewer than three) reads row 0 and is discarded below
      const float* w0 = cls_w + static_cast<size_t>(j0 < 0 ? 0 : j0) * C;
      const float* w1 = cls_w + static_cast<size_t>(j1 < 0 ? 0 : j1) * C;
      const float* w2 = cls_w + static_cast<size_t>(j2 < 0 ? 0 : j2) * C;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, pn = 0.f;
      if (w4) {
#pragma unroll 4
        for (int i = lane; i < C / 4; i += 32) {
          const float4 x0 = __ldg(reinterpret_cast<const float4*>(w0) + i);
          const float4 x1 = __ldg(reinterpret_cast<const float4*>(w1) + i);
          const float4 x2 = __ldg(reinterpret_cast<const float4*>(w2) + i);
          const float4 pv = reinterpret_cast<const float4*>(pr)[i];
          a0 = fmaf(x0.x, pv.x, a0); a0 = fmaf(x0.y, pv.y, a0); a0 = fmaf(x0.z, pv.z, a0); a0 = fmaf(x0.w, pv.w, a0);
          a1 = fmaf(x1.x, pv.x, a1); a1 = fmaf(x1.y, pv.y, a1); a1 = fmaf(x1.z, pv.z, a1); a1 = fmaf(x1.w, pv.w, a1);
          a2 = fmaf(x2.x, pv.x, a2); a2 = fmaf(x2.y, pv.y, a2); a2 = fmaf(x2.z, pv.z, a2); a2 = fmaf(x2.w, pv.w, a2);
          n0 = fmaf(x0.x, x0.x, fmaf(x0.y, x0.y, fmaf(x0.z, x0.z, fmaf(x0.w, x0.w, n0))));
          n1 = fmaf(x1.x, x1.x, fmaf(x1.y, x1.y, fmaf(x1.z, x1.z, fmaf(x1.w, x1.w, n1))));
          n2 = fmaf(x2.x, x2.x, fmaf(x2.y, x2.y, fmaf(x2.z, x2.z, fmaf(x2.w, x2.w, n2))));
          pn = fmaf(pv.x, pv.x, fmaf(pv.y, pv.y, fmaf(pv.z, pv.z, fmaf(pv.w, pv.w, pn))));
        }
      } else {
        for (int i = lane; i < C; i += 32) {
          const float x0 = __ldg(w0 + i), x1 = __ldg(w1 + i), x2 = __ldg(w2 + i), pv = pr[i];
          a0 = fmaf(x0, pv, a0); a1 = fmaf(x1, pv, a1); a2 = fmaf(x2, pv, a2);
          n0 = fmaf(x0, x0, n0); n1 = fmaf(x1, x1, n1); n2 = fmaf(x2, x2, n2);
          pn = fmaf(pv, pv, pn);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o); n0 += __shfl_xor_sync(0xffffffffu, n0, o);
        n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        pn += __shfl_xor_sync(0xffffffffu, pn, o);
      }
      if (lane == 0) {
        t_log[c0 + c][0] = (j0 < 0) ? -INFINITY : a0 + __ldg(cls_b + j0);
        t_log[c0 + c][1] = (j1 < 0) ? -INFINITY : a1 + __ldg(cls_b + j1);
        t_log[c0 + c][2] = (j2 < 0) ? -INFINITY : a2 + __ldg(cls_b + j2);
        t_wn2[c0 + c] = __float_as_int(fmaxf(j0 < 0 ? 0.f : n0, fmaxf(j1 < 0 ? 0.f : n1, j2 < 0 ? 0.f : n2)));
        t_pn2[c0 + c] = pn;
      }
    }
    __syncthreads();
  }
  // class-max of every candidate, screen noise over all re-scored pairs
  if (tid < ncand) {
    float m = -INFINITY, d2 = 0.f;
    int nv = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (t_cls[tid][t] >= 0) {
        const float lg = t_log[tid][t];
        m = fmaxf(m, lg);
        const float d = t_scr[tid][t] - lg;
        d2 += d * d;
        ++nv;
      }
    }
    cand_max[tid] = m;
    atomicAdd(&s_d2, d2);
    atomicAdd(&s_nv, nv);
  }
  __syncthreads();
  if (tid < ncand) {
    // classes outside the best three: screen logit <= m4
    const float sigma_meas = sqrtf(s_d2 / static_cast<float>(s_nv > 0 ? s_nv : 1));
    const float sigma_exp = 2.34e-3f * sqrtf(t_pn2[tid] * __int_as_float(t_wn2[tid]) / static_cast<float>(C));
    const float sigma = fmaxf(sigma_meas, 0.5f * sigma_exp);
    const float m4 = t_scr[tid][3], cm = cand_max[tid];
    if (m4 > -INFINITY && !(cm - m4 > 8.f * sigma + 4e-7f * fmaxf(fabsf(cm), fabsf(m4)))) atomicOr(&s_bad, 1);
  }
  const int nsel = min(nwin, k);
  // order of the <= 32 candidates (class-max desc, window asc): every candidate counts the ones
  // before it and writes itself to that slot
  if (tid < ncand) {
    const float v = cand_max[tid];
    const int w = cand[tid];
    int rank = 0;
    for (int j = 0; j < ncand; ++j) {
      const float s = cand_max[j];
      rank += (s > v || (s == v && cand[j] < w)) ? 1 : 0;
    }
    order[rank] = tid;
  }
  __syncthreads();
  if (tid == 0) {
    nsel_out[b] = nsel;
    if (approx_max != nullptr) {
      for (int i = 0; i < k; ++i) approx_max[static_cast<size_t>(b) * k + i] = (i < nsel) ? cand_max[order[i]] : 0.f;
      runner_up[b] = (ncand > nsel) ? cand_max[order[nsel]] : -INFINITY;
    }
    bool bad = s_bad != 0;
    // completeness of the candidate list (as region_finalize_select_kernel)
    if (nwin > ncand && !bad) {
      float t_min = INFINITY, s2 = 0.f;
      for (int c = 0; c < ncand; ++c) {
        const float sv = cand_val[c];
        t_min = fminf(t_min, sv);
        const float d = sv - cand_max[c];
        s2 += d * d;
      }
      const float sigma = sqrtf(s2 / static_cast<float>(ncand));
      const float kth = cand_max[order[nsel - 1]];
      bad = !(kth - t_min > 8.f * sigma + 4e-7f * fmaxf(fabsf(kth), fabsf(t_min)));
    }
    if (bad && n_uncertified != nullptr) n_uncertified[1 + atomicAdd(n_uncertified, 1)] = b;
  }
  for (int i = tid; i < k; i += kSelThreads)
    idx[static_cast<size_t>(b) * k + i] = (i < nsel) ? static_cast<int64_t>(cand[order[i]]) : -1;
  // cls_out[b, cls, i]: -inf for the classes not re-scored, zero beyond nsel (model/siamese.py:207-208)
  for (int t = tid; t < ((dbg & 4) ? 0 : ncls * k); t += kSelThreads) {
    const int i = t % k;
    cls_out[static_cast<size_t>(b) * ncls * k + t] = (i < nsel) ? -INFINITY : 0.f;
  }
  __syncthreads();
  if (tid < nsel * 3) {
    const int i = tid / 3, t = tid - i * 3, o = order[i];
    const int j = t_cls[o][t];
    if (j >= 0) cls_out[(static_cast<size_t>(b) * ncls + j) * k + i] = t_log[o][t];
  }
  // ||crop||: sqrt(sum over the window of the per-pixel energy + eps)
  for (int i = warp; i < ((dbg & 2) ? 0 : k); i += kSelThreads / 32) {
    if (i < nsel) {
      const int win = cand[order[i]];
      const int h = win / Wo, w = win - h * Wo;
      double s = 0.0;
#pragma unroll 8
      for (int t = lane; t < fh * fw * ngroups; t += 32) {   // unrolled: eight loads in flight per lane
        const int g = t / (fh * fw), r = t - g * (fh * fw);
        const int dy = r / fw, dx = r - dy * fw;
        s += static_cast<double>(__ldg(e_part + (static_cast<size_t>(b) * ngroups + g) * HW + (h + dy) * W + w + dx));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) win_norm[static_cast<size_t>(b) * k + i] = sqrtf(static_cast<float>(s) + eps);
    } else if (lane == 0) {
      win_norm[static_cast<size_t>(b) * k + i] = 1.f;
    }
  }
}

// ------------------------------------------------------------------ 4b. exact selection (second line)
// For images the fast path could not certify: one CTA per image re-scores candidates of the
// screen EXACTLY -- window means and logits from the fp32 inputs with fp64 accumulation -- in
// ROUNDS: the first round scores the 32 best windows of the screen; while the completeness
// certificate fails (a window not yet scored could still belong to the top k: its screen score is
// within 8 sigma of the exact k-th best), the slots of the candidates that fell out of the top k
// are refilled with the next best windows of the screen and scored, until the certificate holds
// or every window of the map has been scored exactly -- so the result is never left
// uncertified (the round-1 second line scored 32 candidates once and could return a list it had
// flagged incomplete).  Slow (each round streams the whole classifier up to four times per
// CTA), used only for the images isb_region_select / isb_region_logits report.
constexpr int kSelChunk = 8;      // candidates re-scored per pass over the classifier weights

__global__ void __launch_bounds__(kSelThreads)
region_select_exact_kernel(const float* __restrict__ x, int C, int H, int W, int fh, int fw,
                     const float* __restrict__ cls_w, const float* __restrict__ cls_b, int ncls,
                     const float* __restrict__ screen,   // [B*HoWo] class-max by the bf16 GEMM
                     const float* __restrict__ e_part, int nblk, int k, int ncand, float eps,
                     int64_t* __restrict__ idx, int* __restrict__ nsel_out,
                     float* __restrict__ cls_out, float* __restrict__ win_norm,
                     int* __restrict__ n_uncertified) {
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
  __shared__ float cand_scr[kSelMaxCand];
  __shared__ int fresh[kSelMaxCand];      // slots to (re)score in this round
  __shared__ int s_nfresh, s_done, s_scored;
  __shared__ float s_s2;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = x + static_cast<size_t>(b) * C * HW;
  const int nsel = min(nwin, k);

  for (int i = tid; i < nwin; i += kSelThreads) sc[i] = screen[static_cast<size_t>(b) * nwin + i];
  if (tid == 0) {
    s_nfresh = ncand;                     // round 1: every slot is fresh
    for (int c = 0; c < ncand; ++c) fresh[c] = c;
    s_done = 0; s_scored = 0; s_s2 = 0.f;
  }
  __syncthreads();
  const double inv_area = 1.0 / static_cast<double>(fh * fw);
  for (;;) {
    const int nfresh = s_nfresh;
    // ---- fill the fresh slots with the best windows of the screen not scored yet
    for (int f = 0; f < nfresh; ++f) {
      float v = -INFINITY;
      int vi = 0x7FFFFFFF;
      for (int i = tid; i < nwin; i += kSelThreads) {
        const float s = sc[i];
        if (s > v || (s == v && i < vi)) { v = s; vi = i; }
      }
      float bv; int bi;
      block_argmax(v, vi, s_val, s_idx, bv, bi);
      if (tid == 0) { cand[fresh[f]] = bi; cand_scr[fresh[f]] = bv; sc[bi] = -INFINITY; }
      __syncthreads();
    }
    // ---- exact logits of the fresh candidates (fp64 accumulation of fp32 products)
    for (int c0 = 0; c0 < nfresh; c0 += kSelChunk) {
      const int nc = min(kSelChunk, nfresh - c0);
      for (int i = tid; i < nc * C; i += kSelThreads) {
        const int ci = i / C, ch = i - ci * C;
        const int win = cand[fresh[c0 + ci]];
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
            logits[fresh[c0 + t] * ncls + j] = static_cast<float>(a + static_cast<double>(__ldg(cls_b + j)));
        }
      }
      __syncthreads();
    }
    // ---- exact class-max of the fresh candidates
    for (int f = warp; f < nfresh; f += kSelThreads / 32) {
      const int c = fresh[f];
      float m = -INFINITY;
      for (int j = lane; j < ncls; j += 32) m = fmaxf(m, logits[c * ncls + j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) cand_max[c] = m;
    }
    __syncthreads();
    // ---- best screen score among the windows not scored yet
    float rem = -INFINITY;
    {
      float v = -INFINITY;
      int vi = 0x7FFFFFFF;
      for (int i = tid; i < nwin; i += kSelThreads) {
        const float s = sc[i];
        if (s > v || (s == v && i < vi)) { v = s; vi = i; }
      }
      int bi;
      block_argmax(v, vi, s_val, s_idx, rem, bi);
    }
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
      // screen noise, measured on everything scored so far
      float s2 = s_s2;
      for (int f = 0; f < nfresh; ++f) {
        const float d = cand_scr[fresh[f]] - cand_max[fresh[f]];
        s2 += d * d;
      }
      s_s2 = s2;
      s_scored += nfresh;
      const float sigma = sqrtf(s2 / static_cast<float>(s_scored));
      const float kth = cand_max[order[nsel - 1]];
      // completeness: no window left (rem = -inf), or the best one left cannot reach the k-th
      const bool complete = rem == -INFINITY ||
                            (kth - rem > 8.f * sigma + 4e-7f * fmaxf(fabsf(kth), fabsf(rem)));
      const int nfree = ncand - nsel;   // slots outside the top k: refilled in the next round
      if (complete) {
        s_done = 1;
      } else if (nfree <= 0) {          // k fills every slot: cannot continue, report the image
        s_done = 1;
        if (n_uncertified != nullptr) n_uncertified[1 + atomicAdd(n_uncertified, 1)] = b;
      } else {
        const int left = nwin - s_scored;                 // > 0 here: rem is a real window's score
        s_nfresh = nfree < left ? nfree : left;
        for (int f = 0; f < s_nfresh; ++f) fresh[f] = order[nsel + f];
      }
    }
    __syncthreads();
    if (s_done) break;
  }
  if (tid == 0) nsel_out[b] = nsel;
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


// ------------------------------------------------------------------ 5. gather
// One CTA = one image x 8 warps x G units of CW channels; every warp streams ITS OWN channels:
// it stages the rows [r0, r1) of its CW planes that the image's selected windows touch (bulk
// copies when aligned: one per unit when the windows span the whole map, else one per plane)
// through its own ring of NST slots and its own mbarriers, NST - 1 units ahead of the
// arithmetic, and never meets the other warps after the per-image set-up (window offsets,
// norms) -- the CTA-wide barriers of the former block-cooperative loop were its top stall.
// A warp produces, 64 consecutive outputs per pass, the elements
//   u[e] = sum_i x[b, c, h_i + dy, w_i + dx] / norm_i + nsel * shift[e],  e = (c, dy, dx)
// of its channels -- one contiguous run of the operand row per unit.
constexpr int kGatherThreads = 256;
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kGatherMaxStages = 4;
constexpr int kGatherDefaultStages = 2;
constexpr int kGatherDefaultG = 4;      // units per warp

constexpr int kGatherRegWin = 8;

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ float lds_f32_off(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}

// The outputs [e_begin, e_end) of the channels ONE WARP has staged.  The warp takes 64 consecutive
// outputs per pass, lane l the elements base + l and base + 32 + l: consecutive lanes read
// consecutive pixels of a window row (at most 2-way bank conflicts) and store 64 contiguous
// bytes per instruction.  NW >= 0: the number of summed windows as a compile-time constant --
// one address add, one ld.shared and one FMA per window, offsets and reciprocal norms in
// registers, explicit 32-bit shared addresses (the generic-pointer form cost ten instructions
// per window: ncu source view, profiles/r01_ncu_gather_*).  NW < 0: any count, out of smem.
template <int NW>
__device__ __forceinline__ void gather_block(const float* X, int c0, int pl, int W, int area, int fw, int Kin,
                                             int e_begin, int e_end, const uint32_t (&r_off4)[kGatherRegWin],
                                             const float (&r_norm)[kGatherRegWin], const int* s_off,
                                             const float* s_norm, int nsel, float fn,
                                             const float* __restrict__ shift, uint16_t* uh, uint16_t* ul) {
  const int lane = threadIdx.x & 31;
  const uint32_t xs = ptx::smem_u32(X);
  for (int base = e_begin + lane; base < e_end; base += 64) {
    float u[2], sh[2];
    uint32_t a[2];
    bool live[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int ee = base + 32 * t;
      live[t] = ee < Kin && ee < e_end;      // beyond: zero padding, or the next block's element (not stored)
      const int el = live[t] ? ee : e_begin;
      sh[t] = __ldg(shift + el);             // issued ahead of the shared-memory loads
      const int c = el / area, r = el - c * area;
      const int dy = r / fw, dx = r - dy * fw;
      a[t] = xs + static_cast<uint32_t>((c - c0) * pl + dy * W + dx) * 4u;
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float acc = 0.f;
      if (NW >= 0) {
#pragma unroll
        for (int i = 0; i < (NW > 0 ? NW : 0); ++i) acc = fmaf(lds_f32(a[t] + r_off4[i]), r_norm[i], acc);
      } else {
        for (int i = 0; i < nsel; ++i) acc = fmaf(lds_f32(a[t] + static_cast<uint32_t>(s_off[i]) * 4u), s_norm[i], acc);
      }
      u[t] = live[t] ? fmaf(fn, sh[t], acc) : 0.f;
    }
    // bf16 hi / lo of both values: one packed convert each (round to nearest even)
    uint32_t hi2, lo2;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(u[1]), "f"(u[0]));
    const float r0f = u[0] - __uint_as_float(hi2 << 16);
    const float r1f = u[1] - __uint_as_float(hi2 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(r1f), "f"(r0f));
    uh[base] = static_cast<uint16_t>(hi2 & 0xFFFFu);
    if (ul != nullptr) ul[base] = static_cast<uint16_t>(lo2 & 0xFFFFu);
    if (base + 32 < e_end) {
      uh[base + 32] = static_cast<uint16_t>(hi2 >> 16);
      if (ul != nullptr) ul[base + 32] = static_cast<uint16_t>(lo2 >> 16);
    }
  }
}

template <int FHW>   // FHW = 7: 7 x 7 window with compile-time index arithmetic; 0: generic
__global__ void __launch_bounds__(kGatherThreads, 4)
region_gather_kernel(const float* __restrict__ x, int C, int H, int W, int fh_, int fw_, int k, int k_sum,
                     int CW, int G, int NST, int row_align, const int* __restrict__ image_list,
                     const int* __restrict__ n_list, const int64_t* __restrict__ idx,
                     const int* __restrict__ nsel_in,
                     const float* __restrict__ win_norm, const float* __restrict__ shift,
                     uint16_t* __restrict__ U_hi, uint16_t* __restrict__ U_lo, int64_t ldu,
                     float* __restrict__ win_mean) {
  extern __shared__ __align__(128) uint8_t gat_smem_raw[];
  __shared__ int s_off[kSelMaxCand];
  __shared__ float s_norm[kSelMaxCand];
  __shared__ int s_r0, s_r1;
  __shared__ __align__(8) uint64_t bars[kGatherWarps][kGatherMaxStages];
  const int fh = FHW ? FHW : fh_, fw = FHW ? FHW : fw_;
  // optional image list (fix-up pass): grid.y covers the whole batch, rows beyond *n_list exit
  if (image_list != nullptr && static_cast<int>(blockIdx.y) >= *n_list) return;
  const int b = (image_list != nullptr) ? image_list[blockIdx.y] : blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wo = W - fw + 1, HW = H * W, area = fh * fw;
  const int nall = nsel_in[b];                 // windows listed for this image (means are taken of all)
  const int nsel = min(nall, k_sum);           // leading windows summed into the operand
  // window rows of this image: every lane of warp 0 reads one index, the range is a warp reduction
  if (tid < 32) {
    int r0 = H, r1 = 0;
    for (int i = tid; i < nall; i += 32) {
      const int win = static_cast<int>(idx[static_cast<size_t>(b) * k + i]);
      const int h = win / Wo;
      r0 = min(r0, h);
      r1 = max(r1, h + fh);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r0 = min(r0, __shfl_xor_sync(0xffffffffu, r0, o));
      r1 = max(r1, __shfl_xor_sync(0xffffffffu, r1, o));
    }
    if (tid == 0) {
      if (nall == 0) { r0 = 0; r1 = 0; }
      r0 = (r0 / row_align) * row_align;                       // keep every plane's range 16-byte aligned
      r1 = min(H, ((r1 + row_align - 1) / row_align) * row_align);
      s_r0 = r0; s_r1 = r1;
      for (int w = 0; w < kGatherWarps; ++w)
        for (int st = 0; st < NST; ++st) ptx::mbar_init(&bars[w][st], 1);
      ptx::fence_barrier_init();
    }
  }
  __syncthreads();
  const int r0 = s_r0, nrows = s_r1 - s_r0;
  const int pl = nrows * W;                                  // floats staged per plane
  if (tid < nall) {
    const int win = static_cast<int>(idx[static_cast<size_t>(b) * k + tid]);
    const int h = win / Wo, w = win - h * Wo;
    s_off[tid] = (h - r0) * W + w;
    // the operand is rounded to bf16 hi + lo (2^-17) below: a reciprocal multiply instead of the
    // reference's division (model/custom_modules.py:56) changes nothing that survives that rounding
    s_norm[tid] = 1.f / win_norm[static_cast<size_t>(b) * k + tid];
  }
  __syncthreads();   // the only CTA-wide meeting points: from here on every warp is on its own

  // this warp's channels: unit u covers [cta_c0 + (u * 8 + warp) * CW, + CW)
  const int cta_c0 = blockIdx.x * (kGatherWarps * CW * G);
  const size_t slot_floats = static_cast<size_t>(CW) * HW;
  float* Xw = reinterpret_cast<float*>(gat_smem_raw) + static_cast<size_t>(warp) * NST * slot_floats;
  const float* xsrc = x + static_cast<size_t>(b) * C * HW + r0 * W;    // + channel * HW
  auto unit_c0 = [&](int u) { return cta_c0 + (u * kGatherWarps + warp) * CW; };
  auto unit_nch = [&](int u) { return max(0, min(CW, C - unit_c0(u))); };
  const bool bulk = ((reinterpret_cast<uintptr_t>(xsrc) & 15) == 0) && ((HW & 3) == 0) && ((pl & 3) == 0) && pl > 0;
  // the warp's copies of unit u into ring slot u % NST (all lanes call it)
  auto issue = [&](int u) {
    const int nch = unit_nch(u);
    if (nch > 0) {
      float* dst = Xw + static_cast<size_t>(u % NST) * slot_floats;
      uint64_t* bar = &bars[warp][u % NST];
      const float* src0 = xsrc + static_cast<size_t>(unit_c0(u)) * HW;
      const uint32_t total = static_cast<uint32_t>(nch) * pl * 4u;
      if (lane == 0) {
        ptx::fence_proxy_async();   // the slot's previous contents were read through the generic proxy
        ptx::mbar_arrive_expect_tx(bar, total);
      }
      __syncwarp();
      if (pl == HW) {
        // whole planes: the unit is ONE contiguous range -> a few large copies
        for (uint32_t off = lane * 8192u; off < total; off += 32u * 8192u)
          ptx::bulk_load_1d(reinterpret_cast<uint8_t*>(dst) + off, reinterpret_cast<const uint8_t*>(src0) + off,
                            min(8192u, total - off), bar);
      } else {
        for (int cb = lane; cb < nch; cb += 32)
          ptx::bulk_load_1d(dst + cb * pl, src0 + static_cast<size_t>(cb) * HW, static_cast<uint32_t>(pl) * 4u, bar);
      }
    }
  };
  if (bulk)
    for (int u = 0; u < NST - 1 && u < G; ++u) issue(u);

  const int Kin = C * area;
  const int KinP = (Kin + 7) & ~7;
  const float fn = static_cast<float>(nsel);
  uint16_t* uh = U_hi + static_cast<size_t>(b) * ldu;
  uint16_t* ul = (U_lo != nullptr) ? U_lo + static_cast<size_t>(b) * ldu : nullptr;
  // the (at most kGatherRegWin) summed windows live in registers; more fall back to smem
  uint32_t r_off4[kGatherRegWin];   // byte offsets of the summed windows inside a staged plane
  float r_norm[kGatherRegWin];
#pragma unroll
  for (int i = 0; i < kGatherRegWin; ++i) {
    r_off4[i] = (i < nsel) ? static_cast<uint32_t>(s_off[i]) * 4u : 0u;
    r_norm[i] = (i < nsel) ? s_norm[i] : 0.f;
  }

  for (int u = 0; u < G; ++u) {
    const int c0 = unit_c0(u);
    const int nch = unit_nch(u);
    if (nch <= 0) break;                          // warp-uniform: the following units start even later
    float* X = Xw + static_cast<size_t>(u % NST) * slot_floats;
    if (bulk) {
      if (u + NST - 1 < G) issue(u + NST - 1);   // into the slot unit u - 1 has released (syncwarp below)
      ptx::mbar_wait(&bars[warp][u % NST], static_cast<uint32_t>((u / NST) & 1));
    } else {
      const float* src0 = xsrc + static_cast<size_t>(c0) * HW;
      for (int i = lane; i < nch * pl; i += 32) {
        const int cb = i / pl, o = i - cb * pl;
        X[i] = __ldg(src0 + static_cast<size_t>(cb) * HW + o);
      }
      __syncwarp();
    }
    // by-product: the exact fp32 mean of every selected window (row-major sum, then / area --
    // AvgPool2d's own arithmetic, model/siamese.py:187), input of isb_region_logits
    if (win_mean != nullptr) {
      const float farea = static_cast<float>(area);
      const uint32_t xs = ptx::smem_u32(X);
      for (int t = lane; t < nch * nall; t += 32) {
        const int i = t / nch, cb = t - i * nch;   // consecutive lanes: consecutive channels of one window
        float sum = 0.f;
        if (FHW == 7) {
          // explicit shared addresses, the seven taps of a row as immediates
          uint32_t ra = xs + static_cast<uint32_t>(cb * pl + s_off[i]) * 4u;
#pragma unroll
          for (int dy = 0; dy < 7; ++dy) {
            sum += lds_f32_off<0>(ra);  sum += lds_f32_off<4>(ra);  sum += lds_f32_off<8>(ra);
            sum += lds_f32_off<12>(ra); sum += lds_f32_off<16>(ra); sum += lds_f32_off<20>(ra);
            sum += lds_f32_off<24>(ra);
            ra += static_cast<uint32_t>(W) * 4u;
          }
        } else {
          const float* pw = X + cb * pl + s_off[i];
          for (int dy = 0; dy < fh; ++dy)
            for (int dx = 0; dx < fw; ++dx) sum += pw[dy * W + dx];
        }
        win_mean[(static_cast<size_t>(b) * k + i) * C + c0 + cb] = sum / farea;
      }
    }
    const int e_begin = c0 * area;
    // the warp that holds the image's last channel also writes the zero padding [Kin, KinP)
    const int e_end = (c0 + nch >= C) ? KinP : (c0 + nch) * area;
    switch (nsel <= kGatherRegWin ? nsel : -1) {
      case 0: gather_block<0>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 1: gather_block<1>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 2: gather_block<2>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 3: gather_block<3>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 4: gather_block<4>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 5: gather_block<5>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 6: gather_block<6>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 7: gather_block<7>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      case 8: gather_block<8>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
      default: gather_block<-1>(X, c0, pl, W, area, fw, Kin, e_begin, e_end, r_off4, r_norm, s_off, s_norm, nsel, fn, shift, uh, ul); break;
    }
    __syncwarp();   // slot u % NST is free for the copy issued at the top of unit u + 1
  }
}

// ------------------------------------------------------------------ 4b. gather, small maps (7 x 7 windows)
// Maps of up to 256 pixels (14 x 14: the reference's 448-px input) with at most 8 listed windows.
// The warp-per-channel-stream kernel above spends ~5 instructions per useful one on such maps
// (117 M warp instructions per 256 x 2048 x 14 x 14 batch, issue-bound at 160 us: its units are 4
// planes, so unit overhead and per-element index arithmetic dominate).  Here a CTA stages 32 whole
// planes per step (ONE bulk copy: they are contiguous in NCHW; ring of NST steps) and every thread
// owns one (channel, dx) column of the 7 x 7 output patch:
//   for each listed window i: the 7 values x[c, h_i + dy, w_i + dx], dy = 0..6 -- ONE address add and
//   seven ld.shared with immediate row offsets; their sum is the thread's share of the window mean,
//   and for the summed windows they go into the seven accumulators u[c, dy, dx] with one FMA each.
// The four channels of a warp are spaced so that its 28 active lanes (4 channels x 7 dx) hit 28
// different banks.  The 32 x 49 outputs of a step are converted to bf16 hi / lo in a shared tile
// and leave as 16-byte vectors (the block is one contiguous, 16-byte aligned range of the row).
constexpr int kG7MaxWin = 8;
constexpr int kG7MaxStages = 6;

// WT = map width as a compile-time constant (14), 0 = runtime; SC = channels per step (16 or 32:
// SC / 4 warps per CTA).  Smaller steps in a deeper ring keep more bytes in flight per SM.
template <int WT, int SC>
__global__ void __launch_bounds__(SC * 8)
region_gather7_kernel(const float* __restrict__ x, int C, int H, int W_, int k, int k_sum, int CPB, int NST, int CM,
                      const int* __restrict__ image_list, const int* __restrict__ n_list,
                      const int64_t* __restrict__ idx, const int* __restrict__ nsel_in,
                      const float* __restrict__ win_norm, const float* __restrict__ shift,
                      uint16_t* __restrict__ U_hi, uint16_t* __restrict__ U_lo, int64_t ldu,
                      float* __restrict__ win_mean) {
  constexpr int kG7CsStride = SC + 4;   // colsum [window][dx][channel]: conflict-free both ways
  extern __shared__ __align__(128) uint8_t g7_smem[];
  __shared__ uint32_t s_off[kG7MaxWin];
  __shared__ float s_inv[kG7MaxWin];
  __shared__ __align__(8) uint64_t full[kG7MaxStages];
  if (image_list != nullptr && static_cast<int>(blockIdx.y) >= *n_list) return;
  const int b = (image_list != nullptr) ? image_list[blockIdx.y] : blockIdx.y;
  const int W = WT ? WT : W_;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W, Wo = W - 6;
  const int nall = min(nsel_in[b], min(k, kG7MaxWin));
  const int nsel = min(nall, k_sum);
  const int S = HW;   // planes of a step are contiguous in NCHW: ONE bulk copy per step
  float* planes = reinterpret_cast<float*>(g7_smem);                            // [NST][SC][HW]
  uint16_t* out_hi = reinterpret_cast<uint16_t*>(planes + static_cast<size_t>(NST) * SC * S);   // [32 * 49]
  uint16_t* out_lo = out_hi + SC * 49;
  float* colsum = reinterpret_cast<float*>(out_lo + SC * 49);                // [8][7][36]
  if (tid < kG7MaxWin) {
    uint32_t off = 0u;
    float inv = 0.f;
    if (tid < nall) {
      const int win = static_cast<int>(idx[static_cast<size_t>(b) * k + tid]);
      const int h = win / Wo, w = win - h * Wo;
      off = static_cast<uint32_t>(h * W + w) * 4u;
      // reciprocal multiply instead of the reference's division (model/custom_modules.py:56): the
      // operand is rounded to bf16 hi + lo (2^-17) below, nothing of the difference survives
      inv = 1.f / win_norm[static_cast<size_t>(b) * k + tid];
    }
    s_off[tid] = off;
    s_inv[tid] = inv;
  }
  if (tid == 0) {
    for (int st = 0; st < NST; ++st) ptx::mbar_init(&full[st], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int c_begin = blockIdx.x * CPB, c_end = min(C, c_begin + CPB);
  const int nstep = (c_end - c_begin + SC - 1) / SC;
  const float* xb = x + static_cast<size_t>(b) * C * HW;
  // thread 0 stages step st into slot st % NST: the SC planes are one contiguous range
  auto issue = [&](int st) {
    const int c0 = c_begin + st * SC;
    const int nch = min(SC, c_end - c0);
    uint64_t* bar = &full[st % NST];
    const uint32_t bytes = static_cast<uint32_t>(nch) * HW * 4u;
    ptx::fence_proxy_async();   // the slot was last read through the generic proxy
    ptx::mbar_arrive_expect_tx(bar, bytes);
    ptx::bulk_load_1d(planes + static_cast<size_t>(st % NST) * SC * S, xb + static_cast<size_t>(c0) * HW, bytes, bar);
  };
  if (tid == 0)
    for (int st = 0; st < NST && st < nstep; ++st) issue(st);

  // this thread's column: dx = lane % 7 of channel cl of the step.  The four channels of a warp are
  // CM apart, CM * HW = 8 or 24 (mod 32) words (14 x 14: CM = 2), so that its 28 active lanes hit 28
  // different banks although the planes are not padded; the warps tile the SC channels in groups
  // of 4 * CM
  const int c_sub = lane / 7, dx = lane - c_sub * 7;
  const int cl = (warp / CM) * (4 * CM) + (warp % CM) + CM * c_sub;
  uint32_t off[kG7MaxWin];
  float inv[kG7MaxWin];
#pragma unroll
  for (int i = 0; i < kG7MaxWin; ++i) { off[i] = s_off[i]; inv[i] = s_inv[i]; }
  const float fn = static_cast<float>(nsel);
  const uint32_t RS = static_cast<uint32_t>(W) * 4u;
  uint16_t* uh = U_hi + static_cast<size_t>(b) * ldu;
  uint16_t* ul = (U_lo != nullptr) ? U_lo + static_cast<size_t>(b) * ldu : nullptr;
  const int Kin = C * 49;

  for (int st = 0; st < nstep; ++st) {
    const int c0 = c_begin + st * SC;
    const int nch = min(SC, c_end - c0);
    ptx::mbar_wait(&full[st % NST], static_cast<uint32_t>((st / NST) & 1));
    if (lane < 28 && cl < nch) {
      const uint32_t base = ptx::smem_u32(planes + (static_cast<size_t>(st % NST) * SC + cl) * S) + dx * 4u;
      float acc[7];
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) acc[dy] = 0.f;
      const float* sh = shift + static_cast<size_t>(c0 + cl) * 49 + dx;
      float shv[7];
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) shv[dy] = __ldg(sh + dy * 7);   // in flight under the shared-memory loads
#pragma unroll
      for (int i = 0; i < kG7MaxWin; ++i) {
        if (i < nall) {   // CTA-uniform
          const uint32_t a = base + off[i];
          float v[7];
          if (WT) {
            v[0] = lds_f32_off<0>(a);           v[1] = lds_f32_off<4 * WT>(a);
            v[2] = lds_f32_off<8 * WT>(a);      v[3] = lds_f32_off<12 * WT>(a);
            v[4] = lds_f32_off<16 * WT>(a);     v[5] = lds_f32_off<20 * WT>(a);
            v[6] = lds_f32_off<24 * WT>(a);
          } else {
#pragma unroll
            for (int dy = 0; dy < 7; ++dy) v[dy] = lds_f32(a + dy * RS);
          }
          if (win_mean != nullptr)
            colsum[(i * 7 + dx) * kG7CsStride + cl] = (((((v[0] + v[1]) + v[2]) + v[3]) + v[4]) + v[5]) + v[6];
          if (i < nsel) {
#pragma unroll
            for (int dy = 0; dy < 7; ++dy) acc[dy] = fmaf(v[dy], inv[i], acc[dy]);
          }
        }
      }
      // + nsel * shift (Shift, model/custom_modules.py:16-18, once per summed window), then bf16 hi / lo
      const int e0 = cl * 49 + dx;
#pragma unroll
      for (int dy = 0; dy < 7; dy += 2) {
        const float u0 = fmaf(fn, shv[dy], acc[dy]);
        const float u1 = (dy + 1 < 7) ? fmaf(fn, shv[dy + 1 < 7 ? dy + 1 : dy], acc[dy + 1 < 7 ? dy + 1 : dy]) : 0.f;
        uint32_t hi2, lo2;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(u1), "f"(u0));
        const float r0f = u0 - __uint_as_float(hi2 << 16);
        const float r1f = u1 - __uint_as_float(hi2 & 0xFFFF0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(r1f), "f"(r0f));
        out_hi[e0 + dy * 7] = static_cast<uint16_t>(hi2 & 0xFFFFu);
        out_lo[e0 + dy * 7] = static_cast<uint16_t>(lo2 & 0xFFFFu);
        if (dy + 1 < 7) {
          out_hi[e0 + (dy + 1) * 7] = static_cast<uint16_t>(hi2 >> 16);
          out_lo[e0 + (dy + 1) * 7] = static_cast<uint16_t>(lo2 >> 16);
        }
      }
    }
    __syncthreads();   // tile and column sums complete; nobody reads the slot any more
    if (tid == 0 && st + NST < nstep) issue(st + NST);
    // the step's 49 * nch outputs: one contiguous range of the row, 16-byte aligned (c0 % 32 == 0)
    {
      const int n_el = nch * 49;
      const size_t g0 = static_cast<size_t>(c0) * 49;
      const int nvec = n_el >> 3;
      for (int v = tid; v < nvec; v += (SC * 8)) {
        *reinterpret_cast<uint4*>(uh + g0 + 8 * v) = *reinterpret_cast<const uint4*>(out_hi + 8 * v);
        if (ul != nullptr) *reinterpret_cast<uint4*>(ul + g0 + 8 * v) = *reinterpret_cast<const uint4*>(out_lo + 8 * v);
      }
      for (int e = (nvec << 3) + tid; e < n_el; e += (SC * 8)) {
        uh[g0 + e] = out_hi[e];
        if (ul != nullptr) ul[g0 + e] = out_lo[e];
      }
      // the CTA that holds the image's last channel also writes the zero padding [Kin, KinP)
      if (c0 + nch >= C && tid < ((Kin + 7) & ~7) - Kin) {
        uh[Kin + tid] = 0;
        if (ul != nullptr) ul[Kin + tid] = 0;
      }
    }
    // by-product: the fp32 mean of every listed window (AvgPool2d, model/siamese.py:187), the input of
    // isb_region_logits: column sums (dy ascending) added left to right, / 49
    if (win_mean != nullptr) {
      for (int t = tid; t < nch * nall; t += (SC * 8)) {
        const int i = t / nch, c = t - i * nch;   // consecutive lanes: consecutive channels of one window
        const float* cs = colsum + (i * 7) * kG7CsStride + c;
        float sum = cs[0];
#pragma unroll
        for (int d = 1; d < 7; ++d) sum += cs[d * kG7CsStride];
        win_mean[(static_cast<size_t>(b) * k + i) * C + c0 + c] = sum / 49.f;
      }
    }
    __syncthreads();   // tile and column sums are free for the next step
  }
}

// ------------------------------------------------------------------ 4c. gather, larger maps (7 x 7 windows)
// Maps of 257 .. 1024 pixels with W % 4 == 0 (32 x 32: the 1024-px input) and at most 8 listed
// windows.  A plane is up to 4 KB here, so a step stages the ROW RANGE that covers the image's
// windows for 8 channels (one bulk copy per channel), one channel per warp.  The channel-stream kernel
// gives such a map ONE channel per warp-unit: 49 outputs and 8 window means per unit, i.e. passes with
// 8 of 32 lanes busy -- 286 M warp instructions per 256 x 2048 x 32 x 32 batch, bound by the L1 /
// shared-memory pipe (83 %) at 36 % of the DRAM throughput (profiles/r02_ncu_regions_32x32.txt).
// Here a lane owns the (dx, window pair) column slice of its warp's channel: lane = 4 dx + wg handles
// windows wg and wg + 4 -- seven ld.shared with immediate row offsets per window, their sum is the
// lane's share of that window's mean -- and the seven accumulators u[c, dy, dx] are completed by two
// xor-shuffles over wg.  (The terms of a sum are therefore added pairwise, not in window order as the
// other two kernels do: equal to fp32 rounding, not bit for bit.)
constexpr int kGLThreads = 256;
constexpr int kGLSC = 8;          // channels per step = warps per CTA
constexpr int kGLCsStride = 12;   // colsum [window][dx][channel]

template <int WT>   // WT = map width as a compile-time constant (32), 0 = runtime
__global__ void __launch_bounds__(kGLThreads)
region_gatherL_kernel(const float* __restrict__ x, int C, int H, int W_, int k, int k_sum, int CPB, int NST,
                      const int* __restrict__ image_list, const int* __restrict__ n_list,
                      const int64_t* __restrict__ idx, const int* __restrict__ nsel_in,
                      const float* __restrict__ win_norm, const float* __restrict__ shift,
                      uint16_t* __restrict__ U_hi, uint16_t* __restrict__ U_lo, int64_t ldu,
                      float* __restrict__ win_mean) {
  extern __shared__ __align__(128) uint8_t gl_smem[];
  __shared__ uint32_t s_off[kG7MaxWin];
  __shared__ float s_inv[kG7MaxWin];
  __shared__ int s_r0, s_r1;
  __shared__ __align__(8) uint64_t full[kG7MaxStages];
  if (image_list != nullptr && static_cast<int>(blockIdx.y) >= *n_list) return;
  const int b = (image_list != nullptr) ? image_list[blockIdx.y] : blockIdx.y;
  const int W = WT ? WT : W_;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W, Wo = W - 6;
  const int nall = min(nsel_in[b], min(k, kG7MaxWin));
  const int nsel = min(nall, k_sum);
  // row range of the image's windows (warp 0), barriers
  if (tid < 32) {
    int r0 = H, r1 = 0;
    if (tid < nall) {
      const int h = static_cast<int>(idx[static_cast<size_t>(b) * k + tid]) / Wo;
      r0 = h;
      r1 = h + 7;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r0 = min(r0, __shfl_xor_sync(0xffffffffu, r0, o));
      r1 = max(r1, __shfl_xor_sync(0xffffffffu, r1, o));
    }
    if (tid == 0) {
      if (nall == 0) { r0 = 0; r1 = 0; }
      s_r0 = r0; s_r1 = r1;
      for (int st = 0; st < NST; ++st) ptx::mbar_init(&full[st], 1);
      ptx::fence_barrier_init();
    }
  }
  __syncthreads();
  const int r0 = s_r0, pl = (s_r1 - s_r0) * W;   // floats staged per plane (a multiple of 4: W % 4 == 0)
  if (tid < kG7MaxWin) {
    uint32_t off = 0u;
    float inv = 0.f;
    if (tid < nall) {
      const int win = static_cast<int>(idx[static_cast<size_t>(b) * k + tid]);
      const int h = win / Wo, w = win - h * Wo;
      off = static_cast<uint32_t>((h - r0) * W + w) * 4u;
      inv = 1.f / win_norm[static_cast<size_t>(b) * k + tid];   // see region_gather7_kernel
    }
    s_off[tid] = off;
    s_inv[tid] = inv;
  }
  __syncthreads();
  float* planes = reinterpret_cast<float*>(gl_smem);                                         // [NST][8][HW]
  uint16_t* out_hi = reinterpret_cast<uint16_t*>(planes + static_cast<size_t>(NST) * kGLSC * HW);   // [8 * 49]
  uint16_t* out_lo = out_hi + kGLSC * 49;
  float* colsum = reinterpret_cast<float*>(out_lo + kGLSC * 49);                              // [8][7][12]
  const int c_begin = blockIdx.x * CPB, c_end = min(C, c_begin + CPB);
  const int nstep = (pl > 0) ? (c_end - c_begin + kGLSC - 1) / kGLSC : 0;
  const float* xb = x + static_cast<size_t>(b) * C * HW + r0 * W;
  // warp 0 stages step st into slot st % NST: the row range of one channel per lane
  auto issue = [&](int st) {
    const int c0 = c_begin + st * kGLSC;
    const int nch = min(kGLSC, c_end - c0);
    uint64_t* bar = &full[st % NST];
    if (lane == 0) {
      ptx::fence_proxy_async();   // the slot was last read through the generic proxy
      ptx::mbar_arrive_expect_tx(bar, static_cast<uint32_t>(nch) * pl * 4u);
    }
    __syncwarp();
    if (lane < nch)
      ptx::bulk_load_1d(planes + (static_cast<size_t>(st % NST) * kGLSC + lane) * HW,
                        xb + static_cast<size_t>(c0 + lane) * HW, static_cast<uint32_t>(pl) * 4u, bar);
  };
  if (warp == 0)
    for (int st = 0; st < NST && st < nstep; ++st) issue(st);

  // lane = 4 dx + wg: column dx of the patch, windows wg and wg + 4; the warp's channel is cl = warp
  const int dx = lane >> 2, wg = lane & 3;
  const bool col = lane < 28;
  const int cl = warp;
  const uint32_t off0 = s_off[wg], off1 = s_off[wg + 4];
  const float inv0 = s_inv[wg], inv1 = s_inv[wg + 4];
  const bool have0 = wg < nall, have1 = wg + 4 < nall, sum0 = wg < nsel, sum1 = wg + 4 < nsel;
  const float fn = static_cast<float>(nsel);
  const uint32_t RS = static_cast<uint32_t>(W) * 4u;
  uint16_t* uh = U_hi + static_cast<size_t>(b) * ldu;
  uint16_t* ul = (U_lo != nullptr) ? U_lo + static_cast<size_t>(b) * ldu : nullptr;
  const int Kin = C * 49;

  for (int st = 0; st < nstep; ++st) {
    const int c0 = c_begin + st * kGLSC;
    const int nch = min(kGLSC, c_end - c0);
    ptx::mbar_wait(&full[st % NST], static_cast<uint32_t>((st / NST) & 1));
    float acc[7];
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) acc[dy] = 0.f;
    const bool act = col && cl < nch;
    float shv[7];
    if (act && wg == 0) {
      const float* sh = shift + static_cast<size_t>(c0 + cl) * 49 + dx;
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) shv[dy] = __ldg(sh + dy * 7);   // in flight under the shared-memory loads
    }
    if (act) {
      const uint32_t base = ptx::smem_u32(planes + (static_cast<size_t>(st % NST) * kGLSC + cl) * HW) + dx * 4u;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t == 0 ? have0 : have1) {
          const uint32_t a = base + (t == 0 ? off0 : off1);
          float v[7];
          if (WT) {
            v[0] = lds_f32_off<0>(a);           v[1] = lds_f32_off<4 * WT>(a);
            v[2] = lds_f32_off<8 * WT>(a);      v[3] = lds_f32_off<12 * WT>(a);
            v[4] = lds_f32_off<16 * WT>(a);     v[5] = lds_f32_off<20 * WT>(a);
            v[6] = lds_f32_off<24 * WT>(a);
          } else {
#pragma unroll
            for (int dy = 0; dy < 7; ++dy) v[dy] = lds_f32(a + dy * RS);
          }
          if (win_mean != nullptr)
            colsum[((wg + 4 * t) * 7 + dx) * kGLCsStride + cl] =
                (((((v[0] + v[1]) + v[2]) + v[3]) + v[4]) + v[5]) + v[6];
          if (t == 0 ? sum0 : sum1) {
            const float iv = (t == 0) ? inv0 : inv1;
#pragma unroll
            for (int dy = 0; dy < 7; ++dy) acc[dy] = fmaf(v[dy], iv, acc[dy]);
          }
        }
      }
    }
    // the four window groups of a column: lanes 4 dx .. 4 dx + 3 (all 32 lanes take part in the shuffles)
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      acc[dy] += __shfl_xor_sync(0xffffffffu, acc[dy], 1);
      acc[dy] += __shfl_xor_sync(0xffffffffu, acc[dy], 2);
    }
    if (act && wg == 0) {
      // + nsel * shift (Shift, model/custom_modules.py:16-18, once per summed window), then bf16 hi / lo
      const int e0 = cl * 49 + dx;
#pragma unroll
      for (int dy = 0; dy < 7; dy += 2) {
        const int d1 = dy + 1 < 7 ? dy + 1 : dy;
        const float u0 = fmaf(fn, shv[dy], acc[dy]);
        const float u1 = (dy + 1 < 7) ? fmaf(fn, shv[d1], acc[d1]) : 0.f;
        uint32_t hi2, lo2;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(u1), "f"(u0));
        const float r0f = u0 - __uint_as_float(hi2 << 16);
        const float r1f = u1 - __uint_as_float(hi2 & 0xFFFF0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(r1f), "f"(r0f));
        out_hi[e0 + dy * 7] = static_cast<uint16_t>(hi2 & 0xFFFFu);
        out_lo[e0 + dy * 7] = static_cast<uint16_t>(lo2 & 0xFFFFu);
        if (dy + 1 < 7) {
          out_hi[e0 + (dy + 1) * 7] = static_cast<uint16_t>(hi2 >> 16);
          out_lo[e0 + (dy + 1) * 7] = static_cast<uint16_t>(lo2 >> 16);
        }
      }
    }
    __syncthreads();   // tile and column sums complete; nobody reads the slot any more
    if (warp == 0 && st + NST < nstep) issue(st + NST);
    {
      const int n_el = nch * 49;                       // 8 channels: 392 elements = 49 16-byte vectors
      const size_t g0 = static_cast<size_t>(c0) * 49;  // c0 % 8 == 0: 16-byte aligned
      const int nvec = n_el >> 3;
      for (int v = tid; v < nvec; v += kGLThreads) {
        *reinterpret_cast<uint4*>(uh + g0 + 8 * v) = *reinterpret_cast<const uint4*>(out_hi + 8 * v);
        if (ul != nullptr) *reinterpret_cast<uint4*>(ul + g0 + 8 * v) = *reinterpret_cast<const uint4*>(out_lo + 8 * v);
      }
      for (int e = (nvec << 3) + tid; e < n_el; e += kGLThreads) {
        uh[g0 + e] = out_hi[e];
        if (ul != nullptr) ul[g0 + e] = out_lo[e];
      }
      if (c0 + nch >= C && tid < ((Kin + 7) & ~7) - Kin) {
        uh[Kin + tid] = 0;
        if (ul != nullptr) ul[Kin + tid] = 0;
      }
    }
    if (win_mean != nullptr) {
      for (int t = tid; t < nch * nall; t += kGLThreads) {
        const int i = t / nch, c = t - i * nch;
        const float* cs = colsum + (i * 7) * kGLCsStride + c;
        float sum = cs[0];
#pragma unroll
        for (int d = 1; d < 7; ++d) sum += cs[d * kGLCsStride];
        win_mean[(static_cast<size_t>(b) * k + i) * C + c0 + c] = sum / 49.f;
      }
    }
    __syncthreads();   // tile and column sums are free for the next step
  }
  // an image without windows (nsel = 0) still owns its operand row: u = 0 * shift
  if (nstep == 0) {
    const size_t g0 = static_cast<size_t>(c_begin) * 49;
    for (int e = tid; e < (c_end - c_begin) * 49; e += kGLThreads) {
      uh[g0 + e] = 0;
      if (ul != nullptr) ul[g0 + e] = 0;
    }
    if (c_end >= C && tid < ((Kin + 7) & ~7) - Kin) {
      uh[Kin + tid] = 0;
      if (ul != nullptr) ul[Kin + tid] = 0;
    }
  }
}

// ------------------------------------------------------------------ 5b. exact logits of the selected windows
// cls_out[b, :, i] = Wc . mean_i + bc in fp32 from the exact window means
// (model/siamese.py:188,216), one CTA per image: the k means sit in shared memory,
// every warp takes classes j, j + 8, ... and forms the k dot products of one
// weight row together.  The k windows are then put in their exact order (class-max
// desc, window asc, :191-194) and checked against the best window that was NOT
// selected: runner_up[b] (its fp32-grade class-max) must stay below the exact k-th
// value by 8 sigma, sigma = rms(fp32-grade - exact) over the k selected.
constexpr int kLogThreads = 256;
constexpr int kLogMaxK = 8;   // windows per pass over the classifier
constexpr int kLogMaxPairs = 256;   // (window, class) contenders evaluated exactly in class-max-only mode

__global__ void __launch_bounds__(kLogThreads)
region_logits_kernel(const float* __restrict__ win_mean, const float* __restrict__ cls_w,
                     const float* __restrict__ cls_b, int C, int ncls, int ke, int k,
                     const int* __restrict__ nsel_in, const float* __restrict__ approx_max,
                     const float* __restrict__ runner_up, const int64_t* __restrict__ idx_in,
                     const float* __restrict__ norm_in, int64_t* __restrict__ idx_out,
                     float* __restrict__ norm_out, int* __restrict__ nsel_out,
                     float* __restrict__ cls_out, const float* __restrict__ approx_cls, float wabs_max,
                     int* __restrict__ changed_list, int* __restrict__ n_changed,
                     int* __restrict__ n_uncertified) {
  extern __shared__ __align__(16) uint8_t log_smem_raw[];
  float* ms = reinterpret_cast<float*>(log_smem_raw);          // [ke][C]
  float* lg = ms + static_cast<size_t>(ke) * C;                 // [ke][ncls]
  __shared__ float wmax[kSelMaxCand];
  __shared__ float tau[kSelMaxCand];
  __shared__ int pair_list[kLogMaxPairs];
  __shared__ int n_pairs;
  __shared__ int order[kSelMaxCand];
  __shared__ int64_t widx[kSelMaxCand];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nall = nsel_in[b];
  const int nsel = min(nall, k);
  for (int i = tid; i < nall * C; i += kLogThreads) ms[i] = win_mean[static_cast<size_t>(b) * ke * C + i];
  if (tid < ke) widx[tid] = idx_in[static_cast<size_t>(b) * ke + tid];
  __syncthreads();
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(cls_w) & 15) == 0;
  // ---- class-max only (eval: the logits themselves are not an output).  A class can
  // hold the exact maximum of window i only if its fp32-grade logit is within tau_i of
  // the fp32-grade maximum, tau_i = twice the worst-case error of the three-product
  // split, 3 * 2^-18 * sum_c |mean_c| * max|w|.  Only those (window, class) pairs are
  // evaluated in true fp32.
  bool pruned = (cls_out == nullptr) && (approx_cls != nullptr);
  if (pruned) {
    if (tid == 0) n_pairs = 0;
    for (int i = warp; i < nall; i += kLogThreads / 32) {
      float sabs = 0.f, am = -INFINITY;
      for (int c = lane; c < C; c += 32) sabs += fabsf(ms[i * C + c]);
      for (int j = lane; j < ncls; j += 32)
        am = fmaxf(am, approx_cls[(static_cast<size_t>(b) * ncls + j) * ke + i]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sabs += __shfl_xor_sync(0xffffffffu, sabs, o);
        am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
      }
      if (lane == 0) { wmax[i] = am; tau[i] = 2.4e-5f * sabs * wabs_max + 1e-6f; }
    }
    __syncthreads();
    for (int t = tid; t < nall * ncls; t += kLogThreads) {
      const int i = t / ncls, j = t - i * ncls;
      if (approx_cls[(static_cast<size_t>(b) * ncls + j) * ke + i] >= wmax[i] - tau[i]) {
        const int pos = atomicAdd(&n_pairs, 1);
        if (pos < kLogMaxPairs) pair_list[pos] = t;
      }
    }
    __syncthreads();
    if (n_pairs > kLogMaxPairs) pruned = false;   // too many contenders: score every class (block-uniform)
  }
  if (pruned) {
    const int np = n_pairs;
    for (int q = tid; q < nall; q += kLogThreads) wmax[q] = -INFINITY;
    __syncthreads();
    for (int q = warp; q < np; q += kLogThreads / 32) {
      const int t = pair_list[q];
      const int i = t / ncls, j = t - i * ncls;
      const float* wr = cls_w + static_cast<size_t>(j) * C;
      float acc = 0.f;
      if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
#pragma unroll 4
        for (int c4 = lane; c4 < C / 4; c4 += 32) {
          const float4 wv = __ldg(w4 + c4);
          const float4 m = reinterpret_cast<const float4*>(ms + i * C)[c4];
          acc = fmaf(wv.x, m.x, acc);
          acc = fmaf(wv.y, m.y, acc);
          acc = fmaf(wv.z, m.z, acc);
          acc = fmaf(wv.w, m.w, acc);
        }
      } else {
        for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(wr + c), ms[i * C + c], acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) lg[q] = acc + __ldg(cls_b + j);
    }
    __syncthreads();
    if (tid == 0) {
      for (int q = 0; q < np; ++q) {
        const int i = pair_list[q] / ncls;
        wmax[i] = fmaxf(wmax[i], lg[q]);
      }
    }
    __syncthreads();
  } else {
  for (int i0 = 0; i0 < nall; i0 += kLogMaxK) {
    const int ni = min(kLogMaxK, nall - i0);
    for (int j = warp; j < ncls; j += kLogThreads / 32) {
      const float* wr = cls_w + static_cast<size_t>(j) * C;
      float acc[kLogMaxK];
#pragma unroll
      for (int t = 0; t < kLogMaxK; ++t) acc[t] = 0.f;
      if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
#pragma unroll 4
        for (int c4 = lane; c4 < C / 4; c4 += 32) {
          const float4 wv = __ldg(w4 + c4);
#pragma unroll
          for (int t = 0; t < kLogMaxK; ++t) {
            if (t < ni) {
              const float4 m = reinterpret_cast<const float4*>(ms + (i0 + t) * C)[c4];
              acc[t] = fmaf(wv.x, m.x, acc[t]);
              acc[t] = fmaf(wv.y, m.y, acc[t]);
              acc[t] = fmaf(wv.z, m.z, acc[t]);
              acc[t] = fmaf(wv.w, m.w, acc[t]);
            }
          }
        }
      } else {
        for (int c = lane; c < C; c += 32) {
          const float wv = __ldg(wr + c);
#pragma unroll
          for (int t = 0; t < kLogMaxK; ++t)
            if (t < ni) acc[t] = fmaf(wv, ms[(i0 + t) * C + c], acc[t]);
        }
      }
      const float bj = __ldg(cls_b + j);
#pragma unroll
      for (int t = 0; t < kLogMaxK; ++t) {
        float a = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && t < ni) lg[(i0 + t) * ncls + j] = a + bj;
      }
    }
  }
  __syncthreads();
  for (int i = warp; i < nall; i += kLogThreads / 32) {
    float m = -INFINITY;
    for (int j = lane; j < ncls; j += 32) m = fmaxf(m, lg[i * ncls + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) wmax[i] = m;
  }
  __syncthreads();
  }
  if (tid == 0) {
    for (int i = 0; i < nall; ++i) order[i] = i;
    for (int i = 1; i < nall; ++i) {
      const int o = order[i];
      int j = i - 1;
      while (j >= 0 && (wmax[order[j]] < wmax[o] || (wmax[order[j]] == wmax[o] && widx[order[j]] > widx[o]))) {
        order[j + 1] = order[j];
        --j;
      }
      order[j + 1] = o;
    }
    nsel_out[b] = nsel;
    // did a runner-up displace one of the k windows the operand was summed over?
    bool changed = false;
    for (int i = 0; i < nsel; ++i) changed = changed || (order[i] >= nsel);
    if (changed && changed_list != nullptr) changed_list[atomicAdd(n_changed, 1)] = b;
    if (n_uncertified != nullptr && nsel > 0) {
      float s2 = 0.f;
      for (int i = 0; i < nall; ++i) {
        const float d = approx_max[static_cast<size_t>(b) * ke + i] - wmax[i];
        s2 += d * d;
      }
      const float sigma = sqrtf(s2 / static_cast<float>(nall));
      const float kth = wmax[order[nsel - 1]];
      const float ru = runner_up[b];   // best fp32-grade class-max among the windows NOT scored here
      if (!(kth - ru > 8.f * sigma + 4e-7f * fabsf(kth))) n_uncertified[1 + atomicAdd(n_uncertified, 1)] = b;
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += kLogThreads) {
    idx_out[static_cast<size_t>(b) * k + i] = (i < nsel) ? widx[order[i]] : -1;
    norm_out[static_cast<size_t>(b) * k + i] = (i < nsel) ? norm_in[static_cast<size_t>(b) * ke + order[i]] : 1.f;
  }
  if (cls_out != nullptr) {
    for (int t = tid; t < ncls * k; t += kLogThreads) {
      const int j = t / k, i = t - j * k;
      cls_out[(static_cast<size_t>(b) * ncls + j) * k + i] = (i < nsel) ? lg[order[i] * ncls + j] : 0.f;
    }
  }
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
  int Ho, Wo, nwin, CB, G, nblk, ngroups, ncand, ncand_max, ldl;
  bool fast;       // region_pool_fast_kernel applies
  FastGeom geom;
  int64_t ldp;
  size_t off_Phi, off_Plo, off_epart, off_screen, off_cand, off_cscreen, off_Ahi, off_Alo, off_logits, off_partials,
      off_rowtop, partials_bytes, total;
  int resc_splits;   // split-K of the candidates' re-score GEMM (its tiles alone do not fill the machine)
  size_t pool_smem, cand_smem;
  int fast_threads;   // block size of region_pool_fast_kernel
  int stages;     // plane buffers of the fast pooling kernel's prefetch ring
  bool tc;        // tensor-core pooling (region_pool_tc_kernel)
  int eplanes;    // energy partial planes per (image, channel group): 1, or the converter lanes of the tc kernel
  TcGeom tcg;
};

static size_t pool_smem_bytes(int CB, int HW, int nwin) {
  size_t words = static_cast<size_t>(nwin) * (CB + 1);
  words += words & 1;
  return static_cast<size_t>(2) * CB * HW * 4 + static_cast<size_t>(kEParts) * HW * 4 + words * 4 + 16;
}

static bool make_region_plan(RegionPlan& p, int64_t B, int C, int H, int W, int ncls, int fh, int fw, int k,
                             int margin) {
  p.Ho = H - fh + 1; p.Wo = W - fw + 1; p.nwin = p.Ho * p.Wo;
  const int HW = H * W;
  p.CB = 0;
  p.fast = false;
  p.stages = 2;
  if (fh == 7 && fw == 7 && H <= 32 && W <= 32 && C % 2 == 0) {
    // fast path: needs CB * W <= 1024 threads for the column pass, CB >= 16 (32-byte
    // sectors on the stores), the skewed transposed planes to fit in place, and enough
    // warps for the row pass
    const int Hop = p.Ho | 1;
    p.fast_threads = (H <= 16 && W <= 16) ? kFastThreadsSmall : kFastThreads;
    // prefetch ring: as many plane buffers (<= want) as leave room for 4 CTAs per SM on small
    // maps (their passes overlap) / fit one SM on large ones
    int want = option(ISB_OPT_POOL_STAGES, kPoolDefaultStages);
    if (want < 2) want = 2;
    if (want > kPoolMaxStages) want = kPoolMaxStages;
    const size_t budget = (p.fast_threads == kFastThreadsSmall) ? 55 * 1024 : 224 * 1024;
    for (int cb : {64, 32, 16}) {
      const size_t fixed = static_cast<size_t>(kEParts) * HW * 4 + 64;
      const size_t stage = static_cast<size_t>(cb) * HW * 4;
      int nst = want;
      while (nst > 2 && nst * stage + fixed > budget) --nst;
      const size_t smem = nst * stage + fixed;
      if (cb * W > p.fast_threads || smem > 224 * 1024 || 31 + W * Hop > HW) continue;
      const FastGeom fg = make_fast_geom(H, W, cb, p.fast_threads);
      if (fg.nseg < 1) continue;
      p.stages = nst;
      p.geom = fg;
      p.pool_smem = smem;
      p.CB = cb;
      p.fast = true;
      break;
    }
  }
  if (!p.fast) {
    for (int cb : {64, 32, 16, 8, 4}) {
      if (pool_smem_bytes(cb, HW, p.nwin) <= 200 * 1024) { p.CB = cb; break; }
    }
    if (p.CB == 0) return false;
    p.pool_smem = pool_smem_bytes(p.CB, HW, p.nwin);
  }
  // tensor-core pooling: maps of <= 256 pixels in whole float4 quads, <= 128 windows, whole
  // 64-channel blocks (any window size: Box is built from fh x fw)
  p.tc = false;
  p.eplanes = 1;
  // Opt-in (ISB_OPT_REGION_POOL_TC = 1) while it is slower than the CUDA-core kernel: 229 us vs 202 us
  // per 256 x 2048 x 14 x 14 batch (profiles/r01_ncu_pool_tc.txt: latency-bound hand-offs
  // between the roles, 29% of DRAM bandwidth); results are identical either way.
  const bool want_tc = option(ISB_OPT_REGION_POOL_TC, 0) == 1;
  if (want_tc && HW <= 256 && HW % 4 == 0 && p.nwin <= 128 && C % kTcCB == 0) {
    TcGeom& t = p.tcg;
    t.nslab = (HW + 63) / 64;
    t.R8 = (p.nwin + 7) & ~7;
    t.lanes_c = kTcConv / (HW / 4);
    if (t.lanes_c > kTcCB) t.lanes_c = kTcCB;
    t.raw_bytes = static_cast<uint32_t>(kTcCB) * HW * 4;
    t.off_buf = static_cast<uint32_t>(t.nslab) * t.R8 * 128;             // Box, then the buffers
    t.off_bars = t.off_buf + kTcBufs * kTcBufBytes;
    const size_t smem = t.off_bars + sizeof(TcBars) + 1024;
    const bool items_ok = (kTcCB + t.lanes_c - 1) / t.lanes_c <= kTcMaxItems;
    if (smem <= 227 * 1024 && items_ok) {
      p.tc = true;
      t.dbg = option(ISB_OPT_TC_DEBUG, 0);
      p.CB = kTcCB;
      p.pool_smem = smem;
      p.eplanes = t.lanes_c;
    }
  }
  p.nblk = (C + p.CB - 1) / p.CB;
  // channel blocks per CTA: enough CTAs to fill the machine a few times over, few
  // enough energy partials (one plane of H*W floats per CTA)
  // the LARGEST power of two that still fills whole waves of resident CTAs to >= 95 % (fewer
  // prologues, energy partials and ring refills per image; measured: 8 at 14 x 14 with four CTAs
  // per SM -- 16 leaves 3.46 waves, 0.64 instead of 0.71 of HBM -- and 32 at 32 x 32 with one CTA
  // per SM, 0.80 instead of 0.76).  ISB_OPT_POOL_G overrides.
  {
    const int64_t slots = static_cast<int64_t>(device_sm_count()) *
                          ((p.fast && p.fast_threads == kFastThreadsSmall) ? 4 : 1);
    p.G = 1;
    for (int g = 2; g <= 128 && g <= p.nblk; g *= 2) {
      const int64_t ctas = static_cast<int64_t>(B) * ((p.nblk + g - 1) / g);
      const int64_t waves = (ctas + slots - 1) / slots;
      if (ctas >= slots && static_cast<double>(ctas) >= 0.95 * static_cast<double>(waves * slots)) p.G = g;
    }
    {
      const int v = option(ISB_OPT_POOL_G, 0);
      if (v >= 1 && v <= 128) p.G = v < p.nblk ? v : p.nblk;
    }
  }
  p.ngroups = (p.nblk + p.G - 1) / p.G;
  if (p.tc) p.tcg.n_units = static_cast<int>(B) * p.ngroups;
  p.ncand_max = k + margin;
  if (p.ncand_max > kSelMaxCand) p.ncand_max = kSelMaxCand;
  p.ncand = p.ncand_max < p.nwin ? p.ncand_max : p.nwin;
  p.ldp = static_cast<int64_t>(align_up(static_cast<size_t>(C), 8));
  p.ldl = static_cast<int>(align_up(static_cast<size_t>(ncls), 4));
  size_t off = 0;
  const size_t Pbytes = static_cast<size_t>(B) * p.nwin * p.ldp * 2;
  const size_t Abytes = static_cast<size_t>(B) * p.ncand_max * p.ldp * 2;
  p.off_Phi = off;     off = align_up(off + Pbytes, 1024);
  p.off_Plo = off;     off = align_up(off + Pbytes, 1024);
  p.off_epart = off;   off = align_up(off + static_cast<size_t>(B) * p.ngroups * p.eplanes * HW * 4, 1024);
  p.off_screen = off;  off = align_up(off + static_cast<size_t>(B) * p.nwin * 4, 1024);
  p.off_cand = off;    off = align_up(off + static_cast<size_t>(B) * p.ncand_max * 4, 1024);
  p.off_cscreen = off; off = align_up(off + static_cast<size_t>(B) * p.ncand_max * 4, 1024);
  p.off_Ahi = off;     off = align_up(off + Abytes, 1024);
  p.off_Alo = off;     off = align_up(off + Abytes, 1024);
  p.off_logits = off;  off = align_up(off + static_cast<size_t>(B) * p.ncand_max * p.ldl * 4, 1024);
  {
    const int64_t Mr = B * p.ncand_max;
    const long long tiles = ((Mr + kBM - 1) / kBM) * ((ncls + kBN - 1) / kBN);
    p.resc_splits = static_cast<int>(148 / (tiles > 0 ? tiles : 1));
    if (p.resc_splits < 1) p.resc_splits = 1;
    if (p.resc_splits > 4) p.resc_splits = 4;
    {
      const int o = option(ISB_OPT_RESC_SPLITS, 0);   // experiment switch (changes the workspace layout)
      if (o >= 1 && o <= 8) p.resc_splits = o;
    }
    p.partials_bytes = p.resc_splits > 1 ? isb_gemm_nt_workspace_bytes(Mr, ncls, C, p.resc_splits) : 0;
  }
  p.off_partials = off; off = align_up(off + p.partials_bytes, 1024);
  p.off_rowtop = off;   off = align_up(off + static_cast<size_t>(B) * p.nwin * 8 * 4, 1024);
  p.total = off;
  p.cand_smem = static_cast<size_t>((p.nwin + 3) & ~3) * 4 * 2;   // scores + radix keys
  return true;
}

}  // namespace isb

using namespace isb;

extern "C" size_t isb_region_select_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W,
                                                    int64_t ncls, int fh, int fw, int k, int margin) {
  if (B <= 0 || C <= 0 || H < fh || W < fw || fh <= 0 || fw <= 0 || k <= 0) return 0;
  RegionPlan p;
  if (!make_region_plan(p, B, (int)C, (int)H, (int)W, (int)ncls, fh, fw, k, margin)) return 0;
  return p.total + 1024;
}

extern "C" int isb_region_select(const float* x, int64_t B, int64_t C, int64_t H, int64_t W,
                                 const float* cls_w, const uint16_t* cls_w_hi, const uint16_t* cls_w_lo,
                                 int64_t ld_w, const float* cls_b, int64_t ncls, int fh, int fw, int k,
                                 int margin, int exact_mode, int64_t* idx, int32_t* nsel, float* cls_out,
                                 float* win_norm, float* approx_max, float* runner_up,
                                 int32_t* n_uncertified, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  ISB_CHECK_ARG(x && cls_w && cls_w_hi && cls_w_lo && cls_b && idx && nsel && cls_out && win_norm,
                "isb_region_select: null pointer");
  ISB_CHECK_ARG((approx_max == nullptr) == (runner_up == nullptr),
                "isb_region_select: approx_max and runner_up go together");
  ISB_CHECK_ARG(B > 0 && C > 0 && ncls > 0 && fh > 0 && fw > 0 && H >= fh && W >= fw,
                "isb_region_select: bad shape (B=%lld C=%lld H=%lld W=%lld window %dx%d)", (long long)B,
                (long long)C, (long long)H, (long long)W, fh, fw);
  ISB_CHECK_ARG(k >= 1 && k <= kSelMaxCand && margin >= 0, "isb_region_select: need 1 <= k <= %d", kSelMaxCand);
  ISB_CHECK_ARG(ld_w >= C && ld_w % 8 == 0, "isb_region_select: bad ld_w");
  ISB_CHECK_ARG(B * (H - fh + 1) * (W - fw + 1) < (1ll << 31) && B * C * H * W < (1ll << 40),
                "isb_region_select: too many windows");
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 3) == 0, "isb_region_select: x must be 4-byte aligned");
  int rc = isb_check_device();
  if (rc) return rc;
  RegionPlan p;
  ISB_CHECK_ARG(make_region_plan(p, B, (int)C, (int)H, (int)W, (int)ncls, fh, fw, k, margin),
                "isb_region_select: feature map too large for the shared-memory tiles (H*W=%lld)",
                (long long)(H * W));
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  if (workspace == nullptr || ws + p.total > static_cast<uint8_t*>(workspace) + workspace_bytes) {
    set_error("isb_region_select: workspace too small (need %zu bytes, got %zu)", p.total + 1024, workspace_bytes);
    return ISB_ERR_WORKSPACE;
  }
  ISB_CHECK_ARG(p.cand_smem <= 200 * 1024, "isb_region_select: too many windows per image (%d)", p.nwin);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint16_t* P_hi = reinterpret_cast<uint16_t*>(ws + p.off_Phi);
  uint16_t* P_lo = reinterpret_cast<uint16_t*>(ws + p.off_Plo);
  float* e_part = reinterpret_cast<float*>(ws + p.off_epart);
  float* screen = reinterpret_cast<float*>(ws + p.off_screen);
  int* cand = reinterpret_cast<int*>(ws + p.off_cand);
  float* cscreen = reinterpret_cast<float*>(ws + p.off_cscreen);
  uint16_t* A_hi = reinterpret_cast<uint16_t*>(ws + p.off_Ahi);
  uint16_t* A_lo = reinterpret_cast<uint16_t*>(ws + p.off_Alo);
  float* logits = reinterpret_cast<float*>(ws + p.off_logits);
  const int64_t M = B * p.nwin;
  if (n_uncertified != nullptr) ISB_CUDA(cudaMemsetAsync(n_uncertified, 0, 4, st));

  // 1. window means (bf16 hi / lo, window-major) + per-pixel energy partials
  PoolParams pp;
  pp.x = x; pp.C = (int)C; pp.H = (int)H; pp.W = (int)W; pp.fh = fh; pp.fw = fw;
  pp.CB = p.CB; pp.G = p.G; pp.stages = p.stages; pp.nblk = p.nblk; pp.ngroups = p.ngroups;
  pp.P_hi = P_hi; pp.P_lo = P_lo; pp.ldp = (int)p.ldp; pp.e_part = e_part;
  dim3 pgrid(p.ngroups, static_cast<unsigned>(B));
  ISB_CHECK_ARG(p.nwin <= kMaxWinPerThread * kPoolThreads, "isb_region_select: too many windows per image (%d)", p.nwin);
  if (p.tc && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    ISB_CUDA(cudaFuncSetAttribute(region_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.pool_smem));
    const int sms = device_sm_count();
    region_pool_tc_kernel<<<p.tcg.n_units < sms ? p.tcg.n_units : sms, kTcThreads, p.pool_smem, st>>>(pp, p.tcg);
    if (p.tcg.dbg & 8) return ISB_OK;   // timing of the pool kernel alone (outputs undefined)
  } else if (p.tc) {
    set_error("isb_region_select: x must be 16-byte aligned for maps of this size");
    return ISB_ERR_INVALID_ARGUMENT;
  } else if (p.fast) {
    // compile-time geometry for the reference's two map sizes, launch-parameter geometry otherwise
    void (*kern)(const PoolParams, const FastGeom);
    if (H == 14 && W == 14 && p.CB == 16) kern = region_pool_fast_kernel<16, kFastThreadsSmall, 4, 14, 14, 16>;
    else if (H == 32 && W == 32 && p.CB == 16) kern = region_pool_fast_kernel<32, kFastThreads, 1, 32, 32, 16>;
    else if (p.fast_threads == kFastThreadsSmall) kern = region_pool_fast_kernel<16, kFastThreadsSmall, 4, 0, 0, 0>;
    else kern = region_pool_fast_kernel<32, kFastThreads, 1, 0, 0, 0>;
    if (option(ISB_OPT_POOL_GENERIC_GEOM, 0) == 1)   // A/B switch: same kernel without the folded constants
      kern = (p.fast_threads == kFastThreadsSmall) ? region_pool_fast_kernel<16, kFastThreadsSmall, 4, 0, 0, 0>
                                                   : region_pool_fast_kernel<32, kFastThreads, 1, 0, 0, 0>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.pool_smem));
    kern<<<pgrid, p.fast_threads, p.pool_smem, st>>>(pp, p.geom);
  } else if (fh == 7 && fw == 7) {
    ISB_CUDA(cudaFuncSetAttribute(region_pool_generic_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.pool_smem));
    region_pool_generic_kernel<7><<<pgrid, kPoolThreads, p.pool_smem, st>>>(pp);
  } else {
    ISB_CUDA(cudaFuncSetAttribute(region_pool_generic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.pool_smem));
    region_pool_generic_kernel<0><<<pgrid, kPoolThreads, p.pool_smem, st>>>(pp);
  }
  ISB_CUDA(cudaGetLastError());

  if (exact_mode < 0) return ISB_OK;   // measurement probe: the pooling pass alone (outputs undefined)

  // 2. window classifier screen + class-max:  screen[m] = max_j (P_hi[m,:] . Wc_hi[j,:] + bc[j])
  CUtensorMap ta, tb;
  rc = make_tmap_bf16_k64(&ta, P_hi, M, C, p.ldp, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_k64(&tb, cls_w_hi, ncls, C, ld_w, kBN);
  if (rc) return rc;
  RowSched sched{static_cast<int>((M + kBM - 1) / kBM), static_cast<int>((ncls + kBN - 1) / kBN),
                 static_cast<int>((C + kBK - 1) / kBK)};
  // option region_top_select (default 1): the epilogue also keeps the four best classes of every window and
  // ONE kernel does candidates + re-score + final selection (region_select_top_kernel); 0: four launches with
  // the re-score of all classes on the tensor cores
  const size_t top_smem_fixed = static_cast<size_t>(2) * ((p.nwin + 3) & ~3) * 4;
  int top_chunk = kTopChunk;
  while (top_chunk > 1 && top_smem_fixed + static_cast<size_t>(top_chunk) * p.ldp * 4 > 160 * 1024) top_chunk >>= 1;
  const bool top_select = !exact_mode && option(ISB_OPT_REGION_TOP_SELECT, 1) != 0 && p.nwin <= kRankSelectMaxWin &&
                          ncls <= (1 << kTopIdxBits) &&
                          top_smem_fixed + static_cast<size_t>(top_chunk) * p.ldp * 4 <= 200 * 1024;
  uint32_t* row_top = reinterpret_cast<uint32_t*>(ws + p.off_rowtop);
  const int sms = device_sm_count();
  if (top_select) {
    RowTopEpiParams ep{screen, row_top, cls_b, static_cast<int>(M), static_cast<int>(ncls)};
    auto kern = gemm_tc_kernel<RowSched, RowTopEpilogue>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    kern<<<sched.m_blocks < sms ? sched.m_blocks : sms, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta, tb, kSingleTerm, sched, ep);
  } else {
    RowMaxEpiParams ep{screen, cls_b, static_cast<int>(M), static_cast<int>(ncls)};
    auto kern = gemm_tc_kernel<RowSched, RowMaxEpilogue>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    kern<<<sched.m_blocks < sms ? sched.m_blocks : sms, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta, tb, kSingleTerm, sched, ep);
  }
  ISB_CUDA(cudaGetLastError());

  if (top_select) {
    const size_t smem = top_smem_fixed + static_cast<size_t>(top_chunk) * p.ldp * 4;
    ISB_CUDA(cudaFuncSetAttribute(region_select_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_select_top_kernel<<<static_cast<unsigned>(B), kSelThreads, smem, st>>>(
        screen, row_top, p.nwin, p.ncand, P_hi, P_lo, (int)p.ldp, (int)C, cls_w, cls_b, (int)ncls, e_part,
        p.ngroups * p.eplanes, (int)H, (int)W, fh, fw, k, 1e-10f, top_chunk, option(ISB_OPT_TC_DEBUG, 0), idx, nsel, cls_out, win_norm, approx_max,
        runner_up, n_uncertified);
    ISB_CUDA(cudaGetLastError());
    return ISB_OK;
  }

  if (exact_mode) {
    // second line: fp64-exact re-score of the candidates straight from the fp32 inputs
    const size_t sel_smem = (static_cast<size_t>((p.nwin + 3) & ~3) + static_cast<size_t>(kSelChunk) * C +
                             static_cast<size_t>(p.ncand) * ncls) * 4;
    ISB_CHECK_ARG(sel_smem <= 200 * 1024, "isb_region_select: exact mode needs %zu bytes of shared memory", sel_smem);
    ISB_CUDA(cudaFuncSetAttribute(region_select_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    region_select_exact_kernel<<<static_cast<unsigned>(B), kSelThreads, sel_smem, st>>>(
        x, (int)C, (int)H, (int)W, fh, fw, cls_w, cls_b, (int)ncls, screen, e_part, p.ngroups * p.eplanes, k, p.ncand,
        1e-10f, idx, nsel, cls_out, win_norm, n_uncertified);
    ISB_CUDA(cudaGetLastError());
    return ISB_OK;
  }

  // 3. candidates of every image + their pooled rows as split operands
  if (p.cand_smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(region_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.cand_smem));
  region_candidates_kernel<<<dim3(static_cast<unsigned>(B), p.nwin <= kRankSelectMaxWin ? 4u : 1u), kSelThreads,
                             p.cand_smem, st>>>(
      screen, p.nwin, p.ncand, p.ncand_max, P_hi, P_lo, (int)p.ldp, cand, cscreen, A_hi, A_lo);
  ISB_CUDA(cudaGetLastError());

  // 4. fp32-grade logits of the candidates: (A_hi, A_lo) . (Wc_hi, Wc_lo)^T + bc
  rc = isb_gemm_nt_split(A_hi, A_lo, p.ldp, cls_w_hi, cls_w_lo, ld_w, B * p.ncand_max, ncls, C, cls_b, logits,
                         p.ldl, p.resc_splits, p.resc_splits > 1 ? ws + p.off_partials : nullptr, p.partials_bytes,
                         stream);
  if (rc) return rc;

  // 5. final order, outputs, crop norms, certificate
  region_finalize_select_kernel<<<static_cast<unsigned>(B), kSelThreads, 0, st>>>(
      logits, p.ldl, (int)ncls, cand, cscreen, p.nwin, p.ncand, p.ncand_max, e_part, p.ngroups * p.eplanes, (int)H, (int)W,
      fh, fw, k, 1e-10f, idx, nsel, cls_out, win_norm, approx_max, runner_up, n_uncertified);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_region_gather(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh,
                                 int fw, int k, int k_sum, const int32_t* image_list,
                                 const int32_t* n_list, const int64_t* idx, const int32_t* nsel,
                                 const float* win_norm, const float* shift, uint16_t* U_hi,
                                 uint16_t* U_lo, int64_t ldu, float* win_mean, void* stream) {
  ISB_CHECK_ARG(k_sum >= 1 && k_sum <= k, "isb_region_gather: need 1 <= k_sum <= k");
  ISB_CHECK_ARG((image_list == nullptr) == (n_list == nullptr), "isb_region_gather: image_list and n_list go together");
  ISB_CHECK_ARG(x && idx && nsel && win_norm && shift && U_hi, "isb_region_gather: null pointer");
  ISB_CHECK_ARG(B > 0 && C > 0 && fh > 0 && fw > 0 && H >= fh && W >= fw, "isb_region_gather: bad shape");
  ISB_CHECK_ARG(k >= 1 && k <= kSelMaxCand, "isb_region_gather: need 1 <= k <= %d", kSelMaxCand);
  const int64_t Kin = C * fh * fw;
  const int64_t KinP = (Kin + 7) / 8 * 8;
  ISB_CHECK_ARG(Kin < (1ll << 30), "isb_region_gather: C*fh*fw too large");
  ISB_CHECK_ARG(ldu >= KinP && ldu % 8 == 0 && (reinterpret_cast<uintptr_t>(U_hi) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(U_lo) & 15) == 0, "isb_region_gather: bad ldu / alignment");
  const int64_t HW = H * W;
  // small maps, 7 x 7 windows, <= 8 listed windows: the plane-block kernel (option gather_small = 0: off)
  if (fh == 7 && fw == 7 && HW <= 256 && HW % 4 == 0 && k <= kG7MaxWin &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && option(ISB_OPT_GATHER_SMALL, 1) != 0) {
    const int SC = option(ISB_OPT_GATHER_CW, 32) == 16 ? 16 : 32;    // channels per step
    // channel spacing inside a warp that spreads its lanes over all banks (see the kernel); 1: none does
    int CM = 1;
    for (int m : {1, 2, 4})
      if ((m * HW) % 32 == 8 || (m * HW) % 32 == 24) { CM = m; break; }
    if (SC % (4 * CM) != 0) CM = 1;
    int NST = option(ISB_OPT_GATHER_STAGES, 2);
    if (NST < 1) NST = 1;
    if (NST > kG7MaxStages) NST = kG7MaxStages;
    // channels per CTA: a multiple of 32, >= 4 waves of CTAs when the batch allows
    int CPB = 256;
    while (CPB > 64 && B * ((C + CPB - 1) / CPB) < 148 * 4 * 4) CPB >>= 1;
    {
      const int g = option(ISB_OPT_GATHER_G, 0);
      if (g >= 1 && g <= 64) CPB = 32 * g;
    }
    const size_t smem = static_cast<size_t>(NST) * SC * HW * 4 + 2 * SC * 49 * 2 + kG7MaxWin * 7 * (SC + 4) * 4;
    dim3 grid(static_cast<unsigned>((C + CPB - 1) / CPB), static_cast<unsigned>(B));
    auto kern = (W == 14) ? (SC == 32 ? region_gather7_kernel<14, 32> : region_gather7_kernel<14, 16>)
                          : (SC == 32 ? region_gather7_kernel<0, 32> : region_gather7_kernel<0, 16>);
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, SC * 8, smem, static_cast<cudaStream_t>(stream)>>>(
        x, (int)C, (int)H, (int)W, k, k_sum, CPB, NST, CM, image_list, n_list, idx, nsel, win_norm, shift, U_hi,
        U_lo, ldu, win_mean);
    ISB_CUDA(cudaGetLastError());
    return ISB_OK;
  }
  // larger maps (up to 1024 pixels, 16-byte rows), <= 8 listed windows: the row-range kernel
  // (option gather_small = 0 switches both plane / row-range kernels off)
  if (fh == 7 && fw == 7 && HW > 256 && HW <= 1024 && W % 4 == 0 && k <= kG7MaxWin &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && option(ISB_OPT_GATHER_SMALL, 1) != 0) {
    int NST = option(ISB_OPT_GATHER_STAGES, 2);
    if (NST < 1) NST = 1;
    if (NST > kG7MaxStages) NST = kG7MaxStages;
    while (NST > 1 && static_cast<size_t>(NST) * kGLSC * HW * 4 > 96 * 1024) --NST;
    int CPB = 128;
    {
      const int g = option(ISB_OPT_GATHER_G, 0);
      if (g >= 1 && g <= 64) CPB = 8 * g;
    }
    const size_t smem = static_cast<size_t>(NST) * kGLSC * HW * 4 + 2 * kGLSC * 49 * 2 + kG7MaxWin * 7 * kGLCsStride * 4;
    dim3 grid(static_cast<unsigned>((C + CPB - 1) / CPB), static_cast<unsigned>(B));
    auto kern = (W == 32) ? region_gatherL_kernel<32> : region_gatherL_kernel<0>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kGLThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        x, (int)C, (int)H, (int)W, k, k_sum, CPB, NST, image_list, n_list, idx, nsel, win_norm, shift, U_hi, U_lo,
        ldu, win_mean);
    ISB_CUDA(cudaGetLastError());
    return ISB_OK;
  }
  // channels per warp unit: <= 4 KB of planes per ring slot (14 x 14: 4 channels, 32 x 32: 1)
  int CW = 1;
  for (int cw : {8, 4, 2, 1}) {
    if (static_cast<size_t>(cw) * HW * 4 <= 4 * 1024) { CW = cw; break; }
  }
  {
    const int cw = option(ISB_OPT_GATHER_CW, 0);
    if (cw >= 1 && cw <= 64) CW = cw;
  }
  const int row_align = (W % 4 == 0) ? 1 : ((W % 2 == 0) ? 2 : 4);
  const size_t slot = static_cast<size_t>(CW) * HW * 4;            // one warp's ring slot
  const int units = static_cast<int>((C + kGatherWarps * CW - 1) / (kGatherWarps * CW));   // unit rounds per image
  // units per warp: the ring needs a few to run ahead of; keep >= 8 waves of CTAs
  int GG = kGatherDefaultG;
  GG = option(ISB_OPT_GATHER_G, GG);
  if (GG < 1) GG = 1;
  while (GG > 1 && B * ((units + GG - 1) / GG) < 148 * 8) GG >>= 1;
  if (GG > units) GG = units;
  int NST = kGatherDefaultStages;
  NST = option(ISB_OPT_GATHER_STAGES, NST);
  if (NST < 1) NST = 1;
  if (NST > kGatherMaxStages) NST = kGatherMaxStages;
  if (NST > GG + 1) NST = GG + 1;                  // no point in more slots than units + 1
  while (NST > 1 && NST * kGatherWarps * slot > 200 * 1024) --NST;
  const size_t smem = NST * kGatherWarps * slot;
  ISB_CHECK_ARG(smem <= 200 * 1024, "isb_region_gather: feature map too large (H*W=%lld)", (long long)HW);
  const int CBg = CW;
  dim3 grid(static_cast<unsigned>((units + GG - 1) / GG), static_cast<unsigned>(B));
  if (fh == 7 && fw == 7) {
    if (smem > 48 * 1024)
      ISB_CUDA(cudaFuncSetAttribute(region_gather_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_gather_kernel<7><<<grid, kGatherThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        x, (int)C, (int)H, (int)W, fh, fw, k, k_sum, CBg, GG, NST, row_align, image_list, n_list, idx, nsel, win_norm, shift, U_hi, U_lo, ldu, win_mean);
  } else {
    if (smem > 48 * 1024)
      ISB_CUDA(cudaFuncSetAttribute(region_gather_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_gather_kernel<0><<<grid, kGatherThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        x, (int)C, (int)H, (int)W, fh, fw, k, k_sum, CBg, GG, NST, row_align, image_list, n_list, idx, nsel, win_norm, shift, U_hi, U_lo, ldu, win_mean);
  }
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_region_logits(const float* win_mean, const float* cls_w, const float* cls_b, int64_t B,
                                 int64_t C, int64_t ncls, int ke, int k, const int32_t* nsel_in,
                                 const float* approx_max, const float* runner_up, const int64_t* idx_in,
                                 const float* norm_in, int64_t* idx_out, float* norm_out, int32_t* nsel_out,
                                 float* cls_out, const float* approx_cls, float wabs_max,
                                 int32_t* changed_list, int32_t* n_changed, int32_t* n_uncertified,
                                 void* stream) {
  ISB_CHECK_ARG(win_mean && cls_w && cls_b && nsel_in && idx_in && norm_in && idx_out && norm_out && nsel_out,
                "isb_region_logits: null pointer");
  ISB_CHECK_ARG(cls_out != nullptr || (approx_cls != nullptr && wabs_max > 0.f),
                "isb_region_logits: without cls_out, approx_cls and max|cls_w| are needed to prune the classes");
  ISB_CHECK_ARG(B > 0 && C > 0 && ncls > 0 && k >= 1 && ke >= k && ke <= kSelMaxCand, "isb_region_logits: bad shape");
  ISB_CHECK_ARG((changed_list == nullptr) == (n_changed == nullptr), "isb_region_logits: changed_list and n_changed go together");
  ISB_CHECK_ARG(n_uncertified == nullptr || (approx_max != nullptr && runner_up != nullptr),
                "isb_region_logits: the certificate needs approx_max and runner_up");
  const size_t smem = (static_cast<size_t>(ke) * C + static_cast<size_t>(ke) * ncls) * 4;
  ISB_CHECK_ARG(smem <= 200 * 1024, "isb_region_logits: ke * (C + ncls) too large for shared memory");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_uncertified != nullptr) ISB_CUDA(cudaMemsetAsync(n_uncertified, 0, 4, st));
  if (n_changed != nullptr) ISB_CUDA(cudaMemsetAsync(n_changed, 0, 4, st));
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(region_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  region_logits_kernel<<<static_cast<unsigned>(B), kLogThreads, smem, st>>>(
      win_mean, cls_w, cls_b, (int)C, (int)ncls, ke, k, nsel_in, approx_max, runner_up, idx_in, norm_in, idx_out,
      norm_out, nsel_out, cls_out, approx_cls, wabs_max, changed_list, n_changed, n_uncertified);
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
