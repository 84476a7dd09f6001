// Cosine top-k search:  bf16 tcgen05 screen with fused streaming top-k  +
// exact (fp64-accumulated) re-rank of the surviving candidates, and the R-way
// merge used after the multi-GPU all-gather.
//
// replaces  sim = torch.mm(Q, DB.t())  (test/siamese_regions_test.py:76,
// utils/train_siamese.py:70) + the descending sort / max / kthvalue that consume
// it (utils/metrics.py:11,13,33) of the reference.
#include <cstdlib>
#include "isb_host.cuh"
#include "isb_topk.cuh"

namespace isb {

// ------------------------------------------------------------------ small kernels
__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
  // round-to-nearest-even; NaN stays NaN
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
__device__ __forceinline__ float bf16_to_f32(uint16_t h) {
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

// one thread = 8 output columns (one 16-byte store)
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                   int64_t ldx, uint16_t* __restrict__ y, int64_t ldy, int part) {
  const int64_t groups_per_row = ldy / 8;
  const int64_t total = rows * groups_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / groups_per_row;
    const int64_t c0 = (i - r * groups_per_row) * 8;
    const float* src = x + r * ldx + c0;
    float v[8];
    if (c0 + 8 <= cols && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? __ldg(src + j) : 0.f;
    }
    uint16_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f = v[j];
      uint16_t b = f32_to_bf16_rn(f);
      for (int t = 0; t < part; ++t) {  // peel `part` leading bf16 terms
        f -= bf16_to_f32(b);
        b = f32_to_bf16_rn(f);
      }
      h[j] = b;
    }
    uint4 o;
    o.x = h[0] | (static_cast<uint32_t>(h[1]) << 16);
    o.y = h[2] | (static_cast<uint32_t>(h[3]) << 16);
    o.z = h[4] | (static_cast<uint32_t>(h[5]) << 16);
    o.w = h[6] | (static_cast<uint32_t>(h[7]) << 16);
    *reinterpret_cast<uint4*>(y + r * ldy + c0) = o;
  }
}

// ------------------------------------------------------------------ exact re-rank
// One CTA per query row.
//   1. radix-select (4 x 8 bits) the kc best screen scores among the row's pool
//      entries (<= n_groups * 128);
//   2. exact score of each selected database row: fp32 products accumulated in
//      fp64 (one warp per candidate, 16-byte loads);
//   3. bitonic sort by (score desc, index asc); write the first k.
constexpr int kRerankThreads = 256;
constexpr float kCertZ = 8.f;   // certificate: the k-th exact score must clear t_min by kCertZ sigma

struct SortEntry {
  double score;
  int col;
};

__device__ __forceinline__ bool entry_before(double sa, int ia, double sb, int ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

// Shared state of the candidate selection (one CTA = one row).
struct SelectSmem {
  int hist[256];
  int total, sel, eq_taken, remaining;
  uint32_t prefix;
  uint32_t min_key;               // smallest screen key among the selected
  double sel_score[kMaxCand];
  int sel_col[kMaxCand];
  float sel_screen[kMaxCand];     // screen score of each selected candidate
  float sel_norm2[kMaxCand];      // squared L2 norm of each selected database row (exact_scores)
};

// Pick the kc best screen scores among the row's pool entries (4 x 8-bit radix
// select), leave their columns in sm.sel_col[0..n) and return n = min(total, kc).
// sm.total = number of pool entries; sm.min_key = key of the worst selected one.
__device__ __forceinline__ int select_pool_candidates(SelectSmem& sm, const uint2* __restrict__ rpool,
                                                      const int* __restrict__ rcnt, int n_groups, int kc) {
  const int tid = threadIdx.x;
  const int slots = n_groups * kMaxCand;
  if (tid == 0) {
    int t = 0;
    for (int g = 0; g < n_groups; ++g) t += rcnt[g];
    sm.total = t;
    sm.prefix = 0;
    sm.remaining = kc;
    sm.sel = 0;
    sm.eq_taken = 0;
    sm.min_key = 0xFFFFFFFFu;
  }
  __syncthreads();
  const int total = sm.total;
  const int want = min(total, kc);

  // threshold key T: the want-th largest key (only needed if total > kc)
  uint32_t T = 0;
  if (total > kc) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      sm.hist[tid] = 0;  // kRerankThreads == 256 bins
      __syncthreads();
      const uint32_t prefix = sm.prefix;
      const uint32_t pmask = (pass == 0) ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int e = tid; e < slots; e += kRerankThreads) {
        if ((e & (kMaxCand - 1)) < rcnt[e / kMaxCand]) {
          const uint32_t key = f2key(rpool[e].x);
          if ((key & pmask) == prefix) atomicAdd(&sm.hist[(key >> shift) & 255], 1);
        }
      }
      __syncthreads();
      if (tid == 0) {
        int rem = sm.remaining;  // how many still to take from keys matching the prefix
        int d = 255;
        for (; d > 0; --d) {
          if (sm.hist[d] >= rem) break;
          rem -= sm.hist[d];
        }
        sm.prefix = prefix | (static_cast<uint32_t>(d) << shift);
        sm.remaining = rem;
      }
      __syncthreads();
    }
    T = sm.prefix;
  }
  // after the 4 passes sm.remaining = how many entries with key == T to take
  const int quota_eq = (total > kc) ? sm.remaining : 0x7FFFFFFF;
  for (int e = tid; e < slots; e += kRerankThreads) {
    if ((e & (kMaxCand - 1)) < rcnt[e / kMaxCand]) {
      const uint2 ent = rpool[e];
      const uint32_t key = f2key(ent.x);
      bool take = key > T;
      if (!take && key == T) take = atomicAdd(&sm.eq_taken, 1) < quota_eq;
      if (take) {
        const int pos = atomicAdd(&sm.sel, 1);
        if (pos < kMaxCand) {
          sm.sel_col[pos] = static_cast<int>(ent.y);
          sm.sel_screen[pos] = __uint_as_float(ent.x);
        }
        atomicMin(&sm.min_key, key);
      }
    }
  }
  __syncthreads();
  return min(sm.sel, want);
}

// Exact scores of the n selected rows of `db` against the query in qs (smem):
// fp32 products accumulated in fp64, one warp per candidate, 16-byte loads.
__device__ __forceinline__ void exact_scores(SelectSmem& sm, const float* qs,
                                             const float* __restrict__ db, int D, int n_sel) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < kMaxCand; c += kRerankThreads / 32) {
    if (c < n_sel) {
      const float4* dr = reinterpret_cast<const float4*>(db + static_cast<size_t>(sm.sel_col[c]) * D);
      double acc = 0.0;
      float n2 = 0.f;
      for (int i = lane; i < D / 4; i += 32) {
        const float4 b = __ldg(dr + i);
        const float4 a = reinterpret_cast<const float4*>(qs)[i];
        acc = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc);
        acc = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc);
        acc = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc);
        acc = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc);
        n2 = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, n2))));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
      }
      if (lane == 0) { sm.sel_score[c] = acc; sm.sel_norm2[c] = n2; }
    } else if (lane == 0) {
      sm.sel_score[c] = -INFINITY;
      sm.sel_col[c] = 0x7FFFFFFF;
      sm.sel_norm2[c] = 0.f;
    }
  }
  __syncthreads();
}

// Completeness certificate of a screened row.  Every database row that is NOT
// among the selected candidates has a screen score <= t_min (the worst selected
// one): the streaming filter only ever drops a score that is <= the kc-th best
// seen so far.  Its exact score is therefore <= t_min + (screen error).  The
// screen error is measured, not assumed: sigma = rms(screen - exact) over the
// row's own candidates.  The row is certified when the exact k-th best score
// clears t_min by z sigma plus an fp32 accumulation floor; otherwise the caller
// re-screens it with fp32-grade operands (isb_topk_resolve) or exhaustively.
// The certificate is statistical (z sigma), not a proof.  A sigma measured on a handful of
// candidates can be far too small (one sample: ~0), so it is floored by the noise the screen's
// operand rounding is EXPECTED to have on dense rows, sigma_floor (screen_sigma_floor below; 0
// for the fp32-grade split-operand screen, whose measured sigma is at fp32-accumulation level).
__device__ __forceinline__ bool row_certified(const SelectSmem& sm, int kc, double sigma2_sum, int n_sel,
                                              double kth_exact, float cert_z, double sigma_floor) {
  if (sm.total < kc) return true;  // nothing was ever dropped for this row
  const double sigma = fmax(sqrt(sigma2_sum / static_cast<double>(n_sel > 0 ? n_sel : 1)), sigma_floor);
  const double t_min = static_cast<double>(__uint_as_float(key2f(sm.min_key)));
  const double floor_ = 4e-7 * fmax(fabs(kth_exact), fabs(t_min)) + 1e-30;
  return kth_exact - t_min > static_cast<double>(cert_z) * sigma + floor_;
}

// Expected rms error of a plain-bf16 screen score on dense rows: both operands are rounded to 8
// significant bits, so a product is off by ~2.3e-3 relative (rms) and the sum of D such terms by
// that times sqrt(sum (a_i b_i)^2) ~ |a||b|/sqrt(D): sigma = 2.34e-3 |a||b| / sqrt(D), calibrated
// on unit Gaussian rows at D = 128 and 2048 (5.2e-5 at D = 2048).  Half of it is used as the
// floor when the sample is large (>= 32 candidates measure sigma well), one and a half times it
// when it is small.
// qn2: squared norm of the query row, sm.sel_norm2: of the candidates (the largest is taken).
__device__ __forceinline__ double screen_sigma_floor(const SelectSmem& sm, int n_sel, double qn2, int D) {
  float bn2 = 0.f;
  for (int c = 0; c < n_sel; ++c) bn2 = fmaxf(bn2, sm.sel_norm2[c]);
  const double dense = 2.34e-3 * sqrt(qn2 * static_cast<double>(bn2) / static_cast<double>(D));
  return (n_sel >= 32 ? 0.5 : 1.5) * dense;
}

// squared norm of the query row staged in qs[0..D): every thread returns the same value
__device__ __forceinline__ double block_norm2(const float* qs, int D, double* red) {
  double a = 0.0;
  for (int i = threadIdx.x; i < D; i += kRerankThreads) a = fma(static_cast<double>(qs[i]), static_cast<double>(qs[i]), a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kRerankThreads / 32; ++w) t += red[w];
  __syncthreads();
  return t;
}

// Bitonic sort of sm.sel_score / sm.sel_col [kMaxCand], best first (score desc, column
// asc); called by every thread of the CTA after a __syncthreads().
__device__ __forceinline__ void sort_selected(SelectSmem& sm) {
  const int tid = threadIdx.x;
  double* sel_score = sm.sel_score;
  int* sel_col = sm.sel_col;
  for (int size = 2; size <= kMaxCand; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < kMaxCand) {
        const int partner = tid ^ stride;
        if (partner > tid) {
          const bool up = (tid & size) == 0;  // "up" blocks sort best-first
          const double sa = sel_score[tid], sb = sel_score[partner];
          const int ia = sel_col[tid], ib = sel_col[partner];
          const bool a_first = entry_before(sa, ia, sb, ib);
          if (a_first != up) {
            sel_score[tid] = sb; sel_score[partner] = sa;
            sel_col[tid] = ib; sel_col[partner] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per listed row.  row_map == nullptr: CTA p handles query row p and pool
// row p.  Otherwise (resolve pass) pool row p belongs to query row row_map[p].
__global__ void __launch_bounds__(kRerankThreads)
rerank_kernel(const float* __restrict__ q, const float* __restrict__ db, int D, int n_groups,
              const uint2* __restrict__ pool, const int* __restrict__ pool_cnt, int kc, int k,
              int64_t idx_offset, const int* __restrict__ row_map, float cert_z, int plain_screen,
              int* __restrict__ unc_rows, int* __restrict__ unc_count,
              float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  float* qs = reinterpret_cast<float*>(rr_smem);  // [D]
  __shared__ SelectSmem sm;
  __shared__ double s_sig2;
  __shared__ double s_red[kRerankThreads / 32];
  double* sel_score = sm.sel_score;
  int* sel_col = sm.sel_col;

  const int prow = blockIdx.x;
  const int row = (row_map != nullptr) ? row_map[prow] : prow;
  const int tid = threadIdx.x;
  if (tid == 0) s_sig2 = 0.0;
  for (int i = tid; i < D / 4; i += kRerankThreads)
    reinterpret_cast<float4*>(qs)[i] = __ldg(reinterpret_cast<const float4*>(q + static_cast<size_t>(row) * D) + i);
  const int n_sel = select_pool_candidates(sm, pool + static_cast<size_t>(prow) * n_groups * kMaxCand,
                                           pool_cnt + static_cast<size_t>(prow) * n_groups, n_groups, kc);
  exact_scores(sm, qs, db, D, n_sel);
  const double qn2 = (unc_count != nullptr && plain_screen) ? block_norm2(qs, D, s_red) : 0.0;
  double sigma_floor = 0.0;
  if (unc_count != nullptr && plain_screen && tid == 0) sigma_floor = screen_sigma_floor(sm, n_sel, qn2, D);

  // screen noise of this row: sum over its candidates of (screen - exact)^2
  if (unc_count != nullptr) {
    double d2 = 0.0;
    if (tid < n_sel) {
      const double d = static_cast<double>(sm.sel_screen[tid]) - sel_score[tid];
      d2 = d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((tid & 31) == 0 && tid < kMaxCand) atomicAdd(&s_sig2, d2);
  }
  __syncthreads();

  // ---- 3. bitonic sort of 128 entries (first 128 threads)
  sort_selected(sm);
  for (int j = tid; j < k; j += kRerankThreads) {
    const bool ok = j < n_sel;
    out_scores[static_cast<size_t>(row) * k + j] = ok ? static_cast<float>(sel_score[j]) : -INFINITY;
    out_idx[static_cast<size_t>(row) * k + j] = ok ? static_cast<int64_t>(sel_col[j]) + idx_offset : -1;
  }
  if (unc_count != nullptr && tid == 0) {
    const bool ok = (k <= n_sel) && row_certified(sm, kc, s_sig2, n_sel, sel_score[k - 1], cert_z, sigma_floor);
    if (!ok) unc_rows[atomicAdd(unc_count, 1)] = row;
  }
}

// ------------------------------------------------------------------ sharded search: candidate exchange
// Row-sharded database: re-ranking k + margin candidates on EVERY shard
// would cost R times the gathers of the single-GPU search although only k + margin of
// the R * (k + margin) candidates can matter.  Instead the shards exchange the screen
// scores of their candidates (4 bytes each), every rank derives the same global
// threshold -- the kc-th best screen score over all shards -- and re-ranks only its own
// candidates at or above it: kc gathers per query in total instead of R * kc.
//
// The selections below are warp-cooperative: one warp owns one query row, its entries sit
// in shared memory and the k-th largest key is found by a 32-step bitwise binary search
// (count of keys >= candidate per step) -- no block barriers, four rows per CTA.
constexpr int kSelWarps = 4;   // rows per CTA (launches may use fewer when a row needs much shared memory)

__host__ __device__ __forceinline__ size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }


// Threshold seeding.  A row that starts its stream from -inf passes ~kc ln(N / kc) values and
// the rows of a warp compact their buffers at the same tiles while they warm up -- a fixed
// ~1.7 M cycles per launch that the MMA pipe spends waiting (1 ms: a quarter of the screen on a
// 125k-row shard).  The kc-th best score of ANY subset of the columns is a lower bound of the
// kc-th best over all of them, so the scores of the first kSeedRows database rows (one plain
// GEMM, [Q, kSeedRows] fp32) give every row a safe starting threshold: its kc-th largest sample
// score minus a slack that covers any difference in accumulation order between the two kernels.
// One warp per row; keys in shared memory, 4 x 8-bit radix select.
constexpr int kSeedRows = 2048;
constexpr float kSeedSlack = 1e-5f;

__global__ void __launch_bounds__(32 * kSelWarps)
seed_threshold_kernel(const float* __restrict__ sample, int64_t Q, int S, int kth, uint32_t* __restrict__ gthr) {
  extern __shared__ __align__(16) uint8_t seed_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kSelWarps + warp;
  if (row >= Q) return;
  uint32_t* keys = reinterpret_cast<uint32_t*>(seed_smem) + static_cast<size_t>(warp) * (S + 256);
  int* hist = reinterpret_cast<int*>(keys + S);
  const float* src = sample + row * S;
  for (int j = lane; j < S; j += 32) keys[j] = f2key(__float_as_uint(__ldg(src + j)));
  __syncwarp();
  const uint32_t T = warp_kth_largest(keys, S, kth, hist);
  if (lane == 0) {
    const float v = __uint_as_float(key2f(T)) - kSeedSlack;
    gthr[row] = (v == v) ? f2key(__float_as_uint(v)) : kKeyNegInf;   // NaN: no seed
  }
}

// (1) the kc best screen entries of every row of the local pool, unsorted (entries above
// the kc-th key in slot order, then as many ties with it as fit: deterministic)
__global__ void __launch_bounds__(32 * kSelWarps)
pool_candidates_kernel(int64_t Q, int n_groups, const uint2* __restrict__ pool,
                       const int* __restrict__ pool_cnt, int kc, int kc_out,
                       float* __restrict__ cand_screen, int* __restrict__ cand_col) {
  extern __shared__ __align__(16) uint8_t pc_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= Q) return;
  const int slots = n_groups * kMaxCand;
  // per warp: keys [slots] | cols [slots] | hist [256] | pref [n_groups + 1]
  const size_t per_warp = static_cast<size_t>(2 * slots + 256 + n_groups + 1);
  uint32_t* keys = reinterpret_cast<uint32_t*>(pc_smem) + static_cast<size_t>(warp) * per_warp;
  uint32_t* cols = keys + slots;
  int* hist = reinterpret_cast<int*>(cols + slots);
  int* pref = hist + 256;
  const uint2* rpool = pool + static_cast<size_t>(row) * slots;
  const int* rcnt = pool_cnt + static_cast<size_t>(row) * n_groups;
  // group g holds rcnt[g] leading entries: exclusive prefix of the counts, then ONE flat pass
  // over all slots (independent loads) compacts the valid entries
  int total = 0;
  for (int g0 = 0; g0 < n_groups; g0 += 32) {
    const int g = g0 + lane;
    const int c = (g < n_groups) ? rcnt[g] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (g < n_groups) pref[g] = total + incl - c;
    total += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) pref[n_groups] = total;
  __syncwarp();
  constexpr int kBatch = 8;   // loads in flight per lane (the smem stores below would otherwise
                              // serialise against the next iteration's reads of pref)
  for (int base = 0; base < slots; base += 32 * kBatch) {
    uint2 ent[kBatch];
    int dst[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = base + u * 32 + lane;
      dst[u] = -1;
      if (e < slots) {
        const int g = e / kMaxCand, j = e - g * kMaxCand;
        const int p0 = pref[g];
        if (j < pref[g + 1] - p0) {
          dst[u] = p0 + j;
          ent[u] = __ldg(rpool + e);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (dst[u] >= 0) {
        keys[dst[u]] = f2key(ent[u].x);
        cols[dst[u]] = ent[u].y;
      }
    }
  }
  __syncwarp();
  const int want = min(total, kc);
  uint32_t T = 0u;
  int n_gt = total;
  if (total > kc) {
    T = warp_kth_largest(keys, total, kc, hist);
    int c = 0;
    for (int e = lane; e < total; e += 32) c += (keys[e] > T) ? 1 : 0;
    n_gt = __reduce_add_sync(0xffffffffu, c);
  }
  float* os = cand_screen + static_cast<size_t>(row) * kc_out;
  int* oc = cand_col + static_cast<size_t>(row) * kc_out;
  int pos_gt = 0, pos_eq = n_gt;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < total; base += 32) {
    const int e = base + lane;
    const uint32_t key = (e < total) ? keys[e] : 0u;
    const bool gt = (e < total) && (total <= kc || key > T);
    const bool eq = (e < total) && !gt && key == T;
    const uint32_t bgt = __ballot_sync(0xffffffffu, gt), beq = __ballot_sync(0xffffffffu, eq);
    if (gt) {
      const int p = pos_gt + __popc(bgt & lt_mask);
      os[p] = __uint_as_float(key2f(key));
      oc[p] = static_cast<int>(cols[e]);
    } else if (eq) {
      const int p = pos_eq + __popc(beq & lt_mask);
      if (p < want) {
        os[p] = __uint_as_float(key2f(key));
        oc[p] = static_cast<int>(cols[e]);
      }
    }
    pos_gt += __popc(bgt);
    pos_eq += __popc(beq);
  }
  for (int j = want + lane; j < kc_out; j += 32) {
    os[j] = -INFINITY;
    oc[j] = -1;
  }
}

// Same contract, one CTA per row (4 x 8-bit radix select): used when a row's pool does not
// fit one warp's share of shared memory (very few query rows => very many n-groups).
__global__ void __launch_bounds__(kRerankThreads)
pool_candidates_cta_kernel(int n_groups, const uint2* __restrict__ pool, const int* __restrict__ pool_cnt,
                           int kc, int kc_out, float* __restrict__ cand_screen, int* __restrict__ cand_col) {
  __shared__ SelectSmem sm;
  const int row = blockIdx.x, tid = threadIdx.x;
  const int n_sel = select_pool_candidates(sm, pool + static_cast<size_t>(row) * n_groups * kMaxCand,
                                           pool_cnt + static_cast<size_t>(row) * n_groups, n_groups, kc);
  for (int j = tid; j < kc_out; j += kRerankThreads) {
    const bool ok = j < n_sel;
    cand_screen[static_cast<size_t>(row) * kc_out + j] = ok ? sm.sel_screen[j] : -INFINITY;
    cand_col[static_cast<size_t>(row) * kc_out + j] = ok ? sm.sel_col[j] : -1;
  }
}

// (2) thr[row] = the kth-largest of the row's R * kc gathered screen scores
// (all_screen [R, Q, kc], -inf = no entry); -inf when fewer than kth entries exist, i.e.
// when no shard dropped anything that could matter.  kth = kc: every shard lists as many
// candidates as the result needs; kth > kc: reduced lists (see isb_topk_rerank_owned).
__global__ void __launch_bounds__(32 * kSelWarps)
global_threshold_kernel(const float* __restrict__ all_screen, int R, int64_t Q, int kc, int kth,
                        float* __restrict__ thr) {
  extern __shared__ __align__(16) uint8_t gt_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kSelWarps + warp;
  if (row >= Q) return;
  const int n = R * kc;
  uint32_t* keys = reinterpret_cast<uint32_t*>(gt_smem) + static_cast<size_t>(warp) * (n + 256);
  int* hist = reinterpret_cast<int*>(keys + n);
  int valid = 0;
  for (int e = lane; e < n; e += 32) {
    const int r = e / kc, j = e - r * kc;
    const uint32_t key = f2key(__float_as_uint(all_screen[(static_cast<size_t>(r) * Q + row) * kc + j]));
    keys[e] = key;
    valid += (key > kKeyNegInf) ? 1 : 0;
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  __syncwarp();
  float t = -INFINITY;
  if (valid >= kth) {
    t = __uint_as_float(key2f(warp_kth_largest(keys, n, kth, hist)));
  } else if (kth > kc) {
    // reduced lists: fewer than kth entries overall, yet a shard whose list is FULL (its last slot
    // is taken) may have dropped rows -- nothing can be concluded: +inf sends the row to the
    // second line (nothing is re-ranked, the merge cannot certify)
    bool full = false;
    for (int r = lane; r < R; r += 32) full |= keys[r * kc + kc - 1] > kKeyNegInf;
    if (__any_sync(0xffffffffu, full)) t = INFINITY;
  }
  if (lane == 0) thr[row] = t;
}

// (3) exact scores of the row's OWN candidates at or above the global threshold, sorted
// best first, plus the row's share of the screen-noise measurement, written as ONE packed
// row of 2k + 2 32-bit words (what the second all-gather exchanges).
__global__ void __launch_bounds__(kRerankThreads)
rerank_owned_kernel(const float* __restrict__ q, const float* __restrict__ db, int D, int kc, int k,
                    const float* __restrict__ cand_screen, const int* __restrict__ cand_col,
                    const float* __restrict__ thr, uint32_t* __restrict__ packed) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  float* qs = reinterpret_cast<float*>(rr_smem);  // [D]
  __shared__ SelectSmem sm;
  __shared__ double s_sig2;
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { s_sig2 = 0.0; sm.sel = 0; }
  for (int i = tid; i < D / 4; i += kRerankThreads)
    reinterpret_cast<float4*>(qs)[i] = __ldg(reinterpret_cast<const float4*>(q + static_cast<size_t>(row) * D) + i);
  __syncthreads();
  const float t = thr[row];
  int n_above = 0;   // listed entries strictly above the threshold
  for (int j = tid; j < kc; j += kRerankThreads) {
    const float s = cand_screen[static_cast<size_t>(row) * kc + j];
    const int c = cand_col[static_cast<size_t>(row) * kc + j];
    n_above += (c >= 0 && s > t) ? 1 : 0;
    if (c >= 0 && s >= t) {
      const int pos = atomicAdd(&sm.sel, 1);
      sm.sel_col[pos] = c;
      sm.sel_screen[pos] = s;
    }
  }
  n_above = __syncthreads_count(n_above);   // kc <= 128 < kRerankThreads: one entry per thread at most
  const int n_sel = sm.sel;
  // Reduced lists (kc shorter than the k + margin the threshold is the rank of): a FULL list whose
  // every entry lies strictly above the threshold may have been cut above it -- rows of this shard
  // that were not listed could still score above thr, and the global certificate would not cover
  // them.  The row is then handed to the merge as uncertifiable (infinite noise).  With full-length
  // lists the condition cannot occur: thr is the kc-th best overall, so no shard holds kc entries
  // strictly above it.
  const bool cut_above_thr = n_above == kc;
  exact_scores(sm, qs, db, D, n_sel);
  double d2 = 0.0;
  if (tid < n_sel) {
    const double d = static_cast<double>(sm.sel_screen[tid]) - sm.sel_score[tid];
    d2 = d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  if ((tid & 31) == 0 && tid < kMaxCand) atomicAdd(&s_sig2, d2);
  __syncthreads();
  sort_selected(sm);
  // packed row: k scores | k local rows (-1 = none) | sum (screen - exact)^2 | candidates scored
  uint32_t* out = packed + static_cast<size_t>(row) * (2 * k + 2);
  for (int j = tid; j < k; j += kRerankThreads) {
    const bool ok = j < n_sel;
    out[j] = __float_as_uint(ok ? static_cast<float>(sm.sel_score[j]) : -INFINITY);
    out[k + j] = static_cast<uint32_t>(ok ? sm.sel_col[j] : -1);
  }
  if (tid == 0) {
    out[2 * k] = __float_as_uint(cut_above_thr ? INFINITY : static_cast<float>(s_sig2));
    out[2 * k + 1] = __float_as_uint(static_cast<float>(n_sel));
  }
}

// ------------------------------------------------------------------ resolve pass operands
// a_hi[p, :], a_lo[p, :] = the two leading bf16 terms of q[rows[p], :]
__global__ void gather_q_terms_kernel(const float* __restrict__ q, const int* __restrict__ rows, int D,
                                      int64_t ldq, uint16_t* __restrict__ a_hi,
                                      uint16_t* __restrict__ a_lo) {
  const int p = blockIdx.x;
  const float* src = q + static_cast<size_t>(rows[p]) * D;
  for (int j = threadIdx.x; j < ldq; j += blockDim.x) {
    const float f = (j < D) ? __ldg(src + j) : 0.f;
    const uint16_t h = f32_to_bf16_rn(f);
    a_hi[p * ldq + j] = h;
    a_lo[p * ldq + j] = f32_to_bf16_rn(f - bf16_to_f32(h));
  }
}

// ------------------------------------------------------------------ exhaustive exact search
// For rows no screen can certify (dozens of database rows within fp32 noise of
// the k-th score, e.g. duplicated entries).  grid (chunks, rows): every CTA
// scores its chunk of the database exactly (fp64 accumulation) and keeps the
// chunk's k best in shared memory; exhaustive_merge_kernel then sorts the
// chunks * k survivors of a row.  Ties -> lower index.
struct ExhEntry {
  double score;
  int col;
  int pad;
};

__global__ void __launch_bounds__(kRerankThreads)
exhaustive_chunk_kernel(const float* __restrict__ q, const float* __restrict__ db, int N, int D, int k,
                        const int* __restrict__ rows, ExhEntry* __restrict__ part) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  float* qs = reinterpret_cast<float*>(rr_smem);  // [D]
  __shared__ double best_s[kMaxCand];
  __shared__ int best_c[kMaxCand];
  __shared__ int worst_pos;      // position of the current worst entry
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = gridDim.x, chunk = blockIdx.x, prow = blockIdx.y;
  const int row = rows[prow];
  for (int i = tid; i < D / 4; i += kRerankThreads)
    reinterpret_cast<float4*>(qs)[i] = __ldg(reinterpret_cast<const float4*>(q + static_cast<size_t>(row) * D) + i);
  for (int i = tid; i < kMaxCand; i += kRerankThreads) { best_s[i] = -INFINITY; best_c[i] = 0x7FFFFFFF; }
  if (tid == 0) worst_pos = 0;
  __syncthreads();
  const int lo = static_cast<int>(static_cast<long long>(chunk) * N / chunks);
  const int hi = static_cast<int>(static_cast<long long>(chunk + 1) * N / chunks);
  constexpr int kWarps = kRerankThreads / 32;
  for (int j0 = lo; j0 < hi; j0 += kWarps) {
    const int j = j0 + warp;
    double acc = -INFINITY;
    if (j < hi) {
      const float4* dr = reinterpret_cast<const float4*>(db + static_cast<size_t>(j) * D);
      acc = 0.0;
      for (int i = lane; i < D / 4; i += 32) {
        const float4 b = __ldg(dr + i);
        const float4 a = reinterpret_cast<const float4*>(qs)[i];
        acc = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc);
        acc = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc);
        acc = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc);
        acc = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    // insert in index order (warp w handles row j0 + w): deterministic
    for (int w = 0; w < kWarps; ++w) {
      __syncthreads();
      if (warp == w && j < hi) {
        const int wp = worst_pos;
        if (entry_before(acc, j, best_s[wp], best_c[wp])) {
          if (lane == 0) { best_s[wp] = acc; best_c[wp] = j; }
          __syncwarp();
          // new worst among the first k slots
          double ws = INFINITY; int wc = -1, wpos = 0;
          for (int i = lane; i < k; i += 32) {
            const double s_ = best_s[i]; const int c_ = best_c[i];
            if (wc == -1 || entry_before(ws, wc, s_, c_)) { ws = s_; wc = c_; wpos = i; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, ws, o);
            const int oc = __shfl_xor_sync(0xffffffffu, wc, o);
            const int op = __shfl_xor_sync(0xffffffffu, wpos, o);
            if (oc != -1 && (wc == -1 || entry_before(ws, wc, os, oc))) { ws = os; wc = oc; wpos = op; }
          }
          if (lane == 0) worst_pos = wpos;
        }
      }
    }
  }
  __syncthreads();
  ExhEntry* dst = part + (static_cast<size_t>(prow) * chunks + chunk) * k;
  for (int i = tid; i < k; i += kRerankThreads) {
    ExhEntry e; e.score = best_s[i]; e.col = best_c[i]; e.pad = 0;
    dst[i] = e;
  }
}

__global__ void __launch_bounds__(256)
exhaustive_merge_kernel(const ExhEntry* __restrict__ part, int chunks, int k, int n_pow2,
                        const int* __restrict__ rows, int64_t idx_offset,
                        float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) uint8_t mg_smem[];
  double* ss = reinterpret_cast<double*>(mg_smem);        // [n_pow2]
  int* sc = reinterpret_cast<int*>(ss + n_pow2);          // [n_pow2]
  const int prow = blockIdx.x, row = rows[prow];
  const int n = chunks * k;
  for (int e = threadIdx.x; e < n_pow2; e += blockDim.x) {
    if (e < n) {
      const ExhEntry x = part[static_cast<size_t>(prow) * n + e];
      ss[e] = x.score; sc[e] = x.col;
    } else {
      ss[e] = -INFINITY; sc[e] = 0x7FFFFFFF;
    }
  }
  __syncthreads();
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
        const int partner = t ^ stride;
        if (partner > t) {
          const bool up = (t & size) == 0;
          const double sa = ss[t], sb = ss[partner];
          const int ia = sc[t], ib = sc[partner];
          if (entry_before(sa, ia, sb, ib) != up) {
            ss[t] = sb; ss[partner] = sa; sc[t] = ib; sc[partner] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool ok = sc[j] != 0x7FFFFFFF;
    out_scores[static_cast<size_t>(row) * k + j] = ok ? static_cast<float>(ss[j]) : -INFINITY;
    out_idx[static_cast<size_t>(row) * k + j] = ok ? static_cast<int64_t>(sc[j]) + idx_offset : -1;
  }
}

// ------------------------------------------------------------------ negative mining
// Per-couple (semi-)hard negative of train/siamese_regions.py:106-129 without the
// N x N matrix: rows of the screen GEMM are the couples' anchors, columns all
// items; the streaming top-k epilogue masks same-label columns and (semi-hard)
// scores >= sim_pos; the survivors are re-checked exactly.

// dst[p, :] = src[rows[p], :]  (bf16 rows, 16-byte copies); row_label[p] = label[rows[p]]
__global__ void gather_rows_kernel(const uint16_t* __restrict__ src, int64_t ld,
                                   const int64_t* __restrict__ rows, uint16_t* __restrict__ dst,
                                   const int* __restrict__ label, int* __restrict__ row_label) {
  const int64_t p = blockIdx.x;
  const int64_t r = rows[p];
  const uint4* s4 = reinterpret_cast<const uint4*>(src + r * ld);
  uint4* d4 = reinterpret_cast<uint4*>(dst + p * ld);
  for (int64_t i = threadIdx.x; i < ld / 8; i += blockDim.x) d4[i] = __ldg(s4 + i);
  if (threadIdx.x == 0) row_label[p] = label[r];
}

// exact dot of every (anchor, positive) couple: one warp per couple
__global__ void pair_dot_kernel(const float* __restrict__ emb, int D, const int64_t* __restrict__ a,
                                const int64_t* __restrict__ b, int64_t P, double* __restrict__ out64,
                                float* __restrict__ out32) {
  const int64_t p = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= P) return;
  const float4* x = reinterpret_cast<const float4*>(emb + a[p] * D);
  const float4* y = reinterpret_cast<const float4*>(emb + b[p] * D);
  double acc = 0.0;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 u = __ldg(x + i), v = __ldg(y + i);
    acc = fma(static_cast<double>(u.x), static_cast<double>(v.x), acc);
    acc = fma(static_cast<double>(u.y), static_cast<double>(v.y), acc);
    acc = fma(static_cast<double>(u.z), static_cast<double>(v.z), acc);
    acc = fma(static_cast<double>(u.w), static_cast<double>(v.w), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) { out64[p] = acc; out32[p] = static_cast<float>(acc); }
}

// block-wide "best valid": larger score first, ties -> lower column
__device__ __forceinline__ void block_best(double& s, int& c, double* red_s, int* red_c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, s, o);
    const int oc = __shfl_xor_sync(0xffffffffu, c, o);
    if (os > s || (os == s && oc < c)) { s = os; c = oc; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red_s[warp] = s; red_c[warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kRerankThreads / 32; ++w)
      if (red_s[w] > red_s[0] || (red_s[w] == red_s[0] && red_c[w] < red_c[0])) {
        red_s[0] = red_s[w]; red_c[0] = red_c[w];
      }
  }
  __syncthreads();
  s = red_s[0];
  c = red_c[0];
}

// Error bound of a couple's screen scores from the candidates scored so far.  With >= 8 samples the
// measured rms (floored by the expected noise) is a usable sigma; a sigma from one or two samples is
// not (a single 4-sigma deviation, ~1 couple in 10^4, would read as a 4x noisier screen and reject
// the couple), so small samples take 2.5 x the floor (measured noise scatters up to 2.2 x around the
// dense-row expectation) and are only widened by an OBSERVED deviation: eps >= 2 max|screen - exact|.
constexpr int kMiningSigmaSamples = 8;
__device__ __forceinline__ double mining_eps(double d2_sum, int cnt, double max_abs_d, float screen_eps,
                                             float sigma_floor) {
  const double sigma = (cnt >= kMiningSigmaSamples)
                           ? fmax(sqrt(d2_sum / static_cast<double>(cnt)), static_cast<double>(sigma_floor))
                           : 2.5 * static_cast<double>(sigma_floor);
  double eps = fmax(static_cast<double>(screen_eps), static_cast<double>(kCertZ) * sigma);
  if (cnt < kMiningSigmaSamples) eps = fmax(eps, 2.0 * max_abs_d);
  return eps;
}

// Exact score (fp32 products, fp64 accumulation) of candidate c of the row, by one warp.
__device__ __forceinline__ double warp_exact_dot(const float* qs, const float* __restrict__ dbrow, int D) {
  const int lane = threadIdx.x & 31;
  const float4* dr = reinterpret_cast<const float4*>(dbrow);
  double acc = 0.0;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 b = __ldg(dr + i);
    const float4 a = reinterpret_cast<const float4*>(qs)[i];
    acc = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc);
    acc = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc);
    acc = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc);
    acc = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// One CTA per couple.  The kc best screen candidates of the couple are re-checked exactly
// PROGRESSIVELY: only ONE negative is wanted, and a candidate whose screen score lies more than
// eps below the best exact valid score found so far cannot win -- the statement the certificate
// makes about the columns outside the candidate list (t_min + eps < winner), applied inside the
// list.  Round 0 scores the candidates within 2 eps of the best screen score; every further
// round scores the ones the current winner does not yet exclude (eps grows with the noise measured
// on the candidates scored so far), until none is left.  On ordinary data one or two rows of D
// floats are gathered per couple instead of kc (16 x 8 KB x 16 384 couples = 2.1 GB at the
// BASELINE configs[2] size: 535 us of the 1.72 ms step).
__global__ void __launch_bounds__(kRerankThreads)
mining_rerank_kernel(const float* __restrict__ emb, int D, const int64_t* __restrict__ anchors,
                     int n_groups, const uint2* __restrict__ pool, const int* __restrict__ pool_cnt,
                     int kc, const double* __restrict__ pos64, int semi_hard, float screen_eps,
                     float sigma_floor, float ub_slack, int progressive, int64_t* __restrict__ neg_idx,
                     float* __restrict__ neg_sim, int* __restrict__ flag, int* __restrict__ unc_rows,
                     int* __restrict__ unc_count, unsigned long long* __restrict__ n_scored) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  float* qs = reinterpret_cast<float*>(rr_smem);
  __shared__ SelectSmem sm;
  __shared__ double red_s[kRerankThreads / 32];
  __shared__ int red_c[kRerankThreads / 32];
  __shared__ double s_sig2;
  __shared__ int s_cnt, s_more;
  __shared__ uint32_t s_topkey, s_maxd;
  __shared__ float s_thr;
  __shared__ unsigned char s_done[kMaxCand];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const float* qrow = emb + anchors[row] * D;
  for (int i = tid; i < D / 4; i += kRerankThreads)
    reinterpret_cast<float4*>(qs)[i] = __ldg(reinterpret_cast<const float4*>(qrow) + i);
  const int n_sel = select_pool_candidates(sm, pool + static_cast<size_t>(row) * n_groups * kMaxCand,
                                           pool_cnt + static_cast<size_t>(row) * n_groups, n_groups, kc);
  const double eps0 = mining_eps(0.0, progressive ? 0 : kMiningSigmaSamples, 0.0, screen_eps, sigma_floor);
  if (tid < kMaxCand) {
    s_done[tid] = 0;
    sm.sel_score[tid] = -INFINITY;
    if (tid >= n_sel) sm.sel_col[tid] = 0x7FFFFFFF;
  }
  if (tid == 0) { s_topkey = 0u; s_thr = -INFINITY; }
  __syncthreads();
  if (progressive) {
    if (tid < n_sel) atomicMax(&s_topkey, f2key(__float_as_uint(sm.sel_screen[tid])));
    __syncthreads();
    if (tid == 0 && n_sel > 0)
      s_thr = static_cast<float>(static_cast<double>(__uint_as_float(key2f(s_topkey))) - 2.0 * eps0);
    __syncthreads();
  }
  double s, eps = eps0;
  int c, scored_total = 0;
  for (;;) {
    const float thr = s_thr;
    for (int i = warp; i < n_sel; i += kRerankThreads / 32) {   // warp-uniform
      if (s_done[i] || !(sm.sel_screen[i] >= thr)) continue;
      const double v = warp_exact_dot(qs, emb + static_cast<size_t>(sm.sel_col[i]) * D, D);
      if ((tid & 31) == 0) { sm.sel_score[i] = v; s_done[i] = 1; }
    }
    if (tid == 0) { s_sig2 = 0.0; s_cnt = 0; s_more = 0; s_maxd = 0u; }
    __syncthreads();
    // screen noise of this couple: rms(screen - exact) over the candidates scored so far
    {
      double d2 = 0.0;
      int n1 = 0;
      if (tid < n_sel && s_done[tid]) {
        const double d = static_cast<double>(sm.sel_screen[tid]) - sm.sel_score[tid];
        d2 = d * d;
        n1 = 1;
        atomicMax(&s_maxd, __float_as_uint(fabsf(static_cast<float>(d)) * 1.000001f));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
      }
      if ((tid & 31) == 0 && n1) { atomicAdd(&s_sig2, d2); atomicAdd(&s_cnt, n1); }
    }
    // reference: excluded if S[i1, j] >= S[i1, i2]  (train/siamese_regions.py:111)
    s = -INFINITY;
    c = 0x7FFFFFFF;
    if (tid < n_sel && s_done[tid]) {
      const double v = sm.sel_score[tid];
      if (!semi_hard || v < pos64[row]) { s = v; c = sm.sel_col[tid]; }
    }
    block_best(s, c, red_s, red_c);   // (its barriers also order the atomics above)
    scored_total = s_cnt;
    eps = mining_eps(s_sig2, scored_total, static_cast<double>(__uint_as_float(s_maxd)), screen_eps, sigma_floor);
    // which of the candidates not yet scored can the current winner not exclude?
    double bound = (c != 0x7FFFFFFF) ? s - eps : -INFINITY;
    uint32_t my_key = 0u;
    if (tid < n_sel && !s_done[tid]) {
      if (static_cast<double>(sm.sel_screen[tid]) >= bound) atomicOr(&s_more, 1);
      my_key = f2key(__float_as_uint(sm.sel_screen[tid]));
    }
    if (tid == 0) s_topkey = 0u;
    __syncthreads();
    if (!s_more) break;
    if (c == 0x7FFFFFFF) {
      // no valid winner yet (semi-hard: everything scored so far is >= sim_pos): go down the list
      if (my_key) atomicMax(&s_topkey, my_key);
      __syncthreads();
      bound = static_cast<double>(__uint_as_float(key2f(s_topkey))) - 2.0 * eps;
    }
    if (tid == 0) s_thr = static_cast<float>(bound) - 1e-7f * fabsf(static_cast<float>(bound));  // never above the fp64 bound
    __syncthreads();
  }
  if (tid == 0) {
    const bool found = c != 0x7FFFFFFF;
    // Certificate: every column NOT selected has a screen score <= the worst selected one,
    // hence an exact score <= that + eps, with eps = max(the caller's absolute bound,
    // kCertZ x the screen noise: mining_eps over this couple's scored candidates).
    bool certified = true;
    if (sm.total > kc) {
      const float t_min = __uint_as_float(key2f(sm.min_key));
      certified = found && (static_cast<double>(t_min) + eps < s);
    }
    // semi-hard: the epilogue dropped the columns with screen >= sim_pos + ub_slack as "surely
    // excluded"; that holds only while the screen noise stays within the slack
    if (semi_hard && eps > static_cast<double>(ub_slack)) certified = false;
    neg_idx[row] = found ? static_cast<int64_t>(c) : -1;
    neg_sim[row] = found ? static_cast<float>(s) : -2.f;  // the reference's fill value (:124)
    flag[row] = certified ? 0 : 1;
    if (!certified && unc_rows != nullptr) unc_rows[atomicAdd(unc_count, 1)] = row;
    if (n_scored != nullptr) atomicAdd(n_scored, static_cast<unsigned long long>(scored_total));
  }
}

// The same re-check with ONE WARP per couple, for kc <= 32 (lane = candidate): the candidates come
// from pool_candidates_kernel (warp-level radix select), so no CTA-wide barrier and no serial
// histogram scan sits on the couple's critical path -- the one-CTA-per-couple kernel above spent
// ~35 us of latency per couple on its selection (16 384 couples = 14 waves = 0.5 ms), far more than
// the gathers.  cand_screen / cand_col [P, kc]: unsorted, (-inf, -1) beyond the couple's entries.
constexpr int kMineWarps = 8;

__global__ void __launch_bounds__(32 * kMineWarps)
mining_rerank_warp_kernel(const float* __restrict__ emb, int D, const int64_t* __restrict__ anchors, int64_t P,
                          const float* __restrict__ cand_screen, const int* __restrict__ cand_col, int kc,
                          const double* __restrict__ pos64, int semi_hard, float screen_eps, float sigma_floor,
                          float ub_slack, int progressive, int64_t* __restrict__ neg_idx,
                          float* __restrict__ neg_sim, int* __restrict__ flag, int* __restrict__ unc_rows,
                          int* __restrict__ unc_count) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kMineWarps + (threadIdx.x >> 5);
  if (row >= P) return;   // warp-uniform
  const float my_screen = (lane < kc) ? cand_screen[row * kc + lane] : -INFINITY;
  const int my_col = (lane < kc) ? cand_col[row * kc + lane] : -1;
  const bool valid = my_col >= 0;
  const int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
  const float* arow = emb + anchors[row] * D;
  const double ub = semi_hard ? pos64[row] : 0.0;
  // screen range of the candidates
  float t_min = valid ? my_screen : INFINITY, top = valid ? my_screen : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t_min = fminf(t_min, __shfl_xor_sync(0xffffffffu, t_min, o));
    top = fmaxf(top, __shfl_xor_sync(0xffffffffu, top, o));
  }
  const double eps0 = mining_eps(0.0, progressive ? 0 : kMiningSigmaSamples, 0.0, screen_eps, sigma_floor);
  float thr = progressive ? static_cast<float>(static_cast<double>(top) - 2.0 * eps0) : -INFINITY;
  double my_exact = -INFINITY;
  bool done = false;
  double s, eps;
  int c;
  for (;;) {
    uint32_t todo = __ballot_sync(0xffffffffu, valid && !done && my_screen >= thr);
    while (todo) {
      const int i = __ffs(todo) - 1;
      todo &= todo - 1;
      const int col = __shfl_sync(0xffffffffu, my_col, i);
      const double v = warp_exact_dot(arow, emb + static_cast<size_t>(col) * D, D);
      if (lane == i) { my_exact = v; done = true; }
    }
    // screen noise of this couple: rms(screen - exact) over the candidates scored so far
    double d2 = 0.0, dmax = 0.0;
    if (done) {
      const double d = static_cast<double>(my_screen) - my_exact;
      d2 = d * d;
      dmax = fabs(d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    const int cnt = __popc(__ballot_sync(0xffffffffu, done));
    eps = mining_eps(d2, cnt, dmax, screen_eps, sigma_floor);
    // reference: excluded if S[i1, j] >= S[i1, i2]  (train/siamese_regions.py:111); larger first, ties -> lower column
    s = -INFINITY;
    c = 0x7FFFFFFF;
    if (done && (!semi_hard || my_exact < ub)) { s = my_exact; c = my_col; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, s, o);
      const int oc = __shfl_xor_sync(0xffffffffu, c, o);
      if (os > s || (os == s && oc < c)) { s = os; c = oc; }
    }
    const bool found = c != 0x7FFFFFFF;
    double bound = found ? s - eps : -INFINITY;
    const bool open = valid && !done;
    if (!__ballot_sync(0xffffffffu, open && static_cast<double>(my_screen) >= bound)) break;
    if (!found) {
      // no valid winner yet (semi-hard: everything scored so far is >= sim_pos): go down the list
      float top2 = open ? my_screen : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) top2 = fmaxf(top2, __shfl_xor_sync(0xffffffffu, top2, o));
      bound = static_cast<double>(top2) - 2.0 * eps;
    }
    thr = static_cast<float>(bound);
    thr -= 1e-7f * fabsf(thr);   // never above the fp64 bound
  }
  if (lane == 0) {
    const bool found = c != 0x7FFFFFFF;
    // Certificate: a column outside the list has a screen score <= the worst listed one (only a full
    // list can have dropped anything), hence an exact score <= t_min + eps
    bool certified = true;
    if (n_valid >= kc) certified = found && (static_cast<double>(t_min) + eps < s);
    // semi-hard: the epilogue dropped the columns with screen >= sim_pos + ub_slack as "surely
    // excluded"; that holds only while the screen noise stays within the slack
    if (semi_hard && eps > static_cast<double>(ub_slack)) certified = false;
    neg_idx[row] = found ? static_cast<int64_t>(c) : -1;
    neg_sim[row] = found ? static_cast<float>(s) : -2.f;  // the reference's fill value (:124)
    flag[row] = certified ? 0 : 1;
    if (!certified && unc_rows != nullptr) unc_rows[atomicAdd(unc_count, 1)] = static_cast<int>(row);
  }
}

// Exhaustive exact pass for the (rare) rows the certificate rejected.
__global__ void __launch_bounds__(kRerankThreads)
mining_bruteforce_kernel(const float* __restrict__ emb, int N, int D, const int* __restrict__ label,
                         const int64_t* __restrict__ anchors, const double* __restrict__ pos64,
                         int semi_hard, const int* __restrict__ flag, int64_t* __restrict__ neg_idx,
                         float* __restrict__ neg_sim, int* __restrict__ n_brute) {
  const int row = blockIdx.x;
  if (flag[row] == 0) return;
  extern __shared__ __align__(16) uint8_t rr_smem[];
  float* qs = reinterpret_cast<float*>(rr_smem);
  __shared__ double red_s[kRerankThreads / 32];
  __shared__ int red_c[kRerankThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t a = anchors[row];
  for (int i = tid; i < D / 4; i += kRerankThreads)
    reinterpret_cast<float4*>(qs)[i] = __ldg(reinterpret_cast<const float4*>(emb + a * D) + i);
  __syncthreads();
  const int lab = label[a];
  const double ub = pos64[row];
  double best = -INFINITY;
  int best_c = 0x7FFFFFFF;
  for (int j = warp; j < N; j += kRerankThreads / 32) {
    if (label[j] == lab) continue;  // warp-uniform
    const float4* dr = reinterpret_cast<const float4*>(emb + static_cast<size_t>(j) * D);
    double acc = 0.0;
    for (int i = lane; i < D / 4; i += 32) {
      const float4 b = __ldg(dr + i);
      const float4 q4 = reinterpret_cast<const float4*>(qs)[i];
      acc = fma(static_cast<double>(q4.x), static_cast<double>(b.x), acc);
      acc = fma(static_cast<double>(q4.y), static_cast<double>(b.y), acc);
      acc = fma(static_cast<double>(q4.z), static_cast<double>(b.z), acc);
      acc = fma(static_cast<double>(q4.w), static_cast<double>(b.w), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((!semi_hard || acc < ub) && (acc > best || (acc == best && j < best_c))) { best = acc; best_c = j; }
  }
  block_best(best, best_c, red_s, red_c);
  if (tid == 0) {
    const bool found = best_c != 0x7FFFFFFF;
    neg_idx[row] = found ? static_cast<int64_t>(best_c) : -1;
    neg_sim[row] = found ? static_cast<float>(best) : -2.f;
    if (n_brute != nullptr) atomicAdd(n_brute, 1);
  }
}

// ------------------------------------------------------------------ R-way merge
// cand_scores / cand_idx: [R, Q, k].  One CTA per query; bitonic sort of the
// R*k (padded to a power of two <= kMergeMax) entries in shared memory.
constexpr int kMergeMax = 4096;

__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ cs, const int64_t* __restrict__ ci, int R, int64_t Q,
                  int k, int n_pow2, float* __restrict__ out_scores, int64_t* __restrict__ out_idx,
                  const float2* __restrict__ stat, const float* __restrict__ thr, float cert_z,
                  int* __restrict__ unc_rows, int* __restrict__ unc_count) {
  extern __shared__ __align__(16) uint8_t mg_smem[];
  int64_t* sidx = reinterpret_cast<int64_t*>(mg_smem);           // [n_pow2]
  float* sscore = reinterpret_cast<float*>(sidx + n_pow2);        // [n_pow2]
  const int64_t row = blockIdx.x;
  const int n = R * k;
  for (int e = threadIdx.x; e < n_pow2; e += blockDim.x) {
    if (e < n) {
      const int r = e / k, j = e - r * k;
      const size_t off = (static_cast<size_t>(r) * Q + row) * k + j;
      sscore[e] = cs[off];
      sidx[e] = ci[off];
    } else {
      sscore[e] = -INFINITY;
      sidx[e] = INT64_MAX;
    }
  }
  __syncthreads();
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
        const int partner = t ^ stride;
        if (partner > t) {
          const bool up = (t & size) == 0;
          const float sa = sscore[t], sb = sscore[partner];
          const int64_t ia = sidx[t], ib = sidx[partner];
          // invalid (index < 0) entries sort last
          const bool va = ia >= 0, vb = ib >= 0;
          const bool a_first = (va != vb) ? va : ((sa > sb) || (sa == sb && ia < ib));
          if (a_first != up) {
            sscore[t] = sb; sscore[partner] = sa;
            sidx[t] = ib; sidx[partner] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool ok = j < n && sidx[j] != INT64_MAX;
    out_scores[row * k + j] = ok ? sscore[j] : -INFINITY;
    out_idx[row * k + j] = ok ? sidx[j] : -1;
  }
  // Completeness certificate of the sharded search (see row_certified): every database row
  // that was not re-ranked on its shard has a screen score <= thr[row]; the screen noise is
  // the rms over ALL shards' re-ranked candidates of the row (stat [R, Q]).
  if (unc_count != nullptr && threadIdx.x == 0) {
    const float t_min = thr[row];
    bool ok = true;
    if (t_min > -INFINITY) {
      double s2 = 0.0, cnt = 0.0;
      for (int r = 0; r < R; ++r) {
        const float2 st = stat[static_cast<size_t>(r) * Q + row];
        s2 += static_cast<double>(st.x);
        cnt += static_cast<double>(st.y);
      }
      const bool have_k = (k - 1 < n) && sidx[k - 1] != INT64_MAX && sidx[k - 1] >= 0;
      const double kth = static_cast<double>(sscore[k - 1]);
      const double sigma = sqrt(s2 / (cnt > 0.0 ? cnt : 1.0));
      const double floor_ = 4e-7 * fmax(fabs(kth), fabs(static_cast<double>(t_min))) + 1e-30;
      ok = have_k && (kth - static_cast<double>(t_min) > static_cast<double>(cert_z) * sigma + floor_);
    }
    if (!ok) unc_rows[atomicAdd(unc_count, 1)] = static_cast<int>(row);
  }
}

// Fast path for k <= kMaxCand: one warp per query.  The k-th best (score, index) entry is
// found by the bitwise search on the score key and, only when more entries tie with it
// than there is room for, a second search on the index; the k survivors are sorted by a
// warp-synchronous bitonic network (4 entries per lane).
__device__ __forceinline__ bool merge_before(float sa, int64_t ia, float sb, int64_t ib) {
  // invalid (index < 0 or the INT64_MAX padding) entries sort last
  const bool va = ia >= 0 && ia != INT64_MAX, vb = ib >= 0 && ib != INT64_MAX;
  if (va != vb) return va;
  return (sa > sb) || (sa == sb && ia < ib);
}

// PACKED: the shards' lists arrive as ONE all-gathered buffer, packed [R, Q, 2k + 2] 32-bit
// words per (shard, query): k scores (fp32) | k local rows (int32, -1 = none) | the two
// stat words; global index = local row + row_offset[shard].  cs / ci / stat are unused.
template <bool PACKED>
__global__ void __launch_bounds__(32 * kSelWarps)
topk_merge_warp_kernel(const float* __restrict__ cs, const int64_t* __restrict__ ci,
                       const uint32_t* __restrict__ packed, const int64_t* __restrict__ row_offset, int R,
                       int64_t Q, int k, float* __restrict__ out_scores, int64_t* __restrict__ out_idx,
                       const float2* __restrict__ stat, const float* __restrict__ thr, float cert_z,
                       int* __restrict__ unc_rows, int* __restrict__ unc_count) {
  extern __shared__ __align__(16) uint8_t mw_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kSelWarps + warp;
  if (row >= Q) return;
  const int n = R * k;
  // per warp: keys [n] u32 | sel_idx [kMaxCand] i64 | sel_score [kMaxCand] f32 | hist [256]
  const size_t per_warp = align_up_dev(static_cast<size_t>(n) * 4, 8) + kMaxCand * 12 + 1024;
  uint8_t* base = mw_smem + warp * per_warp;
  uint32_t* keys = reinterpret_cast<uint32_t*>(base);
  int64_t* sidx = reinterpret_cast<int64_t*>(base + align_up_dev(static_cast<size_t>(n) * 4, 8));
  float* sscore = reinterpret_cast<float*>(sidx + kMaxCand);
  int* hist = reinterpret_cast<int*>(sscore + kMaxCand);
  const int pw = 2 * k + 2;   // packed words per (shard, query)

  auto entry_score = [&](int e) -> float {
    const int r = e / k, j = e - r * k;
    if (PACKED) return __uint_as_float(packed[(static_cast<size_t>(r) * Q + row) * pw + j]);
    return cs[(static_cast<size_t>(r) * Q + row) * k + j];
  };
  auto entry_idx = [&](int e) -> int64_t {
    const int r = e / k, j = e - r * k;
    if (PACKED) {
      const int col = static_cast<int>(packed[(static_cast<size_t>(r) * Q + row) * pw + k + j]);
      return col >= 0 ? static_cast<int64_t>(col) + row_offset[r] : -1;
    }
    return ci[(static_cast<size_t>(r) * Q + row) * k + j];
  };
  // keys: invalid entries (index < 0) get key 0 (below every real score incl. -inf)
  int valid = 0;
  for (int e = lane; e < n; e += 32) {
    const bool ok = entry_idx(e) >= 0;
    keys[e] = ok ? f2key(__float_as_uint(entry_score(e))) : 0u;
    valid += ok ? 1 : 0;
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  __syncwarp();
  const int want = min(valid, k);
  uint32_t T = 1u;            // every valid key is >= kKeyNegInf > 1
  int n_gt = 0;
  int64_t I = INT64_MAX;      // ties with T are taken while index <= I
  // up to kMaxCand valid entries (the usual case after a candidate exchange: only the
  // k + margin candidates above the global threshold were re-ranked at all) go straight to
  // the sort; otherwise the k best are selected first
  const bool need_select = valid > kMaxCand;
  if (need_select) {
    T = warp_kth_largest(keys, n, k, hist);
    int c = 0, ceq = 0;
    for (int e = lane; e < n; e += 32) {
      c += (keys[e] > T) ? 1 : 0;
      ceq += (keys[e] == T) ? 1 : 0;
    }
    n_gt = __reduce_add_sync(0xffffffffu, c);
    const int n_eq = __reduce_add_sync(0xffffffffu, ceq);
    const int room = k - n_gt;
    if (n_eq > room) {
      // the `room` lowest indices among the ties: smallest I with count(idx <= I) >= room
      uint64_t lo = 0;   // bitwise search for the room-th smallest index (indices are >= 0)
      for (int bit = 62; bit >= 0; --bit) {
        const uint64_t cand = lo | (1ull << bit);
        int cc = 0;   // ties with index < cand
        for (int e = lane; e < n; e += 32)
          if (keys[e] == T && static_cast<uint64_t>(entry_idx(e)) < cand) ++cc;
        cc = __reduce_add_sync(0xffffffffu, cc);
        if (cc < room) lo = cand;     // fewer than room ties lie below cand: the answer is >= cand
      }
      I = static_cast<int64_t>(lo);
    }
  } else {
    T = 1u;
  }
  // collect the survivors
  const uint32_t lt_mask = (1u << lane) - 1u;
  int pos = 0;
  for (int b0 = 0; b0 < n; b0 += 32) {
    const int e = b0 + lane;
    bool take = false;
    if (e < n) {
      const uint32_t key = keys[e];
      take = !need_select ? (key != 0u) : (key > T || (key == T && (I == INT64_MAX || entry_idx(e) <= I)));
    }
    const uint32_t bt = __ballot_sync(0xffffffffu, take);
    if (take) {
      const int p = pos + __popc(bt & lt_mask);
      if (p < kMaxCand) {
        sscore[p] = entry_score(e);
        sidx[p] = entry_idx(e);
      }
    }
    pos += __popc(bt);
  }
  for (int j = min(pos, kMaxCand) + lane; j < kMaxCand; j += 32) {
    sscore[j] = -INFINITY;
    sidx[j] = INT64_MAX;
  }
  __syncwarp();
  // warp-synchronous bitonic sort of kMaxCand = 128 entries (64 compare-exchanges per stage)
  for (int size = 2; size <= kMaxCand; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < kMaxCand / 2; t += 32) {
        const int lo_i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi_i = lo_i | stride;
        const bool up = (lo_i & size) == 0;
        const float sa = sscore[lo_i], sb = sscore[hi_i];
        const int64_t ia = sidx[lo_i], ib = sidx[hi_i];
        if (merge_before(sa, ia, sb, ib) != up) {
          sscore[lo_i] = sb; sscore[hi_i] = sa;
          sidx[lo_i] = ib; sidx[hi_i] = ia;
        }
      }
      __syncwarp();
    }
  }
  for (int j = lane; j < k; j += 32) {
    const bool ok = j < want;
    out_scores[row * k + j] = ok ? sscore[j] : -INFINITY;
    out_idx[row * k + j] = ok ? sidx[j] : -1;
  }
  if (unc_count != nullptr && lane == 0) {
    const float t_min = thr[row];
    bool ok = true;
    if (t_min > -INFINITY) {
      double s2 = 0.0, cnt = 0.0;
      for (int r = 0; r < R; ++r) {
        float2 st;
        if (PACKED) {
          const uint32_t* w = packed + (static_cast<size_t>(r) * Q + row) * pw + 2 * k;
          st = make_float2(__uint_as_float(w[0]), __uint_as_float(w[1]));
        } else {
          st = stat[static_cast<size_t>(r) * Q + row];
        }
        s2 += static_cast<double>(st.x);
        cnt += static_cast<double>(st.y);
      }
      const bool have_k = want >= k;
      const double kth = static_cast<double>(sscore[k - 1]);
      const double sigma = sqrt(s2 / (cnt > 0.0 ? cnt : 1.0));
      const double floor_ = 4e-7 * fmax(fabs(kth), fabs(static_cast<double>(t_min))) + 1e-30;
      ok = have_k && (kth - static_cast<double>(t_min) > static_cast<double>(cert_z) * sigma + floor_);
    }
    if (!ok) unc_rows[atomicAdd(unc_count, 1)] = static_cast<int>(row);
  }
}

// ------------------------------------------------------------------ host side
struct SearchPlan {
  int m_blocks, n_tiles, k_blocks, n_groups, grid;
  bool pair;   // CTA-pair screen kernel (256 x 256 tiles): m-blocks are scheduled two at a time
  int64_t ldq;
  size_t off_qbf16, off_cta_buf, off_gthr, off_pool, off_pool_cnt, off_progress, off_seed, total;
  bool seed;   // thresholds seeded from the first kSeedRows database rows
};

// Pick the number of n-groups.  Cost model of a decomposition (in n-tile times): the segments are
// handed out round-robin and the waves run in step (wave barrier), so the screen lasts
//     waves x (tiles of the longest segment + kSegOverheadTiles),
// the overhead being what every segment pays before its MMA pipe runs full: the wave gate, the
// refill of the operand ring and a threshold warm-up (compactions).  Fitted on the mining
// configuration (16 384 x 16 384: 64 row-block pairs x 64 n-tiles on 74 CTA pairs), where ONE group
// on 64 of the 74 pairs (0.99 ms) beats 8 groups in 7 full waves (1.08 ms) and 22 groups (1.41 ms):
// 2.7 tiles per segment.  For the search (hundreds of tiles per segment) the term is immaterial and
// the choice is the one that fills whole waves, as before.
constexpr double kSegOverheadTiles = 3.0;

static int pick_n_groups(int m_blocks, int n_tiles, int grid) {
  int hi = ((grid + m_blocks - 1) / m_blocks) * 8 + 8;
  if (hi > n_tiles) hi = n_tiles;
  if (hi < 1) hi = 1;
  auto efficiency = [&](int ng) {
    const long long segs = static_cast<long long>(m_blocks) * ng;
    const long long waves = (segs + grid - 1) / grid;
    const double longest = static_cast<double>((n_tiles + ng - 1) / ng) + kSegOverheadTiles;
    const double ideal = static_cast<double>(m_blocks) * n_tiles / grid;
    return ideal / (static_cast<double>(waves) * longest);
  };
  double best_eff = -1.0;
  for (int ng = 1; ng <= hi; ++ng) best_eff = efficiency(ng) > best_eff ? efficiency(ng) : best_eff;
  // the FEWEST groups within 1 % of the best: longer segments mean fewer wave barriers, fewer
  // groups straddling a wave boundary (each straddle streams the group's database range from HBM
  // again) and a smaller candidate pool
  for (int ng = 1; ng <= hi; ++ng)
    if (efficiency(ng) >= best_eff - 0.01) return ng;
  return 1;
}

// ISB_OPT_SCREEN_PAIR = 0 keeps the single-CTA 128 x 256 kernel (A/B switch; the workspace layout
// follows the plan, so the option must not change between the size query and the call)
static bool screen_pair_enabled() { return option(ISB_OPT_SCREEN_PAIR, 1) != 0; }

static SearchPlan make_search_plan(int64_t Q, int64_t N, int64_t D, int terms = 1, bool allow_pair = true) {
  SearchPlan p;
  p.m_blocks = static_cast<int>((Q + kBM - 1) / kBM);
  p.n_tiles = static_cast<int>((N + kBN - 1) / kBN);
  p.k_blocks = terms * static_cast<int>((D + kBK - 1) / kBK);
  const int sms = device_sm_count();
  const long long tiles = static_cast<long long>(p.m_blocks) * p.n_tiles;
  p.grid = static_cast<int>(tiles < sms ? tiles : sms);
  // pairs pay off once the whole machine is busy with >= 2 row blocks per n-tile
  const int pair_m = (p.m_blocks + 1) / 2;
  p.pair = allow_pair && screen_pair_enabled() && p.m_blocks >= 2 &&
           static_cast<long long>(pair_m) * p.n_tiles >= sms / 2 && sms >= 2;
  if (p.pair) {
    p.grid = (sms / 2) * 2;
    p.n_groups = pick_n_groups(pair_m, p.n_tiles, p.grid / 2);
  } else {
    p.n_groups = pick_n_groups(p.m_blocks, p.n_tiles, p.grid);
  }
  {
    const int g = option(ISB_OPT_SCREEN_GROUPS, 0);   // experiment switch (changes the workspace layout)
    if (g >= 1) p.n_groups = g < p.n_tiles ? g : p.n_tiles;
  }
  p.ldq = static_cast<int64_t>(align_up(static_cast<size_t>(D), 8));
  size_t off = 0;
  p.off_qbf16 = off;    off = align_up(off + static_cast<size_t>(Q) * p.ldq * 2 * (terms == 3 ? 2 : 1), 1024);
  p.off_cta_buf = off;  off = align_up(off + static_cast<size_t>(p.grid) * kBM * kCap * sizeof(uint2), 1024);
  p.off_gthr = off;     off = align_up(off + static_cast<size_t>(p.m_blocks + 1) * kBM * 4, 1024);
  p.off_pool = off;     off = align_up(off + static_cast<size_t>(Q) * p.n_groups * kMaxCand * sizeof(uint2), 1024);
  p.off_pool_cnt = off; off = align_up(off + static_cast<size_t>(Q) * p.n_groups * 4, 1024);
  p.off_progress = off; off = align_up(off + static_cast<size_t>(kMaxWaves) * 4, 1024);
  // Opt-in (ISB_OPT_SCREEN_SEED = 1): measured on a 125k-row shard the screen gets 0.23 ms faster (4.19 ->
  // 3.96 ms) and the sample GEMM + select cost 0.17 ms -- no net gain yet
  p.seed = terms == 1 && N >= 16ll * kSeedRows && Q >= 128 && option(ISB_OPT_SCREEN_SEED, 0) == 1;
  p.off_seed = off;
  if (p.seed) off = align_up(off + static_cast<size_t>(Q) * kSeedRows * 4, 1024);
  p.total = off;
  return p;
}

// Runs the screen (GEMM + streaming top-k) into the candidate pool.  Shared by
// the search and the negative-mining entry points.
// a_lo / b_lo non-null: split operands, scores = a_hi.b_hi + a_lo.b_hi + a_hi.b_lo
// (plan.k_blocks must then be 3 * the k-blocks of one term).
int launch_topk_screen(const uint16_t* a_bf16, const uint16_t* a_lo, int64_t lda, int64_t Q,
                       const uint16_t* b_bf16, const uint16_t* b_lo, int64_t ldb, int64_t N, int64_t D,
                       int kc, const SearchPlan& plan, uint8_t* ws, const int* col_label,
                       const int* row_label, const float* row_ub, float ub_slack, cudaStream_t st) {
  CUtensorMap ta, tb, ta_lo, tb_lo;
  int rc = make_tmap_bf16_k64(&ta, a_bf16, Q, D, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_k64(&tb, b_bf16, N, D, ldb, kBN);
  if (rc) return rc;
  ta_lo = ta;
  tb_lo = tb;
  int kb_per_term = kSingleTerm;
  if (a_lo != nullptr && b_lo != nullptr) {
    rc = make_tmap_bf16_k64(&ta_lo, a_lo, Q, D, lda, kBM);
    if (rc) return rc;
    rc = make_tmap_bf16_k64(&tb_lo, b_lo, N, D, ldb, kBN);
    if (rc) return rc;
    kb_per_term = plan.k_blocks / 3;
  }

  uint32_t* gthr = reinterpret_cast<uint32_t*>(ws + plan.off_gthr);
  const size_t n_thr = static_cast<size_t>(plan.m_blocks + 1) * kBM;
  fill_u32_kernel<<<static_cast<int>((n_thr + 255) / 256), 256, 0, st>>>(gthr, n_thr, kKeyNegInf);

  if (plan.seed && col_label == nullptr && kb_per_term == kSingleTerm && kc <= kSeedRows) {
    float* sample = reinterpret_cast<float*>(ws + plan.off_seed);
    rc = isb_gemm_nt(a_bf16, lda, b_bf16, ldb, Q, kSeedRows, D, nullptr, sample, kSeedRows, 1, nullptr, 0, st);
    if (rc) return rc;
    const size_t smem = static_cast<size_t>(kSelWarps) * (kSeedRows + 256) * 4;
    seed_threshold_kernel<<<static_cast<unsigned>((Q + kSelWarps - 1) / kSelWarps), 32 * kSelWarps, smem, st>>>(
        sample, Q, kSeedRows, kc, gthr);
    ISB_CUDA(cudaGetLastError());
  }
  // wave barrier of the scheduler: ISB_OPT_SCREEN_WAVESYNC = 0 switches it off
  int* progress = nullptr;
  int window = 0;
  if (option(ISB_OPT_SCREEN_WAVESYNC, 1) != 0 && plan.grid >= device_sm_count() / 2) {
    progress = reinterpret_cast<int*>(ws + plan.off_progress);
    ISB_CUDA(cudaMemsetAsync(progress, 0, static_cast<size_t>(kMaxWaves) * 4, st));
  }
  TopkSched sched{plan.m_blocks, plan.n_tiles, plan.n_groups, plan.k_blocks, progress, window};
  TopkEpiParams ep;
  ep.Q = static_cast<int>(Q);
  ep.N = static_cast<int>(N);
  ep.n_groups = plan.n_groups;
  ep.kc = kc;
  ep.cta_buf = reinterpret_cast<uint2*>(ws + plan.off_cta_buf);
  ep.gthr = gthr;
  ep.pool = reinterpret_cast<uint2*>(ws + plan.off_pool);
  ep.pool_cnt = reinterpret_cast<int*>(ws + plan.off_pool_cnt);
  ep.col_label = col_label;
  ep.row_label = row_label;
  ep.row_ub = row_ub;
  ep.ub_slack = ub_slack;

  if (plan.pair) {
    // pair kernel: every CTA loads its own half (128 rows) of a 256-row B tile
    rc = make_tmap_bf16_k64(&tb, b_bf16, N, D, ldb, kPairBRows);
    if (rc) return rc;
    tb_lo = tb;
    if (kb_per_term != kSingleTerm) {
      rc = make_tmap_bf16_k64(&tb_lo, b_lo, N, D, ldb, kPairBRows);
      if (rc) return rc;
    }
    sched.m_blocks = (plan.m_blocks + 1) / 2;   // 256-row blocks
    if (col_label != nullptr) {
      auto kern = gemm_tc_pair_kernel<TopkSched, TopkEpilogue<true>>;
      ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
      kern<<<plan.grid, kGemmThreads, kPairSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, kb_per_term, sched, ep);
    } else {
      auto kern = gemm_tc_pair_kernel<TopkSched, TopkEpilogue<false>>;
      ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
      kern<<<plan.grid, kGemmThreads, kPairSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, kb_per_term, sched, ep);
    }
  } else if (col_label != nullptr) {
    auto kern = gemm_tc_kernel<TopkSched, TopkEpilogue<true>>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    kern<<<plan.grid, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, kb_per_term, sched, ep);
  } else {
    auto kern = gemm_tc_kernel<TopkSched, TopkEpilogue<false>>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    kern<<<plan.grid, kGemmThreads, kGemmSmemBytes, st>>>(ta, tb, ta_lo, tb_lo, kb_per_term, sched, ep);
  }
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

}  // namespace isb

using namespace isb;

extern "C" int isb_f32_to_bf16(const float* x, int64_t rows, int64_t cols, int64_t ldx, uint16_t* y,
                               int64_t ldy, int part, void* stream) {
  ISB_CHECK_ARG(x && y, "isb_f32_to_bf16: null pointer");
  ISB_CHECK_ARG(rows >= 0 && cols >= 0 && ldx >= cols && ldy >= cols, "isb_f32_to_bf16: bad shape");
  ISB_CHECK_ARG(ldy % 8 == 0, "isb_f32_to_bf16: ldy (%lld) must be a multiple of 8", (long long)ldy);
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "isb_f32_to_bf16: y must be 16-byte aligned");
  ISB_CHECK_ARG(part >= 0 && part <= 2, "isb_f32_to_bf16: part must be 0, 1 or 2");
  if (rows == 0 || ldy == 0) return ISB_OK;
  const int64_t total = rows * (ldy / 8);
  const int64_t blocks = (total + 255) / 256;
  const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
  f32_to_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, cols, ldx, y, ldy, part);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" size_t isb_topk_search_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k, int margin) {
  (void)k; (void)margin;
  if (Q <= 0 || N <= 0 || D <= 0) return 0;
  return make_search_plan(Q, N, D).total + 1024;  // slack: the base is aligned up to 1024
}

static int check_search_args(const char* fn, int64_t Q, int64_t N, int64_t D, int k, int margin) {
  ISB_CHECK_ARG(Q > 0 && N > 0 && D > 0, "%s: empty problem (Q=%lld N=%lld D=%lld)", fn, (long long)Q,
                (long long)N, (long long)D);
  ISB_CHECK_ARG(N < (1ll << 31) && Q < (1ll << 31), "%s: Q and N must be < 2^31 per call", fn);
  ISB_CHECK_ARG(D % 8 == 0, "%s: D (%lld) must be a multiple of 8 (pad with zeros)", fn, (long long)D);
  ISB_CHECK_ARG(k >= 1 && margin >= 0 && k + margin <= ISB_MAX_CANDIDATES,
                "%s: need 1 <= k, k + margin <= %d (k=%d margin=%d)", fn, ISB_MAX_CANDIDATES, k, margin);
  ISB_CHECK_ARG(k <= N, "%s: k (%d) > N (%lld)", fn, k, (long long)N);
  return isb_check_device();
}

static int carve_workspace(const char* fn, const SearchPlan& plan, void* workspace, size_t workspace_bytes,
                           uint8_t** ws) {
  *ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  if (workspace == nullptr || *ws + plan.total > static_cast<uint8_t*>(workspace) + workspace_bytes) {
    set_error("%s: workspace too small (need %zu bytes incl. alignment slack, got %zu)", fn,
              plan.total + 1024, workspace_bytes);
    return ISB_ERR_WORKSPACE;
  }
  return ISB_OK;
}

extern "C" int isb_topk_screen(const float* q, int64_t Q, const uint16_t* db_bf16, int64_t N, int64_t D,
                               int64_t ld_bf16, int k, int margin, void* workspace,
                               size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(q && db_bf16, "isb_topk_screen: null pointer");
  int rc = check_search_args("isb_topk_screen", Q, N, D, k, margin);
  if (rc) return rc;
  ISB_CHECK_ARG(ld_bf16 >= D && ld_bf16 % 8 == 0, "isb_topk_screen: bad ld_bf16");
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(db_bf16) & 15) == 0,
                "isb_topk_screen: inputs must be 16-byte aligned");
  const SearchPlan plan = make_search_plan(Q, N, D);
  uint8_t* ws;
  rc = carve_workspace("isb_topk_screen", plan, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  uint16_t* q_bf16 = reinterpret_cast<uint16_t*>(ws + plan.off_qbf16);
  rc = isb_f32_to_bf16(q, Q, D, D, q_bf16, plan.ldq, 0, stream);
  if (rc) return rc;
  int kc = k + margin;
  if (kc > N) kc = static_cast<int>(N);
  return launch_topk_screen(q_bf16, nullptr, plan.ldq, Q, db_bf16, nullptr, ld_bf16, N, D, kc, plan, ws,
                            nullptr, nullptr, nullptr, 0.f, static_cast<cudaStream_t>(stream));
}

static int launch_rerank(const char* fn, const float* q, const float* db_f32, int64_t D, int64_t n_rows,
                         const SearchPlan& plan, uint8_t* ws, int kc, int k, int64_t idx_offset,
                         const int* row_map, int plain_screen, int32_t* unc_rows, int32_t* unc_count,
                         float* out_scores, int64_t* out_idx, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(D) * 4;
  ISB_CHECK_ARG(smem <= 160 * 1024, "%s: D too large for the re-rank kernel", fn);
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (unc_count != nullptr) ISB_CUDA(cudaMemsetAsync(unc_count, 0, 4, st));
  rerank_kernel<<<static_cast<unsigned>(n_rows), kRerankThreads, smem, st>>>(
      q, db_f32, static_cast<int>(D), plan.n_groups, reinterpret_cast<const uint2*>(ws + plan.off_pool),
      reinterpret_cast<const int*>(ws + plan.off_pool_cnt), kc, k, idx_offset, row_map, kCertZ, plain_screen,
      unc_rows, unc_count, out_scores, out_idx);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_topk_rerank(const float* q, int64_t Q, const float* db_f32, int64_t N, int64_t D, int k,
                               int margin, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                               int32_t* uncertified_rows, int32_t* n_uncertified, void* workspace,
                               size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(q && db_f32 && out_scores && out_idx, "isb_topk_rerank: null pointer");
  ISB_CHECK_ARG((uncertified_rows == nullptr) == (n_uncertified == nullptr),
                "isb_topk_rerank: uncertified_rows and n_uncertified go together");
  int rc = check_search_args("isb_topk_rerank", Q, N, D, k, margin);
  if (rc) return rc;
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(db_f32) & 15) == 0,
                "isb_topk_rerank: inputs must be 16-byte aligned");
  const SearchPlan plan = make_search_plan(Q, N, D);
  uint8_t* ws;
  rc = carve_workspace("isb_topk_rerank", plan, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  int kc = k + margin;
  if (kc > N) kc = static_cast<int>(N);
  return launch_rerank("isb_topk_rerank", q, db_f32, D, Q, plan, ws, kc, k, idx_offset, nullptr, 1,
                       uncertified_rows, n_uncertified, out_scores, out_idx, static_cast<cudaStream_t>(stream));
}

extern "C" int isb_topk_search(const float* q, int64_t Q, const float* db_f32, const uint16_t* db_bf16,
                               int64_t N, int64_t D, int64_t ld_bf16, int k, int margin,
                               int64_t idx_offset, float* out_scores, int64_t* out_idx,
                               int32_t* uncertified_rows, int32_t* n_uncertified, void* workspace,
                               size_t workspace_bytes, void* stream) {
  int rc = isb_topk_screen(q, Q, db_bf16, N, D, ld_bf16, k, margin, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return isb_topk_rerank(q, Q, db_f32, N, D, k, margin, idx_offset, out_scores, out_idx, uncertified_rows,
                         n_uncertified, workspace, workspace_bytes, stream);
}

// ---- second line: fp32-grade (split-operand) re-screen of the listed rows
extern "C" size_t isb_topk_resolve_workspace_bytes(int64_t n_rows, int64_t N, int64_t D) {
  if (n_rows <= 0 || N <= 0 || D <= 0) return 0;
  return make_search_plan(n_rows, N, D, 3).total + 1024;
}

extern "C" int isb_topk_resolve(const float* q, const float* db_f32, const uint16_t* db_bf16,
                                const uint16_t* db_lo_bf16, int64_t N, int64_t D, int64_t ld_bf16, int k,
                                int margin, int64_t idx_offset, const int32_t* rows, int64_t n_rows,
                                float* out_scores, int64_t* out_idx, int32_t* uncertified_rows,
                                int32_t* n_uncertified, void* workspace, size_t workspace_bytes,
                                void* stream) {
  ISB_CHECK_ARG(q && db_f32 && db_bf16 && db_lo_bf16 && rows && out_scores && out_idx && uncertified_rows &&
                n_uncertified, "isb_topk_resolve: null pointer");
  int rc = check_search_args("isb_topk_resolve", n_rows, N, D, k, margin);
  if (rc) return rc;
  ISB_CHECK_ARG(ld_bf16 >= D && ld_bf16 % 8 == 0, "isb_topk_resolve: bad ld_bf16");
  const SearchPlan plan = make_search_plan(n_rows, N, D, 3);
  uint8_t* ws;
  rc = carve_workspace("isb_topk_resolve", plan, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint16_t* a_hi = reinterpret_cast<uint16_t*>(ws + plan.off_qbf16);
  uint16_t* a_lo = a_hi + static_cast<size_t>(n_rows) * plan.ldq;
  gather_q_terms_kernel<<<static_cast<unsigned>(n_rows), 256, 0, st>>>(q, rows, (int)D, plan.ldq, a_hi, a_lo);
  ISB_CUDA(cudaGetLastError());
  int kc = k + margin;
  if (kc > N) kc = static_cast<int>(N);
  rc = launch_topk_screen(a_hi, a_lo, plan.ldq, n_rows, db_bf16, db_lo_bf16, ld_bf16, N, D, kc, plan, ws,
                          nullptr, nullptr, nullptr, 0.f, st);
  if (rc) return rc;
  return launch_rerank("isb_topk_resolve", q, db_f32, D, n_rows, plan, ws, kc, k, idx_offset, rows, 0,
                       uncertified_rows, n_uncertified, out_scores, out_idx, st);
}

// ---- last line: exhaustive exact search of the listed rows
static int exhaustive_chunks(int64_t n_rows, int64_t N) {
  long long c = (148ll * 4 + n_rows - 1) / n_rows;
  if (c > 32) c = 32;
  if (c > N) c = N;
  if (c < 1) c = 1;
  return static_cast<int>(c);
}

extern "C" size_t isb_topk_exhaustive_workspace_bytes(int64_t n_rows, int64_t N, int k) {
  if (n_rows <= 0 || N <= 0 || k <= 0) return 0;
  return align_up(static_cast<size_t>(n_rows) * exhaustive_chunks(n_rows, N) * k * sizeof(ExhEntry), 1024) + 1024;
}

extern "C" int isb_topk_exhaustive(const float* q, const float* db_f32, int64_t N, int64_t D, int k,
                                   int64_t idx_offset, const int32_t* rows, int64_t n_rows,
                                   float* out_scores, int64_t* out_idx, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(q && db_f32 && rows && out_scores && out_idx, "isb_topk_exhaustive: null pointer");
  int rc = check_search_args("isb_topk_exhaustive", n_rows, N, D, k, 0);
  if (rc) return rc;
  const size_t need = isb_topk_exhaustive_workspace_bytes(n_rows, N, k);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  if (workspace == nullptr || ws + need - 1024 > static_cast<uint8_t*>(workspace) + workspace_bytes) {
    set_error("isb_topk_exhaustive: workspace too small (need %zu bytes, got %zu)", need, workspace_bytes);
    return ISB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = exhaustive_chunks(n_rows, N);
  const size_t smem = static_cast<size_t>(D) * 4;
  ISB_CHECK_ARG(smem <= 160 * 1024, "isb_topk_exhaustive: D too large");
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(exhaustive_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ExhEntry* part = reinterpret_cast<ExhEntry*>(ws);
  exhaustive_chunk_kernel<<<dim3(chunks, static_cast<unsigned>(n_rows)), kRerankThreads, smem, st>>>(
      q, db_f32, (int)N, (int)D, k, rows, part);
  ISB_CUDA(cudaGetLastError());
  int n_pow2 = 2;
  while (n_pow2 < chunks * k) n_pow2 <<= 1;
  const size_t msmem = static_cast<size_t>(n_pow2) * 12;
  if (msmem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(exhaustive_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  exhaustive_merge_kernel<<<static_cast<unsigned>(n_rows), 256, msmem, st>>>(part, chunks, k, n_pow2, rows,
                                                                            idx_offset, out_scores, out_idx);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

static int launch_merge(const float* cand_scores, const int64_t* cand_idx, const uint32_t* packed,
                        const int64_t* row_offset, int R, int64_t Q, int k, float* out_scores,
                        int64_t* out_idx, const float2* stat, const float* thr, int32_t* unc_rows,
                        int32_t* unc_count, cudaStream_t st) {
  if (k <= kMaxCand) {
    const size_t smem = kSelWarps * (align_up_dev(static_cast<size_t>(R) * k * 4, 8) + kMaxCand * 12 + 1024);
    const unsigned grid = static_cast<unsigned>((Q + kSelWarps - 1) / kSelWarps);
    if (packed != nullptr) {
      auto kern = topk_merge_warp_kernel<true>;
      if (smem > 48 * 1024)
        ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 32 * kSelWarps, smem, st>>>(nullptr, nullptr, packed, row_offset, R, Q, k, out_scores, out_idx,
                                               nullptr, thr, kCertZ, unc_rows, unc_count);
    } else {
      auto kern = topk_merge_warp_kernel<false>;
      if (smem > 48 * 1024)
        ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 32 * kSelWarps, smem, st>>>(cand_scores, cand_idx, nullptr, nullptr, R, Q, k, out_scores,
                                               out_idx, stat, thr, kCertZ, unc_rows, unc_count);
    }
  } else {
    ISB_CHECK_ARG(packed == nullptr, "isb_topk_merge: k > %d is not supported with packed lists", kMaxCand);
    int n_pow2 = 2;
    while (n_pow2 < R * k) n_pow2 <<= 1;
    const size_t smem = static_cast<size_t>(n_pow2) * 12;
    if (smem > 48 * 1024)
      ISB_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<static_cast<unsigned>(Q), 256, smem, st>>>(cand_scores, cand_idx, R, Q, k, n_pow2,
                                                                  out_scores, out_idx, stat, thr, kCertZ,
                                                                  unc_rows, unc_count);
  }
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_topk_merge(const float* cand_scores, const int64_t* cand_idx, int R, int64_t Q, int k,
                              float* out_scores, int64_t* out_idx, void* stream) {
  ISB_CHECK_ARG(cand_scores && cand_idx && out_scores && out_idx, "isb_topk_merge: null pointer");
  ISB_CHECK_ARG(R >= 1 && Q >= 0 && k >= 1, "isb_topk_merge: bad shape");
  ISB_CHECK_ARG(static_cast<int64_t>(R) * k <= kMergeMax, "isb_topk_merge: R*k (%lld) > %d",
                (long long)R * k, kMergeMax);
  if (Q == 0) return ISB_OK;
  return launch_merge(cand_scores, cand_idx, nullptr, nullptr, R, Q, k, out_scores, out_idx, nullptr, nullptr,
                      nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

// ---- sharded search with candidate exchange (three local stages around two all-gathers)
extern "C" int isb_topk_candidates(int64_t Q, int64_t N, int64_t D, int k, int margin, int kc_out,
                                   float* cand_screen, int32_t* cand_col, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  ISB_CHECK_ARG(cand_screen && cand_col, "isb_topk_candidates: null pointer");
  int rc = check_search_args("isb_topk_candidates", Q, N, D, k, margin);
  if (rc) return rc;
  const SearchPlan plan = make_search_plan(Q, N, D);
  uint8_t* ws;
  rc = carve_workspace("isb_topk_candidates", plan, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  int kc = k + margin;   // what isb_topk_screen kept per row (a shard shorter than that keeps all its rows)
  if (kc > N) kc = static_cast<int>(N);
  ISB_CHECK_ARG(kc_out >= kc && kc_out <= ISB_MAX_CANDIDATES, "isb_topk_candidates: need %d <= kc_out <= %d", kc,
                ISB_MAX_CANDIDATES);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint2* pool = reinterpret_cast<const uint2*>(ws + plan.off_pool);
  const int* pool_cnt = reinterpret_cast<const int*>(ws + plan.off_pool_cnt);
  const size_t per_warp = (static_cast<size_t>(plan.n_groups) * kMaxCand * 2 + 256 + plan.n_groups + 1) * 4;
  int wpc = kSelWarps;
  while (wpc > 1 && per_warp * wpc > 96 * 1024) wpc >>= 1;
  if (per_warp * wpc <= 200 * 1024) {
    const size_t smem = per_warp * wpc;
    if (smem > 48 * 1024)
      ISB_CUDA(cudaFuncSetAttribute(pool_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pool_candidates_kernel<<<static_cast<unsigned>((Q + wpc - 1) / wpc), 32 * wpc, smem, st>>>(
        Q, plan.n_groups, pool, pool_cnt, kc, kc_out, cand_screen, cand_col);
  } else {
    pool_candidates_cta_kernel<<<static_cast<unsigned>(Q), kRerankThreads, 0, st>>>(
        plan.n_groups, pool, pool_cnt, kc, kc_out, cand_screen, cand_col);
  }
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_topk_global_threshold(const float* all_screen, int R, int64_t Q, int kc, int kth, float* thr,
                                         void* stream) {
  ISB_CHECK_ARG(all_screen && thr, "isb_topk_global_threshold: null pointer");
  ISB_CHECK_ARG(R >= 1 && Q >= 0 && kc >= 1 && kc <= ISB_MAX_CANDIDATES, "isb_topk_global_threshold: bad shape");
  if (kth <= 0) kth = kc;
  ISB_CHECK_ARG(kth >= kc && kth <= ISB_MAX_CANDIDATES, "isb_topk_global_threshold: need kc <= kth <= %d",
                ISB_MAX_CANDIDATES);
  if (Q == 0) return ISB_OK;
  const size_t smem = static_cast<size_t>(kSelWarps) * (static_cast<size_t>(R) * kc + 256) * 4;
  ISB_CHECK_ARG(smem <= 200 * 1024, "isb_topk_global_threshold: R * kc (%d) too large", R * kc);
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(global_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  global_threshold_kernel<<<static_cast<unsigned>((Q + kSelWarps - 1) / kSelWarps), 32 * kSelWarps, smem,
                            static_cast<cudaStream_t>(stream)>>>(all_screen, R, Q, kc, kth, thr);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_topk_rerank_owned(const float* q, int64_t Q, const float* db_f32, int64_t N, int64_t D,
                                     int k, int kc, const float* cand_screen, const int32_t* cand_col,
                                     const float* thr, uint32_t* packed, void* stream) {
  ISB_CHECK_ARG(q && db_f32 && cand_screen && cand_col && thr && packed, "isb_topk_rerank_owned: null pointer");
  ISB_CHECK_ARG(Q >= 0 && N > 0 && D > 0 && D % 8 == 0, "isb_topk_rerank_owned: bad shape");
  ISB_CHECK_ARG(k >= 1 && k <= ISB_MAX_CANDIDATES && kc >= 1 && kc <= ISB_MAX_CANDIDATES,
                "isb_topk_rerank_owned: need 1 <= k, kc <= %d", ISB_MAX_CANDIDATES);
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(db_f32) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(packed) & 3) == 0, "isb_topk_rerank_owned: misaligned input");
  if (Q == 0) return ISB_OK;
  const size_t smem = static_cast<size_t>(D) * 4;
  ISB_CHECK_ARG(smem <= 160 * 1024, "isb_topk_rerank_owned: D too large for the re-rank kernel");
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(rerank_owned_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rerank_owned_kernel<<<static_cast<unsigned>(Q), kRerankThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      q, db_f32, static_cast<int>(D), kc, k, cand_screen, cand_col, thr, packed);
  ISB_CUDA(cudaGetLastError());
  return ISB_OK;
}

extern "C" int isb_topk_merge_certified(const uint32_t* packed_all, const int64_t* row_offsets,
                                        const float* thr, int R, int64_t Q, int k, float* out_scores,
                                        int64_t* out_idx, int32_t* uncertified_rows, int32_t* n_uncertified,
                                        void* stream) {
  ISB_CHECK_ARG(packed_all && row_offsets && thr && out_scores && out_idx && uncertified_rows && n_uncertified,
                "isb_topk_merge_certified: null pointer");
  ISB_CHECK_ARG(R >= 1 && Q >= 0 && k >= 1 && k <= ISB_MAX_CANDIDATES, "isb_topk_merge_certified: bad shape");
  ISB_CHECK_ARG(static_cast<int64_t>(R) * k <= kMergeMax, "isb_topk_merge_certified: R*k (%lld) > %d",
                (long long)R * k, kMergeMax);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ISB_CUDA(cudaMemsetAsync(n_uncertified, 0, 4, st));
  if (Q == 0) return ISB_OK;
  return launch_merge(nullptr, nullptr, packed_all, row_offsets, R, Q, k, out_scores, out_idx, nullptr, thr,
                      uncertified_rows, n_uncertified, st);
}

// ------------------------------------------------------------------ a13 entry point
struct MiningPlan {
  SearchPlan sp;
  size_t off_row_label, off_pos64, off_pos32, off_flag, off_cand_screen, off_cand_col, total;
};

static MiningPlan make_mining_plan(int64_t P, int64_t N, int64_t D, int split) {
  MiningPlan m;
  m.sp = make_search_plan(P, N, D, split ? 3 : 1);
  size_t off = m.sp.total;
  m.off_row_label = off; off = align_up(off + static_cast<size_t>(P) * 4, 1024);
  m.off_pos64 = off;     off = align_up(off + static_cast<size_t>(P) * 8, 1024);
  m.off_pos32 = off;     off = align_up(off + static_cast<size_t>(P) * 4, 1024);
  m.off_flag = off;      off = align_up(off + static_cast<size_t>(P) * 4, 1024);
  m.off_cand_screen = off; off = align_up(off + static_cast<size_t>(P) * 32 * 4, 1024);   // warp path: kc <= 32
  m.off_cand_col = off;    off = align_up(off + static_cast<size_t>(P) * 32 * 4, 1024);
  m.total = off;
  return m;
}

constexpr int kMiningCand = 16;   // candidates re-checked exactly per couple (<= kMaxCand; 32: +0.2 ms, 8: -0.1 ms)

extern "C" size_t isb_select_negatives_workspace_bytes(int64_t P, int64_t N, int64_t D, int split) {
  if (P <= 0 || N <= 0 || D <= 0) return 0;
  return make_mining_plan(P, N, D, split).total + 1024;
}

extern "C" int isb_select_negatives(const float* emb, const uint16_t* emb_hi, const uint16_t* emb_lo,
                                    int64_t ld, int64_t N, int64_t D, const int32_t* label,
                                    const int64_t* anchors, const int64_t* positives, int64_t P,
                                    int semi_hard, float screen_eps, float sigma_floor, int64_t* neg_idx,
                                    float* neg_sim, float* pos_sim, int32_t* uncertified_rows,
                                    int32_t* n_uncertified, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  ISB_CHECK_ARG(emb && emb_hi && label && anchors && positives && neg_idx && neg_sim && pos_sim,
                "isb_select_negatives: null pointer");
  ISB_CHECK_ARG(P > 0 && N > 0 && D > 0 && N < (1ll << 31) && P < (1ll << 31), "isb_select_negatives: bad shape");
  ISB_CHECK_ARG(uncertified_rows == nullptr || n_uncertified != nullptr,
                "isb_select_negatives: uncertified_rows needs n_uncertified");
  ISB_CHECK_ARG(screen_eps >= 0.f && sigma_floor >= 0.f, "isb_select_negatives: negative error bound");
  ISB_CHECK_ARG(D % 8 == 0 && ld % 8 == 0 && ld >= D, "isb_select_negatives: D and ld must be multiples of 8, ld >= D");
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(emb) & 15) == 0 && (reinterpret_cast<uintptr_t>(emb_hi) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(emb_lo) & 15) == 0, "isb_select_negatives: inputs must be 16-byte aligned");
  int rc = isb_check_device();
  if (rc) return rc;
  const int split = emb_lo != nullptr ? 1 : 0;
  const MiningPlan mp = make_mining_plan(P, N, D, split);
  ISB_CHECK_ARG(mp.sp.ldq == ld, "isb_select_negatives: ld (%lld) must equal D rounded up to 8", (long long)ld);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  if (workspace == nullptr || ws + mp.total > static_cast<uint8_t*>(workspace) + workspace_bytes) {
    set_error("isb_select_negatives: workspace too small (need %zu bytes, got %zu)", mp.total + 1024, workspace_bytes);
    return ISB_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint16_t* a_hi = reinterpret_cast<uint16_t*>(ws + mp.sp.off_qbf16);   // [P, ld] anchors' rows
  uint16_t* a_lo = split ? a_hi + static_cast<size_t>(P) * ld : nullptr;
  int* row_label = reinterpret_cast<int*>(ws + mp.off_row_label);
  double* pos64 = reinterpret_cast<double*>(ws + mp.off_pos64);
  float* pos32 = reinterpret_cast<float*>(ws + mp.off_pos32);
  int* flag = reinterpret_cast<int*>(ws + mp.off_flag);

  gather_rows_kernel<<<static_cast<unsigned>(P), 128, 0, st>>>(emb_hi, ld, anchors, a_hi, label, row_label);
  ISB_CUDA(cudaGetLastError());
  if (split) {
    gather_rows_kernel<<<static_cast<unsigned>(P), 128, 0, st>>>(emb_lo, ld, anchors, a_lo, label, row_label);
    ISB_CUDA(cudaGetLastError());
  }
  pair_dot_kernel<<<static_cast<unsigned>((P * 32 + 255) / 256), 256, 0, st>>>(emb, (int)D, anchors, positives, P, pos64, pos32);
  ISB_CUDA(cudaGetLastError());
  ISB_CUDA(cudaMemcpyAsync(pos_sim, pos32, static_cast<size_t>(P) * 4, cudaMemcpyDeviceToDevice, st));
  if (n_uncertified != nullptr) ISB_CUDA(cudaMemsetAsync(n_uncertified, 0, 4, st));

  // One negative per couple is wanted: a short candidate list is enough -- the certificate
  // (worst candidate's screen score + screen_eps < the exact winner) sends the rows it is not
  // enough for to the brute-force pass.  The exact re-check gathers kc rows of D floats per
  // couple: 128 candidates were 17 GB of random 8 KB reads at the 16k x 2048 configuration.
  int kc = option(ISB_OPT_MINING_KC, kMiningCand);
  if (kc < 1 || kc > kMaxCand) kc = kMiningCand;
  if (kc > N) kc = static_cast<int>(N);
  // semi-hard: columns scoring >= sim_pos (+ the screen's error bound) are masked in the epilogue
  // four times the certificate's z: the noise measured on a couple's ~16 best candidates scatters
  // around the dense-row expectation (up to 2.2 x on clustered descriptors), and a couple is
  // rejected when its measured noise exceeds the slack
  const float ub_slack = fmaxf(screen_eps, 4.f * kCertZ * sigma_floor);
  rc = launch_topk_screen(a_hi, a_lo, ld, P, emb_hi, emb_lo, ld, N, D, kc, mp.sp, ws, label, row_label,
                          semi_hard ? pos32 : nullptr, ub_slack, st);
  if (rc) return rc;
  const size_t smem = static_cast<size_t>(D) * 4;
  ISB_CHECK_ARG(smem <= 160 * 1024, "isb_select_negatives: D too large");
  if (smem > 48 * 1024)
    ISB_CUDA(cudaFuncSetAttribute(mining_bruteforce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int progressive = option(ISB_OPT_MINING_PROGRESSIVE, 1) != 0 ? 1 : 0;
  const uint2* pool = reinterpret_cast<const uint2*>(ws + mp.sp.off_pool);
  const int* pool_cnt = reinterpret_cast<const int*>(ws + mp.sp.off_pool_cnt);
  // warp path (kc <= 32 and a row's pool fits a warp's share of shared memory): warp-level candidate
  // selection, then one warp per couple; otherwise one CTA per couple
  const size_t per_warp = (static_cast<size_t>(mp.sp.n_groups) * kMaxCand * 2 + 256 + mp.sp.n_groups + 1) * 4;
  int wpc = kSelWarps;
  while (wpc > 1 && per_warp * wpc > 96 * 1024) wpc >>= 1;
  if (kc <= 32 && per_warp * wpc <= 200 * 1024 && option(ISB_OPT_MINING_PROGRESSIVE, 1) != 2) {
    float* cand_screen = reinterpret_cast<float*>(ws + mp.off_cand_screen);
    int* cand_col = reinterpret_cast<int*>(ws + mp.off_cand_col);
    const size_t psmem = per_warp * wpc;
    if (psmem > 48 * 1024)
      ISB_CUDA(cudaFuncSetAttribute(pool_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    pool_candidates_kernel<<<static_cast<unsigned>((P + wpc - 1) / wpc), 32 * wpc, psmem, st>>>(
        P, mp.sp.n_groups, pool, pool_cnt, kc, kc, cand_screen, cand_col);
    ISB_CUDA(cudaGetLastError());
    mining_rerank_warp_kernel<<<static_cast<unsigned>((P + kMineWarps - 1) / kMineWarps), 32 * kMineWarps, 0, st>>>(
        emb, (int)D, anchors, P, cand_screen, cand_col, kc, pos64, semi_hard, screen_eps, sigma_floor, ub_slack,
        progressive, neg_idx, neg_sim, flag, uncertified_rows, n_uncertified);
  } else {
    if (smem > 48 * 1024)
      ISB_CUDA(cudaFuncSetAttribute(mining_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mining_rerank_kernel<<<static_cast<unsigned>(P), kRerankThreads, smem, st>>>(
        emb, (int)D, anchors, mp.sp.n_groups, pool, pool_cnt, kc, pos64, semi_hard, screen_eps, sigma_floor,
        ub_slack, progressive, neg_idx, neg_sim, flag, uncertified_rows, n_uncertified, nullptr);
  }
  ISB_CUDA(cudaGetLastError());
  if (uncertified_rows == nullptr) {   // no second line requested: resolve the rejected couples here
    mining_bruteforce_kernel<<<static_cast<unsigned>(P), kRerankThreads, smem, st>>>(
        emb, (int)N, (int)D, label, anchors, pos64, semi_hard, flag, neg_idx, neg_sim, n_uncertified);
    ISB_CUDA(cudaGetLastError());
  }
  return ISB_OK;
}

#ifdef ISB_TIMELINE
// debug builds (make TIMELINE=1): copies the role timeline of the last CTA-pair screen launch,
// [256][8] cycle counters (see isb_gemm_core.cuh), to host memory
extern "C" int isb_debug_timeline(long long* out) {
  ISB_CUDA(cudaDeviceSynchronize());
  ISB_CUDA(cudaMemcpyFromSymbol(out, isb::isb_timeline, sizeof(long long) * 256 * isb::kTimelineSlots));
  return ISB_OK;
}
#endif
