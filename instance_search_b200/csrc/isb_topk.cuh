// Streaming per-row top-k fused into the tcgen05 GEMM epilogue.
//
// Each epilogue thread owns one row (= one TMEM lane = one query).  It keeps a
// threshold `thr` and a private candidate buffer of kCap (score, column) pairs
// in global memory (L2-resident).  A score passes when score > thr; passing is
// rare once thr has warmed up (expected k*ln(N/k) passes per row over N
// columns).  When a row's buffer holds more than kTrig entries after a tile,
// the warp compacts it cooperatively: a 32-step radix select finds the kc-th
// largest key, the kc best are kept, thr becomes that key and is published with
// atomicMax to a global per-row threshold that every CTA working on the same
// rows re-reads at each tile.  At the end of a segment (one m-block x one group
// of n-tiles) the surviving <= kc candidates are copied to the row's slot of a
// candidate pool [row][group][kMaxCand] that the merge / re-rank kernel reads.
#pragma once

#include "isb_gemm_core.cuh"

namespace isb {

constexpr int kCap = 512;        // private buffer entries per row
constexpr int kTrig = 256;       // compact when a row holds more than this after a tile (kTrig + 256 <= kCap)
constexpr int kCompactSlack = 64;   // an in-stream compaction may keep up to kc + this many entries
constexpr int kMaxCand = 128;    // == ISB_MAX_CANDIDATES
constexpr int kMaxWaves = 1024;  // wave-barrier counters of the screen scheduler

// order-preserving float <-> uint32 key (larger float <=> larger key)
__host__ __device__ __forceinline__ uint32_t f2key(uint32_t bits) {
  return bits ^ ((bits & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t key2f(uint32_t key) {
  return key ^ ((key & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu);
}
constexpr uint32_t kKeyNegInf = 0x007FFFFFu;  // f2key(bits(-inf))

struct TopkEpiParams {
  int Q;                 // valid rows
  int N;                 // valid columns
  int n_groups;          // pool slots per row
  int kc;                // candidates kept per (row, group), <= kMaxCand
  uint2* cta_buf;        // [gridDim.x][128][kCap]  (score bits, column)
  uint32_t* gthr;        // [m_blocks*128] global per-row threshold keys
  uint2* pool;           // [Q][n_groups][kMaxCand]
  int* pool_cnt;         // [Q][n_groups]
  // optional column mask (hard-negative mining): a column is a candidate only if
  // col_label[col] != row_label[row] and score < row_ub[row] + ub_slack
  const int* col_label;  // [N] or null
  const int* row_label;  // [Q] or null
  const float* row_ub;   // [Q] or null
  float ub_slack;
};

// Warp-cooperative: keep the largest entries of buf[0..cnt) (cnt <= kCap), packed to the
// front, and return the threshold key T they all reach.  At least `keep` entries survive and
// at most keep + slack (n_out): the bit-by-bit select stops as soon as the count of keys >= T
// is within `slack` of `keep` -- a threshold only has to be a LOWER bound of the keep-th best
// key, and an early exit saves most of the ~25 select rounds (slack = 0: exactly the keep
// largest, ties with the keep-th key in slot order).  All 32 lanes call it with identical
// arguments.
// The row's entries come in `raw` (lane l holds entries l, l + 32, ...: load_row_entries), so
// that the caller can fetch the next row's while this one is being reduced.
constexpr int kPerLane = kCap / 32;

__device__ __forceinline__ void load_row_entries(const uint2* buf, int cnt, int lane, uint2 (&raw)[kPerLane]) {
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int j = lane + 32 * i;
    raw[i] = (j < cnt) ? buf[j] : make_uint2(0u, 0u);
  }
}

__device__ __forceinline__ uint32_t warp_compact_row(uint2* buf, int cnt, int keep, int slack, int lane,
                                                     const uint2 (&raw)[kPerLane], int& n_out) {
  constexpr int kPer = kPerLane;
  uint32_t key[kPer], col[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const bool valid = lane + 32 * i < cnt;
    key[i] = valid ? f2key(raw[i].x) : 0u;
    col[i] = raw[i].y;
  }
  __syncwarp();  // every load of this row is ordered before every store below
  // The keys of one row share their leading bits (same sign / exponent): start the
  // bit-by-bit select below the common prefix instead of at bit 31.
  const uint32_t k0 = __shfl_sync(0xffffffffu, key[0], 0);  // entry 0 is valid (cnt > 0)
  uint32_t diff = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) diff |= (lane + 32 * i < cnt) ? (key[i] ^ k0) : 0u;
  diff = __reduce_or_sync(0xffffffffu, diff);
  uint32_t T = k0;
  int c_T = cnt;       // count of keys >= T
  bool exact = true;   // T is the keep-th largest key (all select rounds ran)
  if (diff != 0) {
    const int hb = 31 - __clz(diff);
    T = (hb == 31) ? 0u : (k0 & ~((2u << hb) - 1u));
#pragma unroll 1
    for (int b = hb; b >= 0; --b) {
      const uint32_t trial = T | (1u << b);
      int c = 0;
#pragma unroll
      for (int i = 0; i < kPer; ++i) c += (key[i] >= trial) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= keep) {
        T = trial;
        c_T = c;
        if (c <= keep + slack && b > 0) { exact = false; break; }
      }
    }
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (!exact && c_T <= keep + slack) {
    // keep every entry with key >= T (keep <= c_T <= keep + slack of them)
    int pos = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const bool ge = key[i] >= T && (lane + 32 * i < cnt);
      const uint32_t bge = __ballot_sync(0xffffffffu, ge);
      if (ge) buf[pos + __popc(bge & lt_mask)] = make_uint2(key2f(key[i]), col[i]);
      pos += __popc(bge);
    }
    __syncwarp();
    n_out = pos;
    return T;
  }
  int n_gt = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) n_gt += (key[i] > T) ? 1 : 0;
  n_gt = __reduce_add_sync(0xffffffffu, n_gt);
  const int quota_eq = keep - n_gt;
  int pos_gt = 0, pos_eq = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const bool gt = key[i] > T;
    const bool eq = key[i] == T;
    const uint32_t bgt = __ballot_sync(0xffffffffu, gt);
    const uint32_t beq = __ballot_sync(0xffffffffu, eq);
    if (gt) buf[pos_gt + __popc(bgt & lt_mask)] = make_uint2(key2f(key[i]), col[i]);
    if (eq) {
      const int p = pos_eq + __popc(beq & lt_mask);
      if (p < quota_eq) buf[n_gt + p] = make_uint2(key2f(key[i]), col[i]);
    }
    pos_gt += __popc(bgt);
    pos_eq += __popc(beq);
  }
  __syncwarp();
  n_out = keep;
  return T;
}

template <bool kMasked>
struct TopkEpilogue {
  using Params = TopkEpiParams;
  const Params& p;
  const int row_in_tile;
  const int lane;
  uint2* my_buf;      // this row's private buffer
  uint2* warp_buf;    // buffer of lane 0's row (rows of a warp are consecutive)
  int row;            // global row
  bool row_valid;
  int cnt;
  float thr;
  int my_label;
  float my_ub;

  __device__ TopkEpilogue(const Params& p_, int row_in_tile_)
      : p(p_), row_in_tile(row_in_tile_), lane(row_in_tile_ & 31) {
    uint2* cta = p.cta_buf + static_cast<size_t>(blockIdx.x) * kBM * kCap;
    my_buf = cta + static_cast<size_t>(row_in_tile) * kCap;
    warp_buf = cta + static_cast<size_t>(row_in_tile - lane) * kCap;
    row = 0; row_valid = false; cnt = 0; thr = 0.f; my_label = -1; my_ub = 0.f;
  }

  __device__ __forceinline__ void begin_segment(const Segment& seg) {
    row = seg.m_block * kBM + row_in_tile;
    row_valid = row < p.Q;
    cnt = 0;
    thr = row_valid ? __uint_as_float(key2f(kKeyNegInf)) : __uint_as_float(0x7F800000u);
    if (kMasked) {
      my_label = row_valid ? p.row_label[row] : -1;
      my_ub = (row_valid && p.row_ub != nullptr) ? p.row_ub[row] + p.ub_slack
                                                  : __uint_as_float(0x7F800000u);
    }
  }

  // One accumulator value: append (score, column) to the row's buffer when it
  // beats the threshold.  Branch-free predicated PTX -- the epilogue must stay a
  // few KB of straight-line code (an unrolled, branchy version thrashed the
  // instruction cache and made the MMA pipe wait for TMEM).
  // Once the row's threshold has tightened a value beats it about once in a thousand: the
  // values are first tested four at a time (FMNMX3 + FMNMX + FSETP, 0.75 instructions per
  // value) and the per-value predicated append runs only for a group that holds a hit.
  template <bool kRagged>
  __device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], int col0, int rem,
                                             uint2*& wp) {
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
      float m3;
      asm("max.f32 %0, %1, %2, %3;" : "=f"(m3) : "f"(__uint_as_float(v[4 * g4])),
          "f"(__uint_as_float(v[4 * g4 + 1])), "f"(__uint_as_float(v[4 * g4 + 2])));
      if (__builtin_expect(fmaxf(m3, __uint_as_float(v[4 * g4 + 3])) > thr, 0)) scan_group<kRagged>(v, g4, col0, rem, wp);
    }
  }

  template <bool kRagged>
  __device__ __forceinline__ void scan_group(const uint32_t (&v)[32], int g4, int col0, int rem,
                                             uint2*& wp) {
#pragma unroll
    for (int j = 4 * g4; j < 4 * g4 + 4; ++j) {
      const int col = col0 + j;
      if (kMasked) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t.reg .s32 l;\n\t"
            "setp.gt.f32 p, %1, %2;\n\t"
            "setp.lt.and.f32 p, %1, %6, p;\n\t"
            "setp.lt.and.s32 p, %7, %8, p;\n\t"
            "mad.wide.s32 a, %4, 4, %5;\n\t"
            "mov.s32 l, %9;\n\t"
            "@p ld.global.nc.s32 l, [a];\n\t"
            "setp.ne.and.s32 p, l, %9, p;\n\t"
            "@p st.global.v2.b32 [%0], {%3, %4};\n\t"
            "@p add.u64 %0, %0, 8;\n\t}"
            : "+l"(wp)
            : "f"(__uint_as_float(v[j])), "f"(thr), "r"(v[j]), "r"(col), "l"(p.col_label),
              "f"(my_ub), "r"(j), "r"(rem), "r"(my_label)
            : "memory");
      } else if (kRagged) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.gt.f32 p, %1, %2;\n\t"
            "setp.lt.and.s32 p, %5, %6, p;\n\t"
            "@p st.global.v2.b32 [%0], {%3, %4};\n\t"
            "@p add.u64 %0, %0, 8;\n\t}"
            : "+l"(wp)
            : "f"(__uint_as_float(v[j])), "f"(thr), "r"(v[j]), "r"(col), "r"(j), "r"(rem)
            : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.gt.f32 p, %1, %2;\n\t"
            "@p st.global.v2.b32 [%0], {%3, %4};\n\t"
            "@p add.u64 %0, %0, 8;\n\t}"
            : "+l"(wp)
            : "f"(__uint_as_float(v[j])), "f"(thr), "r"(v[j]), "r"(col)
            : "memory");
      }
    }
  }

  // 256 accumulator columns = 4 iterations of two 32-column chunks, TMEM loads
  // double-buffered against the scan of the previous chunk.
  template <bool kRagged>
  __device__ __forceinline__ void scan_tile(uint32_t tmem_acc, int col0, int rem,
                                            uint64_t* tmem_empty_bar, uint2*& wp) {
    uint32_t v0[32], v1[32];
    ptx::tmem_ld_32x32b_x32(tmem_acc, v0);
#pragma unroll 1
    for (int it = 0; it < kBN / 64; ++it) {
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(tmem_acc + it * 64 + 32, v1);
      scan_chunk<kRagged>(v0, col0 + it * 64, rem - it * 64, wp);
      ptx::tmem_ld_wait();
      if (it + 1 < kBN / 64) {
        ptx::tmem_ld_32x32b_x32(tmem_acc + it * 64 + 64, v0);
      } else {
        // the whole accumulator is in registers: hand the TMEM buffer back
        ptx::tc_fence_before();
        ptx::mbar_arrive(tmem_empty_bar);
      }
      scan_chunk<kRagged>(v1, col0 + it * 64 + 32, rem - it * 64 - 32, wp);
    }
  }

  // slack: how many entries beyond `keep` a compacted row may retain (0 = exactly keep).
  // The rows of a warp tend to need compaction at the same tile (they fill at the same rate
  // while their thresholds warm up), and the MMA pipe waits for the whole series: the entries
  // of the next row are fetched (L2 latency) while the current one is being reduced.
  __device__ __forceinline__ void compact_rows(uint32_t need, int keep, int slack) {
    uint2 raw_a[kPerLane], raw_b[kPerLane];
    int r = __ffs(need) - 1;
    need &= need - 1;
    int c = __shfl_sync(0xffffffffu, cnt, r);
    load_row_entries(warp_buf + static_cast<size_t>(r) * kCap, c, lane, raw_a);
    for (;;) {
      // ---- row r sits in raw_a; prefetch the next row into raw_b
      int r2 = -1, c2 = 0;
      if (need) {
        r2 = __ffs(need) - 1;
        need &= need - 1;
        c2 = __shfl_sync(0xffffffffu, cnt, r2);
        load_row_entries(warp_buf + static_cast<size_t>(r2) * kCap, c2, lane, raw_b);
      }
      finish_row(r, c, keep, slack, raw_a);
      if (r2 < 0) break;
      // ---- row r2 sits in raw_b; prefetch the next row into raw_a
      r = -1;
      if (need) {
        r = __ffs(need) - 1;
        need &= need - 1;
        c = __shfl_sync(0xffffffffu, cnt, r);
        load_row_entries(warp_buf + static_cast<size_t>(r) * kCap, c, lane, raw_a);
      }
      finish_row(r2, c2, keep, slack, raw_b);
      if (r < 0) break;
    }
  }

  __device__ __forceinline__ void finish_row(int r, int c, int keep, int slack, const uint2 (&raw)[kPerLane]) {
    int kept;
    const uint32_t T = warp_compact_row(warp_buf + static_cast<size_t>(r) * kCap, c, keep, slack, lane, raw, kept);
    if (lane == r) {
      cnt = kept;
      thr = fmaxf(thr, __uint_as_float(key2f(T)));
      atomicMax(p.gthr + row, T);
    }
  }

  __device__ __forceinline__ void tile(const Segment& seg, int nt, uint32_t tmem_acc,
                                       uint64_t* tmem_empty_bar) {
    if (row_valid) {
      // thresholds published by the other CTAs that work on these rows
      const uint32_t g = *reinterpret_cast<volatile uint32_t*>(p.gthr + row);
      thr = fmaxf(thr, __uint_as_float(key2f(g)));
    }
    const int col0 = nt * kBN;
    const int rem = p.N - col0;  // valid columns of this tile
    uint2* wp = my_buf + cnt;
    if (kMasked || rem < kBN) scan_tile<true>(tmem_acc, col0, rem, tmem_empty_bar, wp);
    else scan_tile<false>(tmem_acc, col0, rem, tmem_empty_bar, wp);
    cnt = static_cast<int>(wp - my_buf);
    __syncwarp();
    const uint32_t need = __ballot_sync(0xffffffffu, cnt > kTrig);
    if (need) compact_rows(need, p.kc, kCompactSlack);
  }

  __device__ __forceinline__ void end_segment(const Segment& seg) {
    const uint32_t need = __ballot_sync(0xffffffffu, cnt > p.kc);
    if (need) compact_rows(need, p.kc, 0);   // the pool slot holds exactly <= kc entries
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
      const int c = __shfl_sync(0xffffffffu, cnt, r);
      const int grow = __shfl_sync(0xffffffffu, row, r);
      if (grow >= p.Q) continue;  // warp-uniform
      const size_t slot = static_cast<size_t>(grow) * p.n_groups + seg.aux;
      const uint2* src = warp_buf + static_cast<size_t>(r) * kCap;
      uint2* dst = p.pool + slot * kMaxCand;
      for (int j = lane; j < c; j += 32) dst[j] = src[j];
      if (lane == 0) p.pool_cnt[slot] = c;
    }
    __syncwarp();
  }
};

// m-blocks x n-groups segments; consecutive segment ids share an n-group so the
// CTAs running concurrently stream the SAME database tiles (one HBM read, the
// rest L2 hits) while the whole query matrix stays L2-resident.
struct TopkSched {
  ISB_PLAIN_SEGMENT_PASSES
  int m_blocks, n_tiles, n_groups, k_blocks;
  // Wave barrier.  Segments are handed out round-robin, so the workers (CTAs or CTA pairs) go
  // through them in waves; the row blocks of one n-group stream the same database tiles and
  // share them through L2 only while they run in step.  Their start times drift apart from
  // wave to wave (ncu: 84 GB of DRAM reads for a 4.1 GB database in the CTA-pair kernel), so
  // every worker checks in at the start of each segment and waits -- for a bounded time, the
  // barrier is advisory -- until the whole wave has arrived.  wave_sync[w] counts arrivals.
  int* wave_sync;   // [kMaxWaves] zeroed before the launch; null: no barrier
  int window;       // unused (0) unless the barrier is on
  __device__ __forceinline__ int num_segments() const { return m_blocks * n_groups; }
  __device__ __forceinline__ Segment segment(int s) const {
    Segment seg;
    const int g = s / m_blocks;
    seg.m_block = s - g * m_blocks;
    seg.nt_begin = static_cast<int>(static_cast<long long>(g) * n_tiles / n_groups);
    seg.nt_end = static_cast<int>(static_cast<long long>(g + 1) * n_tiles / n_groups);
    seg.kb_begin = 0;
    seg.kb_end = k_blocks;
    seg.aux = g;
    seg.n_tile = 0;
    return seg;
  }
  // called by ALL 32 lanes of the producer warp before n-tile nt of the segment is loaded;
  // wave = how many segments this worker has finished, workers = CTAs (pairs) in the grid
  __device__ __forceinline__ void gate(const Segment& seg, int nt, int wave, int workers) const {
    if (wave_sync == nullptr || nt != seg.nt_begin || wave >= kMaxWaves) return;   // warp-uniform
    if ((threadIdx.x & 31) == 0) {
      const int total = num_segments();
      const int expected = min(workers, total - wave * workers);
      volatile int* cnt = wave_sync + wave;
      atomicAdd(wave_sync + wave, 1);
      for (int polls = 0; polls < 2048 && *cnt < expected; ++polls) __nanosleep(100);
    }
    __syncwarp();
  }
  __device__ __forceinline__ void leave(const Segment&) const {}
};

// k-th largest of n keys in shared memory by a 4 x 8-bit radix select; hist: 256 ints of
// the warp's own scratch.  Precondition n >= kth >= 1; all 32 lanes call it together.
__device__ __forceinline__ uint32_t warp_kth_largest(const uint32_t* keys, int n, int kth, int* hist) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0u;
  int remaining = kth;
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t pmask = (pass == 0) ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int b = lane; b < 256; b += 32) hist[b] = 0;
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
      const uint32_t key = keys[e];
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncwarp();
    // lane L owns bins 255 - 8L ... 248 - 8L; running count from the top bin downwards
    int c[8], local = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = hist[255 - 8 * lane - j];
      local += c[j];
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int excl = incl - local;
    const bool mine = excl < remaining && remaining <= incl;
    const uint32_t bal = __ballot_sync(0xffffffffu, mine);
    if (bal == 0u) return 0u;   // n < kth (precondition violated): everything qualifies
    const int src = __ffs(bal) - 1;
    int d = 0, above = 0;
    if (mine) {
      int run = excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (run + c[j] >= remaining) {
          d = 255 - 8 * lane - j;
          above = run;
          break;
        }
        run += c[j];
      }
    }
    d = __shfl_sync(0xffffffffu, d, src);
    above = __shfl_sync(0xffffffffu, above, src);
    prefix |= static_cast<uint32_t>(d) << shift;
    remaining -= above;
    __syncwarp();
  }
  return prefix;
}

}  // namespace isb
