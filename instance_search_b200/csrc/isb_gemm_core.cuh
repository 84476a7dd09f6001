// Warp-specialised tcgen05 GEMM mainloop for sm_100a:  acc[128 x 256] (fp32, TMEM)
//   = A[m-block, K] (bf16, K-major) . B[n-tile, K]^T (bf16, K-major)
//
//   warp 0      TMA producer   (one lane): global -> 128B-swizzled smem ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma 128x256x16, 4 per k-block,
//                              tcgen05.commit frees the smem slot / publishes the tile
//   warps 2..5  epilogue       (128 threads = 128 TMEM lanes = 128 rows of the tile)
//
// TMEM holds two 256-column accumulators, so the epilogue of tile t overlaps the
// MMAs of tile t+1.  Work is a list of SEGMENTS handed out round-robin to a
// persistent grid (segment s -> CTA s % gridDim.x): a segment is one m-block
// times a run of n-tiles times a k-block range.  What happens to a finished
// accumulator tile is the Epilogue policy's business (streaming top-k filter,
// fp32 store, ...).
#pragma once

#include "isb_ptx.cuh"

namespace isb {

constexpr int kBM = 128;      // rows of A per tile (= TMEM lanes)
constexpr int kBN = 256;      // rows of B per tile (= TMEM columns per accumulator)
constexpr int kBK = 64;       // bf16 per k-block = one 128-byte swizzle row
constexpr int kStages = 4;    // smem ring depth
constexpr int kUmmaK = 16;    // K of one tcgen05.mma kind::f16
constexpr int kGemmThreads = 192;
constexpr int kEpiWarp0 = 2;  // first epilogue warp
constexpr uint32_t kTmemCols = 512;
constexpr int kSingleTerm = 0x7FFFFFFF;  // kb_per_term of a plain (one bf16 term) GEMM

constexpr uint32_t kABytes = kBM * kBK * 2;                       // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 2;                       // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;               // 48 KB
constexpr uint32_t kRingBytes = kStages * kStageBytes;            // 192 KB
constexpr uint32_t kGemmSmemBytes = kRingBytes + 1024 /*align*/ + 256 /*barriers*/;

// ------------------------------------------------------------------ role timeline (debug builds)
// make TIMELINE=1 compiles the CTA-pair kernel with clock64 accounting of where each role
// waits: the totals land in isb_timeline[blockIdx.x][slot] (cycles) and are read back with
// isb_debug_timeline().  Off by default: the shipped kernel carries none of it.
//   0 kernel total   1 producer: wave gate   2 producer: free stage (empty)
//   3 MMA: accumulator drained (tmem_empty + peer)   4 MMA: stage loaded (full)
//   5 epilogue warp 2: accumulator ready (tmem_full)   6 epilogue warp 2: tile()   7 tiles
constexpr int kTimelineSlots = 8;
#ifdef ISB_TIMELINE
__device__ long long isb_timeline[256][kTimelineSlots];
#define ISB_TL_DECL(name) long long name = 0
#define ISB_TL_BEGIN(t) const long long t = clock64()
#define ISB_TL_ADD(acc, t) acc += clock64() - t
#define ISB_TL_STORE(slot, v) isb_timeline[blockIdx.x & 255][slot] = (v)
#else
#define ISB_TL_DECL(name)
#define ISB_TL_BEGIN(t)
#define ISB_TL_ADD(acc, t)
#define ISB_TL_STORE(slot, v)
#endif

struct Segment {
  int m_block;   // rows [m_block*128, +128) of A
  int nt_begin;  // n-tiles [nt_begin, nt_end) of B (256 rows each)
  int nt_end;
  int kb_begin;  // k-blocks [kb_begin, kb_end) (64 columns each)
  int kb_end;
  int aux;       // scheduler-defined (n-group id / k-split id)
  int n_tile;    // chunked schedulers: the one n-tile of B the whole segment works on
};

// Scheduler interface (besides num_segments / segment / gate / leave): every pass `nt` of a
// segment produces one accumulator tile
//   int  b_tile(seg, nt)              n-tile of B loaded for the pass (plain schedulers: nt)
//   void kb_range(seg, nt, kb0, kb1)  k-blocks accumulated in the pass (plain: the segment's)
// A chunked scheduler (PlainSched with chunk_kb) runs SEVERAL passes over the same output tile,
// each over a slice of the k-range: the tensor core's fp32 accumulator loses low bits on every
// one of the thousands of accumulation steps of a long chain (measured: 3.6e-4 relative on a
// 300k-long chain, 5e-6 on 2k-long ones, tools/gemm_precision.py), so long contractions are cut
// into chunks whose partial tiles the epilogue adds up in fp32 outside the tensor core.
// Plain schedulers get the two trivial members from ISB_PLAIN_SEGMENT_PASSES.
#define ISB_PLAIN_SEGMENT_PASSES                                                                    \
  __device__ __forceinline__ int b_tile(const Segment&, int nt) const { return nt; }                \
  __device__ __forceinline__ void kb_range(const Segment& seg, int, int& kb0, int& kb1) const {     \
    kb0 = seg.kb_begin;                                                                             \
    kb1 = seg.kb_end;                                                                               \
  }

struct GemmSmem {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// Epilogue policy interface (all called by the 128 epilogue threads):
//   Epi(params, row_in_tile)                    per-thread state
//   void begin_segment(const Segment&)
//   void tile(const Segment&, int nt, uint32_t tmem_acc, uint64_t* tmem_empty_bar)
//        must read the accumulator (lane = row_in_tile, columns [0,256)) with
//        tcgen05.ld, then tcgen05.fence::before + arrive on tmem_empty_bar
//        (every epilogue thread arrives once per tile).
//   void end_segment(const Segment&)
//
// Split operands: an fp32-grade product of fp32 matrices is the sum of three bf16
// products  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  (isb_f32_to_bf16 parts 0 / 1).  The
// k-block index of a segment then runs over 3 * kb_per_term blocks, kb = 3 * kk + term, and
// the producer switches tensor maps per term -- term 1 reads A from tmap_a_lo, term 2
// reads B from tmap_b_lo -- so neither operand is ever stored K-concatenated.
// Single-term callers pass kb_per_term = INT_MAX (and any valid maps for *_lo).
template <class Sched, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_a_lo,
               const __grid_constant__ CUtensorMap tmap_b_lo, const int kb_per_term,
               const Sched sched, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  GemmSmem* bars = reinterpret_cast<GemmSmem*>(ring + kRingBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    if (kb_per_term != kSingleTerm) {
      ptx::prefetch_tensormap(&tmap_a_lo);
      ptx::prefetch_tensormap(&tmap_b_lo);
    }
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars->tmem_full[b], 1);
      ptx::mbar_init(&bars->tmem_empty[b], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(&bars->tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int num_segments = sched.num_segments();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (lane 0 issues; the whole warp takes part in the scheduler's gate)
    int stage = 0;
    uint32_t phase = 0;
    int wave = 0;
    for (int s = blockIdx.x; s < num_segments; s += gridDim.x, ++wave) {
      const Segment seg = sched.segment(s);
      for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt) {
        sched.gate(seg, nt, wave, gridDim.x);
        if (lane == 0) {
          int kb0, kb1;
          sched.kb_range(seg, nt, kb0, kb1);
          const int bt = sched.b_tile(seg, nt);
          for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* sa = ring + stage * kStageBytes;
            uint8_t* sb = sa + kABytes;
            ptx::mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
            // split operands: the three terms of a k-range run back to back (kb = 3 * kk + term), so
            // the B_hi tile that terms 0 and 1 share is fetched from HBM once and hits L2 the second
            // time (term-major order streamed the whole of B_hi twice: 1.23 GB instead of 0.82 GB
            // of weights per batch in the region projection)
            const bool split = kb_per_term != kSingleTerm;
            const int term = split ? kb % 3 : 0;
            const int kk = split ? kb / 3 : kb;
            const CUtensorMap* ma = (term == 1) ? &tmap_a_lo : &tmap_a;
            const CUtensorMap* mb = (term == 2) ? &tmap_b_lo : &tmap_b;
            // A (queries / activations) is re-read for every n-tile: keep it in L2.
            ptx::tma_load_2d(sa, ma, &bars->full[stage], kk * kBK, seg.m_block * kBM,
                             ptx::kEvictLast);
            ptx::tma_load_2d(sb, mb, &bars->full[stage], kk * kBK, bt * kBN,
                             ptx::kEvictNormal);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        __syncwarp();
      }
      if (lane == 0) sched.leave(seg);
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_iter = 0;
      for (int s = blockIdx.x; s < num_segments; s += gridDim.x) {
        const Segment seg = sched.segment(s);
        for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt, ++acc_iter) {
          const uint32_t buf = acc_iter & 1;
          const uint32_t acc_phase = (acc_iter >> 1) & 1;
          ptx::mbar_wait(&bars->tmem_empty[buf], acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_acc = tmem_base + buf * kBN;
          int kb0, kb1;
          sched.kb_range(seg, nt, kb0, kb1);
          for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(&bars->full[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(ring + stage * kStageBytes);
            const uint64_t da = ptx::make_smem_desc_k_sw128(sa);
            const uint64_t db = ptx::make_smem_desc_k_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in
              // the (addr >> 4) start-address field
              ptx::umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc,
                             (kb > kb0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit(&bars->empty[stage]);  // smem slot free once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&bars->tmem_full[buf]);  // accumulator complete
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    // A warp may only touch TMEM lanes [32*(warp%4), +32).
    const int lane_group = warp & 3;
    const int row_in_tile = lane_group * 32 + lane;
    Epi epi(ep, row_in_tile);
    uint32_t acc_iter = 0;
    for (int s = blockIdx.x; s < num_segments; s += gridDim.x) {
      const Segment seg = sched.segment(s);
      epi.begin_segment(seg);
      for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt, ++acc_iter) {
        const uint32_t buf = acc_iter & 1;
        const uint32_t acc_phase = (acc_iter >> 1) & 1;
        ptx::mbar_wait(&bars->tmem_full[buf], acc_phase);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * kBN + (static_cast<uint32_t>(lane_group * 32) << 16);
        epi.tile(seg, nt, tmem_acc, &bars->tmem_empty[buf]);
      }
      epi.end_segment(seg);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ CTA-pair variant
// Same roles, one tile of 256 x 256 per CTA PAIR (cluster of 2 = the two SMs of a TPC):
//   * each CTA loads ITS 128 rows of A and ITS half (128 rows) of the B tile -- 32 KB per
//     stage instead of 48 KB, so the ring is 6 deep, and the tensor core reads every B row
//     from shared memory once per 256 accumulator rows instead of once per 128 (the single-CTA
//     kernel moves 192 B/clk through shared memory, TMA writes + operand reads, against a
//     128 B/clk port: 74 % tensor-pipe utilisation, profiles/r01_ncu_search_screen_1M_v2.txt)
//   * the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256; both CTAs'
//     TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor.cta_group::2),
//     the leader's commits are multicast to both CTAs' empty / tmem_full barriers
//   * every CTA's epilogue warps drain their own 128 TMEM lanes (rows 128 * rank + lane of
//     the pair tile); the peer's warp 1, which has no MMAs to issue, relays "my accumulator
//     is drained" to the leader (one remote arrive per tile), so the Epilogue policies are the
//     single-CTA ones unchanged.
// The scheduler enumerates segments in units of PAIR row blocks: seg.m_block counts 256-row
// blocks; CTA `rank` works on the 128-row block 2 * seg.m_block + rank.
constexpr int kPairStages = 6;
constexpr int kPairBRows = kBN / 2;                                       // B rows loaded per CTA
constexpr uint32_t kPairBBytes = kPairBRows * kBK * 2;                    // 16 KB
constexpr uint32_t kPairStageBytes = kABytes + kPairBBytes;               // 32 KB
constexpr uint32_t kPairRingBytes = kPairStages * kPairStageBytes;        // 192 KB
constexpr uint32_t kPairSmemBytes = kPairRingBytes + 1024 /*align*/ + 256 /*barriers*/;

struct PairSmem {
  uint64_t full[kPairStages];    // leader's: TMA bytes of BOTH CTAs for the stage have landed
  uint64_t empty[kPairStages];   // each CTA's: the MMAs reading the stage retired (multicast commit)
  uint64_t tmem_full[2];         // each CTA's: accumulator complete (multicast commit)
  uint64_t tmem_empty[2];        // each CTA's: its 128 epilogue threads drained the accumulator
  uint64_t peer_empty[2];        // leader's: the peer drained ITS accumulator (relayed)
  uint32_t tmem_base;
};

template <class Sched, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_a_lo,
                    const __grid_constant__ CUtensorMap tmap_b_lo, const int kb_per_term,
                    const Sched sched, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  PairSmem* bars = reinterpret_cast<PairSmem*>(ring + kPairRingBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    if (kb_per_term != kSingleTerm) {
      ptx::prefetch_tensormap(&tmap_a_lo);
      ptx::prefetch_tensormap(&tmap_b_lo);
    }
    for (int s = 0; s < kPairStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars->tmem_full[b], 1);
      ptx::mbar_init(&bars->tmem_empty[b], 128);
      ptx::mbar_init(&bars->peer_empty[b], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair<kTmemCols>(&bars->tmem_base);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // both CTAs' barriers are initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int num_segments = sched.num_segments();
  ISB_TL_BEGIN(tl_kernel0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // (lane 0 issues; the leader's whole warp takes part in the scheduler's gate, the peer
    // follows through the shared empty barriers)
    int stage = 0;
    uint32_t phase = 0;
    int wave = 0;
    ISB_TL_DECL(tl_gate);
    ISB_TL_DECL(tl_empty);
    for (int s = pair_id; s < num_segments; s += n_pairs, ++wave) {
      const Segment seg = sched.segment(s);
      const int my_m_block = 2 * seg.m_block + static_cast<int>(rank);
      for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt) {
        ISB_TL_BEGIN(tl_g0);
        if (leader) sched.gate(seg, nt, wave, n_pairs);
        ISB_TL_ADD(tl_gate, tl_g0);
        if (lane == 0) {
          int kb0, kb1;
          sched.kb_range(seg, nt, kb0, kb1);
          const int bt = sched.b_tile(seg, nt);
          for (int kb = kb0; kb < kb1; ++kb) {
            ISB_TL_BEGIN(tl_e0);
            ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
            ISB_TL_ADD(tl_empty, tl_e0);
            uint8_t* sa = ring + stage * kPairStageBytes;
            uint8_t* sb = sa + kABytes;
            if (leader) ptx::mbar_arrive_expect_tx(&bars->full[stage], 2 * kPairStageBytes);
            const uint32_t full_leader = ptx::mapa_u32(ptx::smem_u32(&bars->full[stage]), 0);
            // split operands: the three terms of a k-range run back to back (kb = 3 * kk + term), so
            // the B_hi tile that terms 0 and 1 share is fetched from HBM once and hits L2 the second
            // time (term-major order streamed the whole of B_hi twice: 1.23 GB instead of 0.82 GB
            // of weights per batch in the region projection)
            const bool split = kb_per_term != kSingleTerm;
            const int term = split ? kb % 3 : 0;
            const int kk = split ? kb / 3 : kb;
            const CUtensorMap* ma = (term == 1) ? &tmap_a_lo : &tmap_a;
            const CUtensorMap* mb = (term == 2) ? &tmap_b_lo : &tmap_b;
            ptx::tma_load_2d_pair(sa, ma, full_leader, kk * kBK, my_m_block * kBM, ptx::kEvictLast);
            ptx::tma_load_2d_pair(sb, mb, full_leader, kk * kBK, bt * kBN + static_cast<int>(rank) * kPairBRows,
                                  ptx::kEvictNormal);
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
          }
        }
        __syncwarp();
      }
      if (leader && lane == 0) sched.leave(seg);
    }
    if (lane == 0) {
      ISB_TL_STORE(1, tl_gate);
      ISB_TL_STORE(2, tl_empty);
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ------------------------------------------------------------ MMA issuer (leader)
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(2 * kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_iter = 0;
      ISB_TL_DECL(tl_drain);
      ISB_TL_DECL(tl_full);
      for (int s = pair_id; s < num_segments; s += n_pairs) {
        const Segment seg = sched.segment(s);
        for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt, ++acc_iter) {
          const uint32_t buf = acc_iter & 1;
          const uint32_t acc_phase = (acc_iter >> 1) & 1;
          ISB_TL_BEGIN(tl_d0);
          ptx::mbar_wait(&bars->tmem_empty[buf], acc_phase ^ 1);
          ptx::mbar_wait(&bars->peer_empty[buf], acc_phase ^ 1);
          ISB_TL_ADD(tl_drain, tl_d0);
          ptx::tc_fence_after();
          const uint32_t tmem_acc = tmem_base + buf * kBN;
          int kb0, kb1;
          sched.kb_range(seg, nt, kb0, kb1);
          for (int kb = kb0; kb < kb1; ++kb) {
            ISB_TL_BEGIN(tl_f0);
            ptx::mbar_wait(&bars->full[stage], phase);
            ISB_TL_ADD(tl_full, tl_f0);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(ring + stage * kPairStageBytes);
            const uint64_t da = ptx::make_smem_desc_k_sw128(sa);
            const uint64_t db = ptx::make_smem_desc_k_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              ptx::umma_bf16_pair(tmem_acc, da + 2 * k, db + 2 * k, idesc,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit_pair(&bars->empty[stage], 3);   // both CTAs' slots are free
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit_pair(&bars->tmem_full[buf], 3);   // both CTAs' accumulators complete
        }
      }
      ISB_TL_STORE(3, tl_drain);
      ISB_TL_STORE(4, tl_full);
      ISB_TL_STORE(7, static_cast<long long>(acc_iter));
    } else if (lane == 0) {
      // ------------------------------------------------------------ drain relay (peer)
      uint32_t acc_iter = 0;
      for (int s = pair_id; s < num_segments; s += n_pairs) {
        const Segment seg = sched.segment(s);
        for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt, ++acc_iter) {
          const uint32_t buf = acc_iter & 1;
          const uint32_t acc_phase = (acc_iter >> 1) & 1;
          ptx::mbar_wait(&bars->tmem_empty[buf], acc_phase);
          ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&bars->peer_empty[buf]), 0));
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (both CTAs)
    const int lane_group = warp & 3;
    const int row_in_tile = lane_group * 32 + lane;
    Epi epi(ep, row_in_tile);
    uint32_t acc_iter = 0;
    ISB_TL_DECL(tl_ready);
    ISB_TL_DECL(tl_tile);
    for (int s = pair_id; s < num_segments; s += n_pairs) {
      Segment seg = sched.segment(s);
      seg.m_block = 2 * seg.m_block + static_cast<int>(rank);
      epi.begin_segment(seg);
      for (int nt = seg.nt_begin; nt < seg.nt_end; ++nt, ++acc_iter) {
        const uint32_t buf = acc_iter & 1;
        const uint32_t acc_phase = (acc_iter >> 1) & 1;
        ISB_TL_BEGIN(tl_r0);
        ptx::mbar_wait(&bars->tmem_full[buf], acc_phase);
        ISB_TL_ADD(tl_ready, tl_r0);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * kBN + (static_cast<uint32_t>(lane_group * 32) << 16);
        ISB_TL_BEGIN(tl_t0);
        epi.tile(seg, nt, tmem_acc, &bars->tmem_empty[buf]);
        ISB_TL_ADD(tl_tile, tl_t0);
      }
      epi.end_segment(seg);
    }
    if (warp == kEpiWarp0 && lane == 0) {
      ISB_TL_STORE(5, tl_ready);
      ISB_TL_STORE(6, tl_tile);
    }
  }

  if (threadIdx.x == 0) {
    ISB_TL_STORE(0, clock64() - tl_kernel0);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

}  // namespace isb
