/* isb.h -- C ABI of libisb.so, the B200 (sm_100a) implementation of the
 * retrieval hot path of maxgreat/Instance-Search.
 *
 * The reference is pure Python over torch; it has no FFI of its own.  Its
 * boundary for this path is the Python operator surface listed in SURVEY.md
 * section 8b.  Each entry point below names the reference lines it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the caller owns all
 *     buffers, the library never allocates outputs; scratch memory comes in
 *     through a workspace whose size the matching *_workspace_bytes() returns;
 *   - tensors are dense row-major; fp32 = float, bf16 = uint16_t bit pattern,
 *     indices int64_t (torch LongTensor) unless stated, masks/labels int32_t;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous on it and thread-safe per stream;
 *   - return value: ISB_OK or an ISB_ERR_* code; isb_last_error() returns a
 *     thread-local message for the last failing call on this thread;
 *   - no entry point has a CPU fallback: without an sm_100 device they fail.
 */
#ifndef ISB_H_
#define ISB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISB_OK 0
#define ISB_ERR_INVALID_ARGUMENT 1
#define ISB_ERR_CUDA 2
#define ISB_ERR_WORKSPACE 3
#define ISB_ERR_UNSUPPORTED_DEVICE 4

#define ISB_ABI_VERSION 6

/* Largest k (after the screening margin is added) one search call supports. */
#define ISB_MAX_CANDIDATES 128

int isb_abi_version(void);
const char* isb_last_error(void);
/* 0 when the current device is compute capability 10.x, else an error code. */
int isb_check_device(void);

/* Tuning options.  The library reads nothing from the environment: the A/B switches that
 * let a claim of DESIGN.md be re-measured are set explicitly, process-wide, by the caller.
 * Defaults are the measured best.  An option is read when a call makes its launch plan; one
 * that changes a workspace layout (SCREEN_PAIR, SCREEN_SEED, POOL_G) must not change between
 * a *_workspace_bytes() query and the call it sizes.
 * isb_set_option(option, value >= 0) sets, value == -1 restores the default;
 * isb_get_option returns the set value or -1 (default in force / unknown option). */
enum {
  ISB_OPT_SCREEN_PAIR = 0,       /* 1 (default): CTA-pair 256 x 256 screen when it fits; 0: single-CTA kernel */
  ISB_OPT_SCREEN_SEED = 1,       /* 1: seed the screen thresholds from a sample of the database (default 0) */
  ISB_OPT_SCREEN_WAVESYNC = 2,   /* 1 (default): wave barrier of the screen scheduler; 0: off */
  ISB_OPT_MINING_KC = 3,         /* candidates re-checked exactly per couple, 1..128 (default 16) */
  ISB_OPT_POOL_STAGES = 4,       /* ring depth of the pooling kernel (default 3) */
  ISB_OPT_POOL_G = 5,            /* channel blocks per pooling CTA (default: fills whole waves) */
  ISB_OPT_POOL_GENERIC_GEOM = 6, /* 1: pooling kernel without compile-time geometry (default 0) */
  ISB_OPT_REGION_POOL_TC = 7,    /* 1: tensor-core pooling kernel (default 0) */
  ISB_OPT_TC_DEBUG = 8,          /* ablation flags of the tensor-core pooling kernel (default 0) */
  ISB_OPT_GATHER_CW = 9,         /* channels per warp unit of the gather kernel */
  ISB_OPT_GATHER_G = 10,         /* units per warp of the gather kernel */
  ISB_OPT_GATHER_STAGES = 11,    /* ring depth of the gather kernel */
  ISB_OPT_GEMM_PAIR = 12,        /* 1 (default): CTA-pair kernel for isb_gemm_nt[_split] with >= 2 row blocks */
  ISB_OPT_MINING_PROGRESSIVE = 13, /* exact re-check of a couple's candidates: 1 (default) progressive, 0 all of them, 2 progressive with the one-CTA-per-couple kernel */
  ISB_OPT_GATHER_SMALL = 14,     /* 1 (default): plane-block gather kernel for maps of <= 256 pixels; 0: off */
  ISB_OPT_SCREEN_GROUPS = 15,    /* n-groups of the screen's work decomposition (default 0: chosen for whole waves) */
  ISB_OPT_RESC_SPLITS = 16,      /* split-K of the region head's candidate re-score GEMM (default 0: fills the machine) */
  ISB_OPT_REGION_TOP_SELECT = 17, /* 1 (default): one-kernel candidate re-score from the screen's best classes; 0: re-score GEMM */
  ISB_OPT_COUNT_ = 18
};
int isb_set_option(int option, int value);
int isb_get_option(int option);

/* ---------------------------------------------------------------- a1 / a5
 * y[m, :] = x[m, :] / sqrt(sum_j x[m, j]^2 + eps)          (eps INSIDE the sqrt)
 * replaces NormalizeL2Fun.forward, model/custom_modules.py:52-57 (module :70-76;
 * used at model/siamese.py:178,182,222).  x and y may alias. */
int isb_l2norm_rows(const float* x, int64_t M, int64_t F, float eps, float* y, void* stream);

/* ---------------------------------------------------------------- a2
 * y[m, j] = x[m, j] + param[j]
 * replaces ShiftFun.forward, model/custom_modules.py:16-18 (module :28-39). */
int isb_shift_rows(const float* x, const float* param, int64_t M, int64_t F, float* y,
                   void* stream);

/* fp32 [rows, cols] (leading dimension ldx) -> bf16 [rows, ldy], columns
 * [cols, ldy) zero-filled.  ldy must be a multiple of 8 (16-byte rows for TMA).
 * Data preparation for the tensor-core operands (database, queries, weights);
 * `part` selects which bf16 term of the fp32 value is produced:
 *   0: hi = bf16(x)     1: lo = bf16(x - hi)     2: lo2 = bf16(x - hi - lo)  */
int isb_f32_to_bf16(const float* x, int64_t rows, int64_t cols, int64_t ldx, uint16_t* y,
                    int64_t ldy, int part, void* stream);

/* ---------------------------------------------------------------- a8 + a9/a10
 * Cosine top-k:  for every query row, the k database rows with the largest
 * q . db, best first -- the first k entries of the descending sort the
 * reference takes of each row of  sim = torch.mm(Q, DB.t())
 * (test/siamese_regions_test.py:76, utils/train_siamese.py:70, consumed by
 * utils/metrics.py:11,13,33) without materialising the Q x N matrix.
 *
 *   stage 1  bf16 tcgen05 GEMM, fp32 accumulate, streaming per-row top-(k+margin)
 *            filter fused into the TMEM epilogue        (db_bf16, ld_bf16)
 *   stage 2  exact re-rank of the k+margin candidates: dot products of the
 *            fp32 rows accumulated in fp64, sorted, ties -> lower index
 *   stage 2b completeness certificate per query: every database row outside the
 *            candidate list has a screen score <= t_min (the worst candidate's),
 *            so the list contains the true top k -- to 8 sigma of the screen noise,
 *            a statistical statement, not a proof -- when
 *               exact_kth - t_min  >  8 * sigma + 4e-7 * |score|
 *            with sigma = rms(screen - exact) MEASURED on the row's own
 *            candidates and floored by the noise bf16 operand rounding is expected
 *            to have on dense rows (2.34e-3 |q||db| / sqrt(D); half of it from 32
 *            candidates on, 1.5x below: a sigma from a few samples can be ~0).
 *            Rows that fail are appended to uncertified_rows and
 *            counted in *n_uncertified (both device memory, may both be NULL to
 *            skip the test); the caller resolves them with isb_topk_resolve and,
 *            if still uncertified, isb_topk_exhaustive.
 *
 * q        [Q, D] fp32           db_f32  [N, D] fp32
 * db_bf16  [N, ld_bf16] bf16 made by isb_f32_to_bf16(part 0), ld_bf16 >= D, %8 == 0
 * k + margin <= ISB_MAX_CANDIDATES;  k <= N;  N < 2^31
 * idx_offset is added to every returned index (row offset of a database shard)
 * out_scores [Q, k] fp32 (the fp64 dot rounded to fp32), out_idx [Q, k] int64
 * uncertified_rows [Q] int32 (order unspecified), n_uncertified [1] int32 */
size_t isb_topk_search_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k, int margin);
int isb_topk_search(const float* q, int64_t Q, const float* db_f32, const uint16_t* db_bf16,
                    int64_t N, int64_t D, int64_t ld_bf16, int k, int margin,
                    int64_t idx_offset, float* out_scores, int64_t* out_idx,
                    int32_t* uncertified_rows, int32_t* n_uncertified, void* workspace,
                    size_t workspace_bytes, void* stream);

/* The two stages of isb_topk_search as separate calls sharing one workspace
 * (isb_topk_search == screen then rerank).  The screen leaves, per query, the
 * k+margin best bf16-scored candidates in the workspace; the rerank consumes
 * them.  Exposed so a caller can time / overlap the stages. */
int isb_topk_screen(const float* q, int64_t Q, const uint16_t* db_bf16, int64_t N, int64_t D,
                    int64_t ld_bf16, int k, int margin, void* workspace, size_t workspace_bytes,
                    void* stream);
int isb_topk_rerank(const float* q, int64_t Q, const float* db_f32, int64_t N, int64_t D, int k,
                    int margin, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                    int32_t* uncertified_rows, int32_t* n_uncertified, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Second line for the rows a bf16 screen could not certify (near-duplicate-dense
 * neighbourhoods): the same screen with fp32-grade split operands
 *   score = q_hi.db_hi + q_lo.db_hi + q_hi.db_lo      (three bf16 tcgen05 products,
 * error ~1e-6 on unit rows), exact re-rank and the same certificate.  Results of
 * the listed rows are (re)written to out_scores / out_idx [Q, k] at their own row;
 * rows still uncertified are listed in uncertified_rows / n_uncertified.
 *   rows [n_rows] int32 device: query rows to resolve (from isb_topk_search)
 *   db_lo_bf16 = isb_f32_to_bf16(db_f32, part 1), same leading dimension as db_bf16 */
size_t isb_topk_resolve_workspace_bytes(int64_t n_rows, int64_t N, int64_t D);
int isb_topk_resolve(const float* q, const float* db_f32, const uint16_t* db_bf16,
                     const uint16_t* db_lo_bf16, int64_t N, int64_t D, int64_t ld_bf16, int k,
                     int margin, int64_t idx_offset, const int32_t* rows, int64_t n_rows,
                     float* out_scores, int64_t* out_idx, int32_t* uncertified_rows,
                     int32_t* n_uncertified, void* workspace, size_t workspace_bytes, void* stream);

/* Last line: exhaustive exact (fp64-accumulated) scoring of the whole database
 * for the listed rows -- what torch.mm + sort does, ties -> lower index.  Needed
 * only when more than `margin` database rows tie with the k-th score within fp32
 * noise (e.g. duplicated database entries). */
size_t isb_topk_exhaustive_workspace_bytes(int64_t n_rows, int64_t N, int k);
int isb_topk_exhaustive(const float* q, const float* db_f32, int64_t N, int64_t D, int k,
                        int64_t idx_offset, const int32_t* rows, int64_t n_rows, float* out_scores,
                        int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- (e) multi-GPU
 * Merge R per-shard results (after the all-gather): cand_scores [R, Q, k],
 * cand_idx [R, Q, k] (global indices) -> the k best per query, best first,
 * ties -> lower index.  The reference has no multi-device code; this is the
 * exchange step of the row-sharded search. */
int isb_topk_merge(const float* cand_scores, const int64_t* cand_idx, int R, int64_t Q, int k,
                   float* out_scores, int64_t* out_idx, void* stream);

/* Row-sharded search with candidate exchange (one process per GPU, R shards).  Running
 * isb_topk_rerank on every shard costs R times the gathers of the single-GPU search
 * although only k + margin of the R * (k + margin) candidates can reach the result.
 * Instead, after isb_topk_screen on its shard, every rank
 *   1. isb_topk_candidates        lists its kc_out = k + margin best screen entries per query:
 *                                 cand_screen [Q, kc_out] fp32 (-inf = none), cand_col [Q, kc_out]
 *                                 int32 local row (-1 = none); `k`, `margin` as passed to the screen
 *      -- all-gather of cand_screen (4 bytes per candidate) --
 *   2. isb_topk_global_threshold  thr [Q] = the kth-best screen score over all shards
 *                                 (all_screen [R, Q, kc]; kth <= 0: kth = kc); -inf when fewer exist
 *   3. isb_topk_rerank_owned      exact fp64-accumulated scores of ITS candidates >= thr, sorted
 *                                 best first into ONE packed row of 2k + 2 32-bit words per query:
 *                                 k scores (fp32, -inf padded) | k local rows (int32, -1 padded) |
 *                                 sum (screen - exact)^2 | candidates scored   (both fp32)
 *      -- ONE all-gather of the packed rows (8k + 8 bytes per query and shard) --
 *   4. isb_topk_merge_certified   the k best of packed_all [R, Q, 2k + 2] (global index = local row +
 *                                 row_offsets[shard], int64 [R] device), best first, ties -> lower
 *                                 index, plus the completeness certificate of isb_topk_search
 *                                 evaluated globally: every row not re-ranked anywhere has a screen
 *                                 score <= thr, the noise is the rms over all shards' stat words;
 *                                 rows that fail are listed (resolve them with isb_topk_search on
 *                                 every shard + isb_topk_merge).
 * In total k + margin database rows are gathered per query, independent of R.
 * Reduced lists: the global top (k + margin) spreads over the shards, so a shard may list fewer
 * candidates than that (screen with k' + margin' = kc < k + margin; kth = k + margin in step 2; kc
 * may be smaller than k in step 3): its screen keeps fewer rows per query and warms its thresholds
 * up faster.  A shard whose FULL list lies entirely above thr may have been cut above it; step 3
 * then reports infinite noise for the row and step 4 lists it as uncertified.
 * Replaces the same reference lines as isb_topk_search; the reference is single-device. */
int isb_topk_candidates(int64_t Q, int64_t N, int64_t D, int k, int margin, int kc_out,
                        float* cand_screen, int32_t* cand_col, void* workspace,
                        size_t workspace_bytes, void* stream);
int isb_topk_global_threshold(const float* all_screen, int R, int64_t Q, int kc, int kth, float* thr,
                              void* stream);
int isb_topk_rerank_owned(const float* q, int64_t Q, const float* db_f32, int64_t N, int64_t D,
                          int k, int kc, const float* cand_screen, const int32_t* cand_col,
                          const float* thr, uint32_t* packed, void* stream);
int isb_topk_merge_certified(const uint32_t* packed_all, const int64_t* row_offsets, const float* thr,
                             int R, int64_t Q, int k, float* out_scores, int64_t* out_idx,
                             int32_t* uncertified_rows, int32_t* n_uncertified, void* stream);

/* ---------------------------------------------------------------- dense contraction
 * C[M, N] (fp32, leading dimension ldc) = A[M, K] . B[N, K]^T  (+ bias[N])
 * A, B bf16 row-major with leading dimensions lda, ldb (multiples of 8).
 * tcgen05 GEMM with fp32 accumulation; `splits` > 1 splits K across CTAs and
 * reduces the partial tiles in a fixed order (deterministic).  bias may be NULL.
 * Used for: all-pairs similarities  torch.mm(E, E.t())  utils/train_siamese.py:53,
 * test/instance_avg.py:12;  the whitening projection  nn.Linear(100352, D)
 * model/siamese.py:180;  the 1x1-conv window classifier  model/siamese.py:188.
 * isb_gemm_nt_split: fp32-grade product of two fp32 matrices given as split
 * operands (isb_f32_to_bf16 parts 0 and 1, same leading dimension per side):
 *   C = A_hi.B_hi^T + A_lo.B_hi^T + A_hi.B_lo^T   accumulated in ONE TMEM tile
 * (the TMA producer switches tensor maps per term; nothing is K-concatenated). */
size_t isb_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K, int splits);
int isb_gemm_nt(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, int64_t M,
                int64_t N, int64_t K, const float* bias, float* C, int64_t ldc, int splits,
                void* workspace, size_t workspace_bytes, void* stream);
int isb_gemm_nt_split(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda, const uint16_t* B_hi,
                      const uint16_t* B_lo, int64_t ldb, int64_t M, int64_t N, int64_t K,
                      const float* bias, float* C, int64_t ldc, int splits, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- a3
 * Window scoring + selection of RegionDescriptorNet.forward_single,
 * model/siamese.py:186-217, for a batch of trunk feature maps (each image is
 * treated exactly as the reference's batch-1 call, :184):
 *   a = AvgPool2d((fh, fw), stride 1)(x)            (:164-166, :187)
 *   c = Conv2d 1x1 (a) = cls_w . a + cls_b           (:188; model/nn_utils.py:26-39)
 *   m = max over classes (:191); the k' = min(H'W', k) windows with the largest m,
 *   best first (:193-194); cls_out[:, :, i] = c[:, :, window_i] (:216), zero beyond k'
 * A bf16 tcgen05 GEMM with the class-max fused in its epilogue screens all
 * windows; the k + margin best of every image are re-scored with fp32-grade
 * split operands (window means as bf16 hi + lo against cls_w hi + lo, three
 * tcgen05 products, ~1e-6) and the final order / logits come from that pass.
 * Completeness certificate as for the search: every non-candidate window has a
 * screen value <= t_min; an image whose k-th exact class-max does not clear t_min
 * by 8 sigma (sigma = rms(screen - exact) over its candidates) is reported in
 * n_uncertified (device int32 [1 + B], may be NULL): [0] = how many images, [1 + j] =
 * the j-th of them (order unspecified); the caller re-runs those images with the
 * exact second line (exact_mode, margin = 32 - k scores up to 32 windows exactly).
 * The logits of this pass are fp32-GRADE, not fp32: cls_out here is a preview
 * (abs. error ~1e-5 x logit scale); isb_region_logits recomputes the selected
 * windows' logits in true fp32, fixes their order and closes the certificate.
 * approx_max [B, k] (class-max of the selected windows as seen here) and
 * runner_up [B] (class-max of the best window NOT selected, -inf if none) feed it;
 * both may be NULL.
 * Default (ISB_OPT_REGION_TOP_SELECT = 1, up to 512 classes and 1024 windows per image): the
 * screen keeps the four best classes of every window and only the best THREE of each candidate
 * are re-scored (a class outside them has a screen logit <= the fourth best, m4; a window whose
 * re-scored class-max does not clear m4 by 8 sigma puts its image on the n_uncertified list).
 * cls_out then holds the re-scored logits of those classes and -inf for every other class --
 * a class-max preview, which is all isb_region_logits needs from it.  With the option at 0 the
 * candidates are re-scored against all classes (split-operand tensor-core GEMM).
 * exact_mode < 0: measurement probe -- only the pooling pass runs (window means as bf16
 * hi / lo into the workspace, energy partials); no output is written.
 * exact_mode > 0: second line for uncertified batches -- the candidates (use
 * margin = 32 - k) are re-scored from the fp32 inputs with fp64 accumulation; slow,
 * exact, cls_out final.
 *   x [B, C, H, W] fp32 (NCHW)    cls_w [ncls, C] fp32    cls_b [ncls] fp32
 *   cls_w_hi / cls_w_lo [ncls, ld_w] = isb_f32_to_bf16(cls_w, part 0 / 1)
 *   idx [B, k] int64: flat window index r * W' + c, -1 beyond k'    nsel [B] int32 = k'
 *   cls_out [B, ncls, k] fp32      win_norm [B, k] fp32 = sqrt(||crop_i||^2 + 1e-10)
 * k <= 32; windows re-scored per image = min(H'W', k + margin, 32). */
size_t isb_region_select_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t ncls,
                                         int fh, int fw, int k, int margin);
int isb_region_select(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, const float* cls_w,
                      const uint16_t* cls_w_hi, const uint16_t* cls_w_lo, int64_t ld_w,
                      const float* cls_b, int64_t ncls, int fh, int fw, int k, int margin,
                      int exact_mode, int64_t* idx, int32_t* nsel, float* cls_out, float* win_norm,
                      float* approx_max, float* runner_up, int32_t* n_uncertified, void* workspace,
                      size_t workspace_bytes, void* stream);

/* a3 / a4, closing stage.  isb_region_select is called with ke = k + 2 (the k windows
 * plus two runner-ups), isb_region_gather sums the first k into the operand and
 * leaves the exact fp32 mean of all ke windows in win_mean [B, ke, C]; this call
 * computes their logits in true fp32,  cls_w . mean + cls_b  (model/siamese.py:188),
 * orders the ke windows exactly (class-max desc, window asc, :191-194) and writes
 * the final k: idx_out [B, k], norm_out [B, k], nsel_out [B], cls_out [B, ncls, k]
 * (:216).  Images where a runner-up displaced one of the first k are appended to
 * changed_list / n_changed (device; may be NULL): their operand rows must be
 * re-gathered (isb_region_gather with image_list).  Certificate: the k-th exact
 * class-max must clear runner_up[b] -- the best fp32-grade value among the windows
 * NOT scored here -- by 8 sigma, sigma = rms(approx_max - exact) over the ke windows;
 * failures are reported in n_uncertified [1 + B] (count, then the images; NULL: no
 * certificate).
 * cls_out == NULL (eval: the reference's forward returns only the descriptor,
 * model/siamese.py:231): only the class-max is needed, so only the classes whose
 * fp32-grade logit (approx_cls [B, ncls, ke] = isb_region_select's cls_out) lies
 * within twice the worst-case error of the split, 2 * 3 * 2^-18 * sum|mean| * max|cls_w|
 * (wabs_max = max|cls_w|), of the window's best are evaluated in fp32. */
int isb_region_logits(const float* win_mean, const float* cls_w, const float* cls_b, int64_t B,
                      int64_t C, int64_t ncls, int ke, int k, const int32_t* nsel_in,
                      const float* approx_max, const float* runner_up, const int64_t* idx_in,
                      const float* norm_in, int64_t* idx_out, float* norm_out, int32_t* nsel_out,
                      float* cls_out, const float* approx_cls, float wabs_max, int32_t* changed_list,
                      int32_t* n_changed, int32_t* n_uncertified, void* stream);

/* ---------------------------------------------------------------- a4 (operand)
 * u[b, :] = sum_{i < nsel[b]} crop_i / win_norm[b, i]  +  nsel[b] * shift
 * where crop_i = x[b, :, r_i : r_i + fh, c_i : c_i + fw] flattened (C, fh, fw)
 * -- NormalizeL2 + Shift of every selected region (model/siamese.py:218-219 via
 * :178-179), summed BEFORE the projection (nn.Linear is linear; the reference
 * sums after it, :220).  Written as the tensor-core operand of the projection:
 *   U_hi[b] = bf16(u)      U_lo[b] = bf16(u - U_hi)   (U_lo may be NULL)
 * U_hi alone pairs with isb_gemm_nt (plain bf16 projection); U_hi + U_lo pair
 * with isb_gemm_nt_split (fp32-grade).  Rows hold C*fh*fw values zero padded to a
 * multiple of 8; ldu >= that, % 8 == 0.  idx / win_norm are [B, k]; only the first
 * min(nsel[b], k_sum) windows are summed.  win_mean [B, k, C] fp32 (may be NULL): the
 * exact mean of ALL nsel[b] listed windows, by-product for isb_region_logits.
 * image_list / n_list (device, may be NULL): process only the images
 * image_list[0 .. *n_list) -- the fix-up pass after isb_region_logits. */
int isb_region_gather(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh, int fw,
                      int k, int k_sum, const int32_t* image_list, const int32_t* n_list,
                      const int64_t* idx, const int32_t* nsel, const float* win_norm,
                      const float* shift, uint16_t* U_hi, uint16_t* U_lo, int64_t ldu, float* win_mean,
                      void* stream);

/* ---------------------------------------------------------------- a4 (bias) + a5
 * desc[b, :] = l2norm(y[b, :] + nsel[b] * bias)     (model/siamese.py:220-222)
 * y = U . W^T from isb_gemm_nt / isb_gemm_nt_split.  nsel == NULL means 1 (DescriptorNet, :117-122);
 * bias may be NULL. */
int isb_descriptor_finalize(const float* y, int64_t B, int64_t D, const float* bias,
                            const int32_t* nsel, float eps, float* desc, void* stream);

/* ---------------------------------------------------------------- f3: backward of the fused head
 * Training path of RegionDescriptorNet (model/siamese.py:199-222 with the backward passes of
 * model/custom_modules.py:20-25 (Shift) and :59-67 (NormalizeL2)), batched.  With, per image b and
 * selected window i < nsel[b] at idx[b, i]:  crop_i = the flattened C x fh x fw crop,
 *   u_b = sum_i crop_i / sqrt(|crop_i|^2 + eps) + nsel[b] * shift,   y_b = W u_b + nsel[b] * bias,
 *   cls_out[b, :, i] = Wc mean_i + bc,
 * the dense parts (g_u = g_y W, dW = g_y^T u, dWc = g_cls^T mean, g_mean = g_cls Wc) are
 * isb_gemm_nt_split calls; these two entry points are the glue around them.
 *
 * isb_region_crop_stats: crop_norm2[b, i] = |crop_i|^2, crop_dot[b, i] = <crop_i, g_u[b, :]>
 *   (0 when g_u == NULL), win_mean[b, i, :] = per-channel window means (may be NULL); slots
 *   i >= nsel[b] give 0.  g_u [B, ldg] fp32 is the gradient w.r.t. u.
 * isb_region_scatter_grad: g_x[b, c, h, w] = sum over the windows i covering (h, w) of
 *     g_u[b, j] / n_i - x[b, c, h, w] * crop_dot[b, i] / n_i^3  +  g_mean[b, i, c] / (fh * fw),
 *   n_i = sqrt(crop_norm2[b, i] + eps), j = the crop element (c, h - r_i, w - c_i).
 *   g_u / g_mean may be NULL (no descriptor / no cls_out gradient).  Gather form: no atomics,
 *   overlapping windows add in a fixed order.  k <= 32. */
int isb_region_crop_stats(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh, int fw, int k,
                          const int64_t* idx, const int32_t* nsel, const float* g_u, int64_t ldg,
                          float* crop_norm2, float* crop_dot, float* win_mean, void* stream);
int isb_region_scatter_grad(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int fh, int fw, int k,
                            const int64_t* idx, const int32_t* nsel, const float* g_u, int64_t ldg,
                            const float* crop_norm2, const float* crop_dot, const float* g_mean, float eps,
                            float* g_x, void* stream);

/* ---------------------------------------------------------------- a11 + a13
 * Negative selection of create_batch, train/siamese_regions.py:106-129 (same
 * code train/siamese_descriptor.py:108-131), for P positive couples at once and
 * without materialising S = torch.mm(E, E.t()) (utils/train_siamese.py:53):
 *   excl(j)  = label[j] == label[anchor]  ||  (semi_hard && S[anchor, j] >= S[anchor, positive])
 *   neg      = argmax_{j : !excl(j)} S[anchor, j]        (-1 when every j is excluded;
 *              the reference then draws a random negative on the host, :112-129)
 * semi_hard is the reference's `epoch < P.train_epoch_switch`.
 * The rows of a tcgen05 screen GEMM are the anchors; the masks are applied in its
 * streaming top-k epilogue; the survivors (ISB_OPT_MINING_KC per couple, default 16) are
 * re-scored exactly (fp64 accumulation) and the exact conditions re-applied -- progressively:
 * only the candidates the current exact winner does not exclude (screen score within eps of
 * it, the certificate's own bound) are gathered, usually one or two of the 16.
 * Certificate per couple: every column outside the candidate list has a screen score <= the
 * worst candidate's (t_min), hence an exact score <= t_min + eps; the couple is certified when
 * the exact winner clears that, with
 *   eps = max(screen_eps, 8 * max(sigma_measured, sigma_floor)),
 * sigma_measured = rms(screen - exact) over the couple's own candidates.  (Statistical, not a
 * proof: 8 sigma.)  In semi-hard mode the epilogue also drops columns whose screen score is
 * >= S[anchor, positive] + max(screen_eps, 32 * sigma_floor); a couple whose eps exceeds
 * that slack is rejected too.
 *   uncertified_rows != NULL: the rejected couples (indices p) are listed there and counted in
 *     *n_uncertified; the caller re-runs them with a finer screen (split operands) -- the
 *     second line, as isb_topk_resolve is for the search.
 *   uncertified_rows == NULL: they are recomputed by an exhaustive exact pass inside this call
 *     (counted in *n_uncertified when non-NULL).
 *   emb [N, D] fp32 (unit rows); emb_hi / emb_lo [N, ld] bf16 = isb_f32_to_bf16 parts
 *   0 / 1 of emb.  emb_lo == NULL: plain bf16 screen (one tcgen05 product per tile;
 *   screen_eps = 0, sigma_floor = the expected bf16 noise 2.34e-3 |a||b| / sqrt(D)).
 *   emb_lo != NULL: split operands, a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, an fp32-grade screen
 *   (three products per tile; screen_eps = 2e-5 absolute on unit rows, sigma_floor = 0).
 *   label [N] int32; anchors, positives [P] int64
 *   neg_idx [P] int64; neg_sim [P] fp32 (-2 when none); pos_sim [P] fp32 */
size_t isb_select_negatives_workspace_bytes(int64_t P, int64_t N, int64_t D, int split);
int isb_select_negatives(const float* emb, const uint16_t* emb_hi, const uint16_t* emb_lo, int64_t ld,
                         int64_t N, int64_t D, const int32_t* label, const int64_t* anchors,
                         const int64_t* positives, int64_t P, int semi_hard, float screen_eps,
                         float sigma_floor, int64_t* neg_idx, float* neg_sim, float* pos_sim,
                         int32_t* uncertified_rows, int32_t* n_uncertified, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- a9 (reduction)
 * precision1's row reduction, utils/metrics.py:11-13:  kth <= 1 -> sim.max(1);
 * kth > 1 -> sim.kthvalue(N - kth + 1, 1), i.e. the kth LARGEST entry of each row
 * and its column.  sim [Q, N] fp32 with leading dimension ld; val [Q]; idx [Q]
 * int64.  Ties -> lower column first.  1 <= kth <= min(64, N). */
int isb_row_kth_largest(const float* sim, int64_t Q, int64_t N, int64_t ld, int kth, float* val,
                        int64_t* idx, void* stream);

/* ---------------------------------------------------------------- a10 (ranking)
 * What avg_precision needs from  sim[i].sort(descending=True)  (utils/metrics.py:33):
 * the position of the query's positives in that ranking.  For every row q and
 * every listed column c = cols[q, p] (p < P; c < 0 marks an unused slot):
 *   rank[q, p] = #{ j : sim[q, j] > sim[q, c]  or  (sim[q, j] == sim[q, c] and j < c) }
 * (unused slots get -1).  One streaming pass per row, no sort of the row.
 * cols, rank [Q, P] int32; 1 <= P <= 8192. */
int isb_row_ranks(const float* sim, int64_t Q, int64_t N, int64_t ld, const int32_t* cols, int P,
                  int32_t* rank, void* stream);

/* ---------------------------------------------------------------- (f1) DBA
 * instance_avg, test/instance_avg.py:7-33: out[i] = agg / (||agg|| + 1e-10),
 * agg = emb[i] + sum_{j < nn} ((nn - j) / (nn + 1)) * emb[nbr_j], nbr = the other
 * items of label[i] by decreasing emb[i].emb[nbr], nn = their number (capped by k
 * when k >= 0); nn <= 0 -> out[i] = emb[i].  No N x N matrix: only same-label
 * similarities are computed (exactly).  *overflow counts rows whose label has more
 * than 2048 other members (not supported; result uses the first 2048 by index). */
int isb_instance_avg(const float* emb, const int32_t* label, int64_t N, int64_t D, int k, float* out,
                     int32_t* overflow, void* stream);

/* ---------------------------------------------------------------- (f3) training-side operators
 * NormalizeL2Fun.backward, model/custom_modules.py:59-67:
 *   grad_in = (norm2 * g - x * <x, g>) / (norm2 * sqrt(norm2)),  norm2 = sum x^2 + eps */
int isb_l2norm_rows_backward(const float* x, const float* grad_out, int64_t M, int64_t F, float eps,
                             float* grad_in, void* stream);
/* ShiftFun.backward, model/custom_modules.py:20-25: grad_param[j] = sum_m g[m, j]
 * (grad_input is grad_output itself). */
int isb_col_sums(const float* g, int64_t M, int64_t F, float* out, void* stream);
/* TripletLossFun.forward, model/custom_modules.py:153-171.  row_loss [B] (after the
 * clamp), clamp [B] uint8 (1 where the row's loss was <= 0), loss [1]. */
int isb_triplet_loss_forward(const float* anchor, const float* pos, const float* neg, int64_t B,
                             int64_t D, float margin, int size_average, int normalized, float* loss,
                             float* row_loss, uint8_t* clamp, void* stream);
/* TripletLossFun.backward, model/custom_modules.py:173-203.  grad_out [1] device. */
int isb_triplet_loss_backward(const float* anchor, const float* pos, const float* neg, int64_t B,
                              int64_t D, const uint8_t* clamp, const float* grad_out, int size_average,
                              int normalized, float* grad_anchor, float* grad_pos, float* grad_neg,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISB_H_ */
