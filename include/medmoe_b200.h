/* medmoe_b200 — C ABI of the B200-native (sm_100a) MoE + contrastive-loss hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (shivangchopra11/MedMoE) is pure Python and has no FFI of its own; each entry point
 * below names the reference code it replaces (paths relative to the reference root).
 * A Python host binds these with ctypes (medmoe_b200/_lib.py); INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it is documented as "host"; the caller owns
 *     all memory (incl. workspaces) — the library allocates nothing and keeps no pointers;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises;
 *   - return value: 0 = ok, negative = mm_status below; mm_last_error() gives the text
 *     (thread-local: autograd's backward runs on a worker thread);
 *   - bf16 buffers are passed as void*; "rows" buffers use the expert-sorted, 128-row
 *     padded layout produced by mm_dispatch_build.
 */
#ifndef MEDMOE_B200_H
#define MEDMOE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum mm_status {
    MM_OK = 0,
    MM_ERR_BAD_SHAPE = -1,
    MM_ERR_MISALIGNED = -2,
    MM_ERR_UNSUPPORTED = -3,
    MM_ERR_CUDA = -4,
    MM_ERR_NO_DEVICE = -5,
    MM_ERR_WORKSPACE = -6
};

enum mm_epilogue_flags {
    MM_EPI_RELU = 1, MM_EPI_ZERO_PAD = 2, MM_EPI_CAP_SOFTMAX = 4,
    /* the caller guarantees that the tiles (2j, 2j + 1) of the launch never belong to two experts (the 256-row segment
     * alignment mm_dispatch_build produces; always true without tile_info): CTA pairs (tcgen05 cta_group::2) may then
     * share the weight tiles — each CTA stages half.  Results are bit-identical. */
    MM_EPI_PAIR_OK = 8
};

const char* mm_last_error(void);
int mm_abi_version(void);
int mm_device_sm_count(void);
/* kernels launched through this library since it was loaded (bench evidence). */
long long mm_launch_count(void);
/* optional per-kernel CUDA-event timing inside the composite entry points (bench.py): enable, run, collect
 * "name calls total_ms" lines (mm_trace_collect synchronises the device; returns bytes written). */
void mm_trace_enable(int on);
int mm_trace_collect(char* buf, int len);

/* ---- (1) router gate: softmax + top-k -------------------------------------------------
 * replaces src/models/components/swin.py:98-100 (router MLP, softmax, argmax).
 * x [B, D] fp32; W1 [128, D], b1 [128], W2 [K, 128], b2 [K] fp32.
 * out: hidden [B, 128] (post-ReLU, saved for backward), probs [B, K], topk_idx [B, topk] int32
 * (first maximum wins and NaN counts as the maximum, as torch.argmax), topk_w [B, topk] (1.0 when topk == 1,
 * else the selected probs renormalised to sum 1); near_tie [B] int32 (may be NULL): 1 where the gap between the
 * last selected probability and the best one left out is below tie_tol (the north star's "near-tie" report). */
int mm_router_topk(const float* x, int B, int D, const float* W1, const float* b1, const float* W2, const float* b2,
                   int K, int topk, float* hidden, float* probs, int32_t* topk_idx, float* topk_w, int32_t* near_tie,
                   float tie_tol, void* stream);

/* gradient of the returned probabilities (the only gradient path into the router:
 * src/models/medmoe_module.py:235-237 applies cross-entropy to them).
 * out: dlogit [B, K], dhidden [B, 128] scratch; dx [B, D] (may be NULL); dW1, db1, dW2, db2. */
int mm_router_bwd(const float* dprobs, const float* probs, const float* hidden, const float* x, const float* W1,
                  const float* W2, int B, int D, int K, float* dlogit, float* dhidden, float* dx, float* dW1,
                  float* db1, float* dW2, float* db2, void* stream);

/* ---- (2) dispatch ---------------------------------------------------------------------
 * replaces the dense "run all experts, stack, gather" of swin.py:105-108 with a counting
 * sort of the (image, choice) items by expert.  All arrays int32.
 *   P[S], region_base[S], region_tiles[S], chunk_base[S], chunk_cap[S], chunk_tiles[S] : HOST arrays
 *   out: counts [K], offsets [K+1], perm [n] slot->item, inv_perm [n] item->slot, slot_expert [n],
 *        seg_start [S, K] first global row of expert e in region s, slot_row [S, n],
 *        tile_info [sum region_tiles][2] = {expert | -1, valid rows}, chunks [sum chunk_cap][4] =
 *        {expert, first tile, tiles, scale}. */
int mm_dispatch_build(const int32_t* item_expert, int n_items, int K, int S, const int32_t* P,
                      const int32_t* region_base, const int32_t* region_tiles, const int32_t* chunk_base,
                      const int32_t* chunk_cap, const int32_t* chunk_tiles, int32_t* counts, int32_t* offsets,
                      int32_t* perm, int32_t* inv_perm, int32_t* slot_expert, int32_t* seg_start, int32_t* slot_row,
                      int32_t* tile_info, int32_t* chunks, void* stream);

/* permute the S stage-feature tensors (swin.py:139 `stage_feats`, [B, P_s, D_s], fp32 or bf16)
 * into expert-sorted bf16 row segments, zeroing segment padding.  src/dst: HOST arrays of S device pointers. */
int mm_dispatch_rows(const void* const* src, int src_is_f32, void* const* dst, int n_items, int topk, int K, int S,
                     const int32_t* P, const int32_t* D, const int32_t* region_base, const int32_t* perm,
                     const int32_t* slot_row, const int32_t* counts, const int32_t* seg_start, void* stream);
/* transpose of mm_dispatch_rows for the stage-feature gradients (sums the top-k slots of an image). */
int mm_undispatch_rows(const void* const* src, void* const* dst, int dst_is_f32, int n_images, int topk, int S,
                       const int32_t* P, const int32_t* D, const int32_t* region_base, const int32_t* inv_perm,
                       const int32_t* slot_row, void* stream);

int mm_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);
int mm_transpose_cast_f32_bf16(const float* src, void* dst, int batch, int R, int C, void* stream);

/* ---- (3) grouped expert GEMMs on tcgen05 / TMEM / TMA ----------------------------------
 * replaces Expert.proj_convs (Conv1d k=1 + ReLU, swin.py:18-23,41) and Expert.attn_proj[0]
 * (Linear(768,384), swin.py:26-27,63) and their autograd.
 *   out[rows, N] = epi(out_scale * A[rows, K] W_e[N, K]^T + bias_e + aux) masked by gate > 0
 * A, W, aux, gate bf16; out bf16 (or fp32 when out_f32); bias, colsum fp32 [E, N].
 * tile_info == NULL: one dense problem of M rows with expert 0. */
int mm_grouped_gemm_rows(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                         long long ldw, const int32_t* tile_info, int tile_begin, int tile_count, int M,
                         const float* bias, const void* aux, long long ld_aux, const void* gate, long long ld_gate,
                         void* out, long long ld_out, int out_f32, float* colsum, float out_scale, int flags,
                         void* stream);
/* the same product with a rank-1 aux that is never materialised (backward of E4 + ReLU of E1; the rank-1 term is the
 * per-image constant part of d fused / d Y, see mm_interp_softmax_combine_bwd_tc), plus an optional tensor aux:
 *   out[row, :] = (A[row, :] W_e^T + row_coef[row] * vecs[row_vec[row], :] + aux[row, :]) masked by gate > 0
 * row_coef fp32 [rows], row_vec int32 [rows], vecs fp32 [n_vecs, ld_vecs]; aux bf16 or NULL; colsum as above.
 * flags: MM_EPI_PAIR_OK (8) only — the tiles (2j, 2j + 1) of the launch never belong to two experts (mm_dispatch_build's
 * 256-row segment alignment), so CTA pairs (tcgen05 cta_group::2) may share the weight tiles; same results. */
int mm_grouped_gemm_rows_rank1(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                               long long ldw, const int32_t* tile_info, int tile_begin, int tile_count,
                               const float* row_coef, const int32_t* row_vec, const float* vecs, long long ld_vecs,
                               const void* aux, long long ld_aux, const void* gate, long long ld_gate, void* out,
                               long long ld_out, float* colsum, int flags, void* stream);
/* dW[e][N1, N2] += sum_{rows of expert e} A[row, N1]^T B[row, N2]  (fp32 red.add; caller zeroes dW). */
int mm_grouped_gemm_wgrad(const void* A, long long a_rows, int N1, long long lda, const void* B, long long b_rows,
                          int N2, long long ldb, const int32_t* chunks, int chunk_begin, int chunk_count,
                          int tile_base, float* out, void* stream);
/* same, plus colsum[e][i] += sum over the rows of expert e of A[row, i] (fp32 [E, N1], caller zeroes it): the bias
 * gradient that belongs to the weight gradient, from the same MMAs (a constant ones block appended to B). */
int mm_grouped_gemm_wgrad_colsum(const void* A, long long a_rows, int N1, long long lda, const void* B, long long b_rows,
                                 int N2, long long ldb, const int32_t* chunks, int chunk_begin, int chunk_count,
                                 int tile_base, float* out, float* colsum, void* stream);

/* back-to-back expert GEMMs of one scale region (csrc/b2b.cuh): the conv projection (swin.py:41) and the first attention
 * Linear (swin.py:63) in ONE kernel — every 64-column chunk of Y goes from its TMEM accumulator through bias + ReLU
 * into shared memory once, where it is both the A operand of the second GEMM and the source of the TMA store of Y:
 *   Y[rows, D] = ReLU(f[rows, K1] Wp_e[D, K1]^T + bias1_e)      Z[rows, H] = Y W1_e[H, D]^T + bias2_e      (bf16 out)
 * f / Y / Z point at the region's first row; tiles are tile_info[tile_begin .. tile_begin + tile_count).
 * Supported: D = 768, H = 384, K1 % 16 == 0, K1 <= 128 (mm_expert_b2b_fwd_supported); results are bit-identical to the
 * two mm_grouped_gemm_rows launches it replaces.
 * flags & 1: the tiles (2j, 2j + 1) of the launch never belong to two experts (the 256-row segment alignment that
 * mm_dispatch_build produces) -> CTA pairs (tcgen05 cta_group::2) share every weight tile; same results. */
int mm_expert_b2b_fwd_supported(int K1, int D, int H);
int mm_expert_b2b_fwd(const void* f, long long f_rows, int K1, long long ldf, const void* Wp, int E, int D,
                      long long ldwp, const float* bias1, const void* W1, int H, long long ldw1, const float* bias2,
                      const int32_t* tile_info, int tile_begin, int tile_count, void* Y, long long ld_y, void* Z,
                      long long ld_z, int flags, void* stream);

/* ---- finest scale without a sorted copy ---------------------------------------------------
 * P_0 = 3136 / 9216 rows per image are multiples of 64 and expert segments start on 256-row boundaries, so every 64-row
 * group of the finest region's expert-sorted row space lies inside ONE image.  mm_dispatch_group_map writes
 * g64[group] = first row of that group in IMAGE order (img * P0 + 64 j), -1 for padding; the three entry points below
 * address the reference's own [B, P0, D_0] tensors through it with 64-row (32-row) TMA boxes, which replaces the permute
 * of the finest stage feature (input of swin.py:41), its read by the conv weight gradient, and the un-permute of its
 * gradient.  top-k == 1, bf16 features. */
int mm_dispatch_group_map(const int32_t* perm, const int32_t* slot_row0, int n_items, int topk, int P0, int region_base0,
                          int n_groups, int32_t* g64, void* stream);
int mm_expert_b2b_fwd_gather(const void* f, long long f_rows, int K1, long long ldf, const void* Wp, int E, int D,
                             long long ldwp, const float* bias1, const void* W1, int H, long long ldw1, const float* bias2,
                             const int32_t* tile_info, int tile_begin, int tile_count, void* Y, long long ld_y, void* Z,
                             long long ld_z, int flags, const int32_t* f_g64, void* stream);
int mm_grouped_gemm_wgrad_colsum_gather(const void* A, long long a_rows, int N1, long long lda, const void* B, long long b_rows,
                                        int N2, long long ldb, const int32_t* chunks, int chunk_begin, int chunk_count,
                                        int tile_base, float* out, float* colsum, const int32_t* b_g64, void* stream);
int mm_grouped_gemm_rows_scatter(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                                 long long ldw, const int32_t* tile_info, int tile_begin, int tile_count, const float* bias,
                                 void* out, long long out_rows, long long ld_out, const int32_t* out_g64, int flags,
                                 void* stream);

/* ---- (4) interpolate + scale-softmax + weighted combine / scatter-back ------------------
 * replaces swin.py:42-80 (F.interpolate, stack/permute, attn_proj[1:], softmax over scales,
 * weighted sum) and swin.py:108-113 (gather, mean, local_feat layout).
 * Y [rows, D], Z [rows, D/2] bf16; w2 [E, D/2], b2 [E] fp32; Ps HOST [4].
 * out [B, P, D] (bf16 or fp32), beta [n_items, P, 4], gpart [B, nblk, D] scratch, global_feat [B, D] fp32. */
int mm_combine_num_token_blocks(int P);
int mm_combine_num_row_blocks(const int32_t* Ps);
int mm_combine_num_runs(int P);
int mm_combine_num_part_blocks(int P, const int32_t* Ps);
long long mm_combine_bwd_z_scratch_floats(int P, const int32_t* Ps, int D);   /* floats per item of `mom_z` */
/* perm / seg_start / offsets / tile_info0 (the n_tiles0 tile_info entries of the finest-scale region, whose first
 * row-space row is region0_row) / K / total_rows (rows of Y) come from mm_dispatch_build; they enable the
 * tensor-core combine (out = C * Yrows with a sparse coefficient matrix, tcgen05).  tile_info0 == NULL or
 * flags & 1 selects the CUDA-core kernel. */
int mm_interp_softmax_combine_fwd(const void* Y, const void* Z, const float* w2, const float* b2, int B, int topk, int P,
                                  const int32_t* Ps, int D, const int32_t* inv_perm, const int32_t* slot_expert,
                                  const int32_t* slot_row, const float* gate, float* beta, void* out, int out_f32,
                                  float* gpart, float* global_feat, const int32_t* perm, const int32_t* seg_start,
                                  const int32_t* offsets, const int32_t* tile_info0, int n_tiles0, int region0_row, int K,
                                  long long total_rows, int flags, void* stream);
/* backward: dlocal [B, P, D] (bf16/fp32, may be NULL), dglobal [B, D] fp32 (may be NULL) ->
 * dlogit [n_items, P, 8] scratch (dlogit, or the two column-half partial dbeta of the token-centric path), dgate [n_items] (+=, may be NULL), dUT [rows, D] bf16, dZ [rows, D/2] bf16,
 * part [n_items, mm_combine_num_part_blocks, D + 1] scratch, dw2_db1_db2 [K, D + 1] = per expert
 * {dw2 (D/2) | db1 (D/2) | db2}.  mom_u [n_items, mm_combine_num_runs, 2, D] and mom_z [.., 2, D/2] fp32 scratch enable
 * the token-centric path (integer scale ratios); NULL or force_generic selects the generic gather kernel. */
int mm_interp_softmax_combine_bwd(const void* Y, const void* Z, const float* w2, int B, int topk, int P,
                                  const int32_t* Ps, int D, int K, const int32_t* perm, const int32_t* inv_perm,
                                  const int32_t* slot_expert, const int32_t* slot_row, const int32_t* counts,
                                  const int32_t* seg_start, const int32_t* offsets, const float* gate,
                                  const float* beta, const void* dlocal, int dlocal_f32, const float* dglobal,
                                  float* dlogit, float* dgate, void* dUT, void* dZ, float* part, float* dw2_db1_db2,
                                  float* mom_u, float* mom_z, int force_generic, void* stream);

/* backward on the tensor-core / rank-1 path (even integer scale ratios, Ps[0] == P; mm_combine_bwd_*_supported).
 * dF(p) = dlocal[b, p, :] + dglobal[b, :] / P splits into
 *   - the per-image constant: d fused / d Y is rank-1 per image and never written; row_dot [rows] (scratch), row_coef [rows]
 *     fp32 and row_img [rows] int32 feed mm_grouped_gemm_rows_rank1 (vecs = dglobal);
 *   - the local part (dlocal bf16 [B, P, D]): dbeta by a tcgen05 GEMM against the staged Y rows (dbeta_loc [n_items, P, 4]
 *     scratch) and dUT [rows, D] bf16 (the aux of the dY GEMM); mom_u as in mm_interp_softmax_combine_bwd.
 * Either cotangent may be NULL (BASELINE config 2 has dlocal == NULL: the contrastive loss consumes global_feat alone).
 * tile_info0 / n_tiles0 / region0_row / total_rows as in mm_interp_softmax_combine_fwd (needed when dlocal != NULL).
 * zscr: mm_combine_bwd_z_scratch_floats floats per item. */
int mm_combine_bwd_global_supported(int P, const int32_t* Ps, int D);
int mm_combine_bwd_tc_supported(int P, const int32_t* Ps, int D);
int mm_interp_softmax_combine_bwd_tc(const void* Y, const void* Z, const float* w2, int B, int topk, int P, const int32_t* Ps,
                                     int D, int K, const int32_t* perm, const int32_t* inv_perm, const int32_t* slot_expert,
                                     const int32_t* slot_row, const int32_t* counts, const int32_t* seg_start,
                                     const int32_t* offsets, const int32_t* tile_info0, int n_tiles0, int region0_row,
                                     long long total_rows, const float* gate, const float* beta, const void* dlocal,
                                     const float* dglobal, float* row_dot, float* row_coef, int32_t* row_img,
                                     float* dbeta_loc, void* dUT, float* mom_u, float* dgate, void* dZ, float* part,
                                     float* dw2_db1_db2, float* zscr, void* stream);
/* ---- word-patch attention loss (GLORIALocalContrastiveLoss, src/losses.py:954-1026; attention_fn :698-736) ----
 * The large products run on mm_grouped_gemm_rows / mm_grouped_gemm_wgrad; these are the passes in between.  Every [*, N]
 * matrix has N = n_caps * Wp columns, column = caption * Wp + word; words >= cap_len[caption] are masked.
 *   softmax_exp_fwd: E[row, n] = exp(temp1 * softmax over the caption's words of S[row, :]) (bf16), 0 where masked
 *   softmax_exp_bwd: dE (bf16, in place) -> dS through exp, temp1 and the softmax
 *   cos_lse_fwd    : cos[b, n] = cosine(words[n, :], wcU[b, n, :]) (eps 1e-8), sim[b, caption] = log sum_w exp(temp2 cos)
 *                    (agg_mean: log mean); wcU fp32 [B, N, D], words fp32 [N, D]
 *   cos_lse_bwd    : dsim [B, n_caps] -> dwcU bf16 [B, N, D] (zeros where masked), its transpose dwcUT[(b * D + d) * ld_t + n]
 *                    (the K-major weight of the d ctx GEMM) and dwords fp32 [N, D] (the direct part); N % 16 == 0 */
/* score GEMM + first softmax in one kernel for captions padded to exactly 32 word slots:
 * E[row, 32 c + w] = exp(temp1 * softmax_w(<A[row, :], W[32 c + w, :]>)) for w < cap_len[c], 0 otherwise (bf16) */
int mm_local_scores_softmax_exp(const void* A, long long rows, int K, long long lda, const void* W, int n_caps, long long ldw,
                                const int32_t* cap_len, float temp1, void* E, long long ld_e, void* stream);
int mm_local_softmax_exp_fwd(const float* S, long long ld_s, void* E, long long ld_e, long long rows, int n_caps, int Wp,
                             const int32_t* cap_len, float temp1, void* stream);
int mm_local_softmax_exp_bwd(const void* E, long long ld_e, void* dE, long long ld_d, long long rows, int n_caps, int Wp,
                             const int32_t* cap_len, float temp1, void* stream);
int mm_local_cos_lse_fwd(const float* wcU, const float* words, int B, int n_caps, int Wp, int D, const int32_t* cap_len,
                         float temp2, int agg_mean, float* cosv, float* sim, long long ld_sim, void* stream);
int mm_local_cos_lse_bwd(const float* dsim, long long ld_dsim, const float* sim, long long ld_sim, const float* cosv,
                         const float* wcU, const float* words, int B, int n_caps, int Wp, int D, const int32_t* cap_len,
                         float temp2, int agg_mean, void* dwcU, void* dwcUT, long long ld_t, float* dwords, void* stream);
/* A/B switch of the CTA-pair (tcgen05 cta_group::2) row GEMMs, which run wherever a launch passes MM_EPI_PAIR_OK:
 * bit 0 = plain bf16 epilogue GEMMs with a 192 / 256 wide tile, bit 1 = the rank-1 (dY) GEMM; default 3, also
 * MEDMOE_GEMM_PAIR=<bits>.  (The back-to-back kernel has its own switch: MEDMOE_B2B_DEBUG bit 0 = no pairs.) */
void mm_debug_gemm_pair(int on);
/* test hook: 1 = compute dUT with the CUDA-core kernel instead of the tcgen05 one (process-wide) */
void mm_debug_force_cuda_core_dut(int on);
/* the dlocal == NULL special case of the above (kept as its own entry point) */
int mm_interp_softmax_combine_bwd_global(const void* Y, const void* Z, const float* w2, int B, int topk, int P,
                                         const int32_t* Ps, int D, int K, const int32_t* perm, const int32_t* inv_perm,
                                         const int32_t* slot_expert, const int32_t* slot_row, const int32_t* counts,
                                         const int32_t* seg_start, const int32_t* offsets, const float* gate,
                                         const float* beta, const float* dglobal, float* row_dot, float* row_coef,
                                         int32_t* row_img, float* dgate, void* dZ, float* part, float* dw2_db1_db2,
                                         float* zscr, void* stream);

/* ---- (5) global contrastive loss -------------------------------------------------------
 * GLORIA semantics: replaces GLORIAGlobalContrastiveLoss.forward, src/losses.py:766-794.
 * img, txt [B, D] fp32; ws = mm_gloria_workspace_floats(B) floats kept from fwd to bwd. */
long long mm_gloria_workspace_floats(int B);
int mm_gloria_global_fwd(const float* img, const float* txt, int B, int D, float temp, float eps, float* ws,
                         float* loss, void* stream);
int mm_gloria_global_bwd(const float* img, const float* txt, int B, int D, float temp, float eps, float* ws,
                         const float* gout, float* dimg, float* dtxt, void* stream);
/* every expert's fp32 master parameters (separate storages under the reference's names, swin.py:18-30) -> the stacked
 * operands the kernels read, in one launch: kind 0 = weight [rows, cols] -> bf16 copy dst and (dstT non-NULL) bf16
 * transpose [cols, rows]; kind 1 = fp32 vector copy.  src / dst / dstT are HOST arrays of n_jobs DEVICE pointers. */
int mm_pack_expert_params(const void* const* src, void* const* dst, void* const* dstT, const int32_t* rows,
                          const int32_t* cols, const int32_t* kind, int n_jobs, void* stream);

/* ---- (5) fused InfoNCE on the tensor cores (infonce_fused.cu) ---------------------------
 * replaces src/losses.py:558-592 after the gather of :503-524, both directions at once:
 *   logits_a = exp(logit_scale) a all_b^T, logits_b = exp(logit_scale) b all_a^T, labels label0 + r,
 *   loss[0] = sum_r w_r CE(logits_a[r]), loss[1] likewise for logits_b (w = row_w or 1 / R; label_smoothing as in
 *   F.cross_entropy, losses.py:579-583 cross_entropy_kwargs).
 * fp32 in / out; the products run on tcgen05 as 3-way bf16 splits (six cross terms, fp32-exact to rounding).
 * The logits are written only when logits_a / logits_b are non-NULL.  a, b [R, D]; all_a, all_b [N, D] (may alias a / b);
 * lse [2, R]; workspace: mm_infonce_fused_workspace_bytes, 1 KB aligned, kept by the caller for the backward.
 * Needs D % 192 == 0 (mm_infonce_fused_supported); other widths use mm_infonce_fwd / mm_infonce_bwd. */
int mm_infonce_fused_supported(int R, int N, int D);
long long mm_infonce_fused_workspace_bytes(int R, int N, int D);
int mm_infonce_fused_fwd(const float* a, const float* b, const float* all_a, const float* all_b, int R, int N, int D,
                         const float* logit_scale_exp, int label0, const float* row_w, float label_smoothing,
                         void* workspace, float* logits_a, float* logits_b, float* lse, float* loss, void* stream);
/* backward of both directions in one launch: recomputes each logits tile, forms dL in registers and feeds it back to the
 * tensor cores from shared memory.  da, db [R, D], dall_a, dall_b [N, D] fp32 are ACCUMULATED (caller zero-fills them;
 * dall_a may alias da and dall_b alias db when N == R); g_a, g_b: device scalars (upstream gradients), NULL = 0;
 * dscale[0] (+)= d loss / d logit_scale. */
int mm_infonce_fused_bwd(int R, int N, int D, const float* logit_scale_exp, int label0, const float* row_w,
                         float label_smoothing, void* workspace, const float* lse, const float* g_a, const float* g_b,
                         float* da, float* db, float* dall_a, float* dall_b, float* dscale, int accumulate_dscale,
                         void* stream);

/* FLAVA / CLIP semantics, one direction: replaces contrastive_loss_with_temperature,
 * src/losses.py:527-592 (the all-gather of :503-524 stays in torch.distributed).
 * logits [R, N] = *logit_scale_exp * a b_all^T; labels = label0 + row; loss = sum_r w_r (lse_r - logits[r, label]),
 * w == NULL means 1/R. */
int mm_infonce_fwd(const float* a, const float* b_all, int R, int N, int D, const float* logit_scale_exp, int label0,
                   const float* row_w, float* logits, float* lse, float* picked, float* loss, void* stream);
int mm_infonce_bwd(const float* a, const float* b_all, int R, int N, int D, const float* logit_scale_exp, int label0,
                   const float* row_w, const float* logits, const float* lse, const float* gout, float gmul,
                   float* dlogits, float* row_tmp, float* da, float* db_all, float* dscale, int accumulate_dscale,
                   void* stream);
int mm_l2_normalize_fwd(const float* x, int R, int D, float eps, float* y, float* norms, void* stream);
int mm_l2_normalize_bwd(const float* dy, const float* y, const float* norms, int R, int D, float eps, float* dx,
                        void* stream);

/* ---- zero-shot classification (BASELINE config 5; src/eval_zs.py is empty in the reference,
 * semantics from SURVEY §8a row Z): pred[m] = argmax_c cos(img_m, txt_c), first maximum wins. */
int mm_zeroshot_argmax(const float* img, const float* txt, int M, int C, int D, float eps, long long* pred, float* sim,
                       void* stream);

/* ---- embedding exchange over NVLink peer memory (csrc/p2p.cu): the all-gather / reduce-scatter pair of
 * reference src/utils/distributed.py:28-58 (torch.distributed.nn.functional.all_gather and its backward) as used by
 * src/losses.py:503-524, for the ranks of ONE node (world <= 8), without NCCL: peer stores / peer loads + system-scope
 * flags.  Every rank allocates one workspace of mm_p2p_workspace_bytes(world, bytes_per_rank) with mm_p2p_alloc, ships
 * the 64-byte CUDA IPC handle to its peers (any host channel) and maps theirs with mm_p2p_open; peer_bufs is a HOST
 * array of `world` device pointers (entry `rank` = the own workspace).  Calls must be made by all ranks in the same
 * order; step counters live in the workspace and are advanced on the device (CUDA-graph replay safe).
 *   mm_p2p_all_gather:          src [bytes_per_rank] -> dst [world * bytes_per_rank], block q = rank q's src
 *   mm_p2p_reduce_scatter_f32:  grad fp32 [world * bytes_per_rank] -> out [bytes_per_rank] = sum_q grad_q[block rank]
 *                               (ranks summed in ascending order: deterministic)
 * bytes_per_rank must be a multiple of 16 and the same in every call on a workspace.  A rank that waits ~10 s for a
 * peer traps (sticky CUDA error) instead of hanging. */
long long mm_p2p_workspace_bytes(int world, long long bytes_per_rank);
int mm_p2p_alloc(long long bytes, void** ptr_out, void* handle_out);
int mm_p2p_open(const void* handle, void** ptr_out);
int mm_p2p_close(void* ptr);
int mm_p2p_free(void* ptr);
int mm_p2p_all_gather(const void* src, void* dst, long long bytes_per_rank, void* const* peer_bufs, int rank, int world,
                      void* stream);
int mm_p2p_reduce_scatter_f32(const void* grad, void* out, long long bytes_per_rank, void* const* peer_bufs, int rank,
                              int world, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDMOE_B200_H */
