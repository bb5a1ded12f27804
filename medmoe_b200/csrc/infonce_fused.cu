// medmoe_b200 — fused InfoNCE on the tensor cores (north-star kernel 5).
//
// Replaces reference src/losses.py:558-592 (contrastive_loss_with_temperature after the gather of :503-524):
//     logits_a = exp(logit_scale) * a  all_b^T ,  logits_b = exp(logit_scale) * b  all_a^T        (:567-572)
//     loss_a = CE(logits_a, label0 + arange) , loss_b = CE(logits_b, ...)   [optional label smoothing]   (:579-584)
// and its whole backward (d a, d b, d all_a, d all_b, d logit_scale).
//
// fp32 parity on bf16 tensor cores: every fp32 operand is split into three bf16 parts x = h + m + l (24 mantissa
// bits); a product is the six cross terms hh + hm + mh + hl + lh + mm accumulated in fp32 in TMEM (dropped terms are
// <= 2^-24 relative), so the logits agree with an fp32 GEMM to fp32 rounding.  The split is written once per step by
// `nce_split_kernel` as [rows, 3 D] bf16 (= [h | m | l]); the six terms are six k ranges of ONE tcgen05 k loop.
//
//   nce_fwd_kernel  grid (column tiles, row tiles, 2 directions), one 128 x 128 logits tile per CTA:
//       TMA (SW128) -> tcgen05.mma (6 D / 16 MMAs) -> TMEM; the epilogue (thread = row) reduces its row of the tile to
//       (max, sum exp, sum x), picks the label column and (only if the caller wants them) stores the logits; the last CTA
//       to finish combines the per-tile partials into lse[r] and the two losses in a fixed order (deterministic).
//       The logits never reach HBM unless asked for.
//   nce_bwd_kernel  same grid; recomputes its logits tile into TMEM, turns it into
//       dL = coef[r] (softmax - (1 - eps) onehot - eps / N) in registers, writes dL as a 2-way bf16 split into shared
//       memory and feeds it straight back to the tensor cores: as the K-major A operand of  d rows += t dL cols  and,
//       the SAME bytes read as an MN-major A operand (= dL^T), of  d cols += dL^T (t rows);  both accumulate in TMEM
//       192 output columns at a time (double buffered) and leave through red.global.add.v4.f32.
//
// Roles per CTA (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include "mm_common.cuh"
#include "api_internal.h"

namespace mm {

constexpr int NCE_T = 128;                 // tile edge (rows and columns of the logits tile)
constexpr int NCE_THREADS = 256;
constexpr int NCE_FWD_STAGES = 4;
constexpr int NCE_FWD_STAGE_BYTES = 32768; // A 128 x 64 bf16 + B 128 x 64 bf16
constexpr int NCE_BWD_STAGES = 3;
constexpr int NCE_ND = 192;                // output columns per accumulator of the second-stage products
constexpr int NCE_BWD_STAGE_BYTES = NCE_ND * 128 * 2;   // 128 k-rows x 192 columns bf16 = 48 KB
constexpr int NCE_DL_BYTES = 2 * 16384;    // one bf16 part of dL: two 64-column blocks of 128 rows x 128 B

struct NceArgs {
    int R, N, D;                 // local rows, gathered columns, embedding width
    int n_ct, n_rt;
    int label0;
    const float* row_w;          // [R] weights of the row losses (mask path) or nullptr (1 / R)
    float smoothing;             // label_smoothing of F.cross_entropy
    float* logits[2];            // [R, N] fp32 or nullptr
    float* part;                 // [2][R][n_ct][4]  (max, sum exp, sum x, -)
    float* picked;               // [2][R]
    float* lse;                  // [2][R]
    float* loss;                 // [2]
    unsigned* counter;
    // backward
    const float* gout[2];        // upstream gradients of loss_a / loss_b (device scalars; nullptr = 0)
    const float* scale;          // exp(logit_scale)
    float* drow[2];              // [R, D]: dir 0 -> d a, dir 1 -> d b
    float* dcol[2];              // [N, D]: dir 0 -> d all_b, dir 1 -> d all_a
    float* dscale_part;          // [2 * n_rt * n_ct]
    float* dscale;
    int accumulate_dscale;
};

// ------------------------------------------------------------------------------------
// split: x (optionally * scale) -> [h | m | l] bf16
// ------------------------------------------------------------------------------------
struct NceSplitJob { const float* src; __nv_bfloat16* dst; int rows; int scaled; };
struct NceSplitArgs { NceSplitJob job[4]; int D; const float* scale; };

__global__ void __launch_bounds__(256) nce_split_kernel(const NceSplitArgs a) {
    const NceSplitJob j = a.job[blockIdx.y];
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j.src == nullptr || r >= j.rows) return;
    const float sc = j.scaled ? __ldg(a.scale) : 1.0f;
    const float* x = j.src + static_cast<size_t>(r) * a.D;
    __nv_bfloat16* o = j.dst + static_cast<size_t>(r) * 3 * a.D;
    for (int i = lane * 4; i < a.D; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(x + i);
        const float f[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
        __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            h[t] = __float2bfloat16_rn(f[t]);
            const float r1 = f[t] - __bfloat162float(h[t]);
            m[t] = __float2bfloat16_rn(r1);
            l[t] = __float2bfloat16_rn(r1 - __bfloat162float(m[t]));
        }
        *reinterpret_cast<uint2*>(o + i) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(o + a.D + i) = *reinterpret_cast<const uint2*>(m);
        *reinterpret_cast<uint2*>(o + 2 * a.D + i) = *reinterpret_cast<const uint2*>(l);
    }
}

// the six cross terms of a 3 x 3 split product, largest first: (part of the row operand, part of the column operand)
MM_DEVINL int nce_term_a(int t) { return (0x120100 >> (4 * t)) & 15; }   // 0 0 1 0 2 1
MM_DEVINL int nce_term_b(int t) { return (0x102010 >> (4 * t)) & 15; }   // 0 1 0 2 0 1

// logits tile: TMA producer and MMA issuer, shared by the forward and the backward kernel.  Stage s of the ring starts at
// ring + s * stage_bytes; A (rows) sits at +0, B (columns) at +16384.
MM_DEVINL void nce_produce_logits(const CUtensorMap* tmRow, const CUtensorMap* tmCol, uint8_t* ring, int stage_bytes, int n_stages,
                                  uint64_t* full, uint64_t* empty, int D, int row0, int col0, int& stage, uint32_t& phase) {
    const int nkb = D >> 6;
    for (int t = 0; t < 6; ++t) {
        const int ka = nce_term_a(t) * D, kb0 = nce_term_b(t) * D;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], 32768);
            uint8_t* dst = ring + stage * stage_bytes;
            tma_load_2d(dst, tmRow, &full[stage], ka + kb * 64, row0);
            tma_load_2d(dst + 16384, tmCol, &full[stage], kb0 + kb * 64, col0);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    }
}
MM_DEVINL void nce_issue_logits(uint32_t d_tmem, uint8_t* ring, int stage_bytes, int n_stages, uint64_t* full, uint64_t* empty,
                                int D, int& stage, uint32_t& phase) {
    constexpr uint32_t idesc = make_idesc_bf16(NCE_T, NCE_T, 0, 0);
    const uint64_t d_base = make_smem_desc(smem_u32(ring), 16, 1024);
    const int total = 6 * (D >> 6);
    for (int i = 0; i < total; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t da = smem_desc_advance(d_base, stage * stage_bytes);
        const uint64_t db = smem_desc_advance(da, 16384);
        umma_bf16(d_tmem, da, db, idesc, i != 0);
        umma_bf16(d_tmem, smem_desc_advance(da, 32), smem_desc_advance(db, 32), idesc, 1);
        umma_bf16(d_tmem, smem_desc_advance(da, 64), smem_desc_advance(db, 64), idesc, 1);
        umma_bf16(d_tmem, smem_desc_advance(da, 96), smem_desc_advance(db, 96), idesc, 1);
        umma_commit(&empty[stage]);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NCE_THREADS, 1)
nce_fwd_kernel(const __grid_constant__ CUtensorMap tmRow0, const __grid_constant__ CUtensorMap tmRow1,
               const __grid_constant__ CUtensorMap tmCol0, const __grid_constant__ CUtensorMap tmCol1, const NceArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NCE_FWD_STAGES * NCE_FWD_STAGE_BYTES);
    uint64_t* empty = full + NCE_FWD_STAGES;
    uint64_t* tfull = empty + NCE_FWD_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    __shared__ float s_red[2][NCE_THREADS];
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x, rt = blockIdx.y, dir = blockIdx.z;
    const CUtensorMap* tmRow = dir ? &tmRow1 : &tmRow0;
    const CUtensorMap* tmCol = dir ? &tmCol1 : &tmCol0;

    if (threadIdx.x == 0) { tma_prefetch_desc(tmRow); tma_prefetch_desc(tmCol); }
    if (threadIdx.x == 32) {
        for (int s = 0; s < NCE_FWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, NCE_T); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            nce_produce_logits(tmRow, tmCol, ring, NCE_FWD_STAGE_BYTES, NCE_FWD_STAGES, full, empty, a.D, rt * NCE_T, ct * NCE_T, stage, phase);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            nce_issue_logits(tmem_base, ring, NCE_FWD_STAGE_BYTES, NCE_FWD_STAGES, full, empty, a.D, stage, phase);
            umma_commit(tfull);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int r = rt * NCE_T + q * 32 + lane;
        const bool row_valid = r < a.R;
        const int label = a.label0 + r;
        float mx = -INFINITY, se = 0.f, sx = 0.f, pk = 0.f;
        bool have_pk = false;
        float* lrow = a.logits[dir] ? a.logits[dir] + static_cast<size_t>(r) * a.N : nullptr;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < NCE_T / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + c * 32, v);
            tmem_ld_wait();
            const int n0 = ct * NCE_T + c * 32;
            const int nv = min(32, a.N - n0);          // valid columns of this chunk (may be <= 0)
            if (nv <= 0) continue;
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nv) cmax = fmaxf(cmax, __uint_as_float(v[j]));
            const float newm = fmaxf(mx, cmax);
            float add = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < nv) {
                    const float f = __uint_as_float(v[j]);
                    add += expf(f - newm);
                    sx += f;
                    if (n0 + j == label) { pk = f; have_pk = true; }
                }
            }
            se = se * expf(mx - newm) + add;          // mx = -inf on the first chunk: exp(-inf) = 0
            mx = newm;
            if (lrow && row_valid) {
                if (nv == 32 && (a.N & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(lrow + n0 + 4 * j) =
                            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (j < nv) lrow[n0 + j] = __uint_as_float(v[j]);
                }
            }
        }
        if (row_valid) {
            float4* p = reinterpret_cast<float4*>(a.part) + (static_cast<size_t>(dir) * a.R + r) * a.n_ct + ct;
            *p = make_float4(mx, se, sx, 0.f);
            if (have_pk) a.picked[dir * a.R + r] = pk;
        }
        tc_fence_before();
    }

    // ---- the last CTA to finish turns the partials into lse[r] and the two losses (fixed order: deterministic) ----
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, NCE_T);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        s_last = (atomicAdd(a.counter, 1u) == total - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float acc[2] = {0.f, 0.f};
    for (int idx = threadIdx.x; idx < 2 * a.R; idx += NCE_THREADS) {
        const int d = idx / a.R, r = idx - d * a.R;
        const float4* p = reinterpret_cast<const float4*>(a.part) + static_cast<size_t>(idx) * a.n_ct;
        float M = -INFINITY;
        for (int t = 0; t < a.n_ct; ++t) M = fmaxf(M, __ldcg(&p[t].x));
        float S = 0.f, X = 0.f;
        for (int t = 0; t < a.n_ct; ++t) {
            const float4 e = __ldcg(p + t);
            if (e.y > 0.f) S += e.y * expf(e.x - M);
            X += e.z;
        }
        const float lse = M + logf(S);
        a.lse[idx] = lse;
        const float w = a.row_w ? a.row_w[r] : 1.0f / static_cast<float>(a.R);
        const float row_loss = lse - (1.0f - a.smoothing) * __ldcg(a.picked + idx) - (a.smoothing / static_cast<float>(a.N)) * X;
        acc[d] += w * row_loss;
    }
    s_red[0][threadIdx.x] = acc[0];
    s_red[1][threadIdx.x] = acc[1];
    __syncthreads();
    for (int o = NCE_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_red[0][threadIdx.x] += s_red[0][threadIdx.x + o]; s_red[1][threadIdx.x] += s_red[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.loss[0] = s_red[0][0];
        a.loss[1] = s_red[1][0];
        *a.counter = 0u;          // ready for the next launch (CUDA-graph replays included)
    }
}

// ------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NCE_THREADS, 1)
nce_bwd_kernel(const __grid_constant__ CUtensorMap tmRow0, const __grid_constant__ CUtensorMap tmRow1,
               const __grid_constant__ CUtensorMap tmCol0, const __grid_constant__ CUtensorMap tmCol1, const NceArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint8_t* sDL = smem + NCE_BWD_STAGES * NCE_BWD_STAGE_BYTES;        // [2 parts (h, m)][2 column blocks][128 rows x 128 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(sDL + 2 * NCE_DL_BYTES);
    uint64_t* empty = full + NCE_BWD_STAGES;
    uint64_t* sfull = empty + NCE_BWD_STAGES;      // logits tile complete
    uint64_t* dlready = sfull + 1;                 // dL written to shared memory (4 epilogue warps)
    uint64_t* tfull = dlready + 1;                 // [2]
    uint64_t* tempty = tfull + 2;                  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    __shared__ float s_ds[4];
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x, rt = blockIdx.y, dir = blockIdx.z;
    const CUtensorMap* tmRow = dir ? &tmRow1 : &tmRow0;
    const CUtensorMap* tmCol = dir ? &tmCol1 : &tmCol0;
    const int n_dch = a.D / NCE_ND;                // output column chunks of the second-stage products
    const int n_groups = 2 * n_dch;                // kind 0: d rows (dL * cols), kind 1: d cols (dL^T * rows)

    if (threadIdx.x == 0) { tma_prefetch_desc(tmRow); tma_prefetch_desc(tmCol); }
    if (threadIdx.x == 32) {
        for (int s = 0; s < NCE_BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(sfull, 1);
        mbar_init(dlready, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;       // [0, 128): logits tile; [128, 320), [320, 512): second-stage accumulators

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            nce_produce_logits(tmRow, tmCol, ring, NCE_BWD_STAGE_BYTES, NCE_BWD_STAGES, full, empty, a.D, rt * NCE_T, ct * NCE_T, stage, phase);
            // second stage: B operand = 128 k-rows x 192 columns of the h / m part of the column (kind 0) or row (kind 1) operand
            for (int g = 0; g < n_groups; ++g) {
                const int kind = g / n_dch, d0 = (g - kind * n_dch) * NCE_ND;
                const CUtensorMap* tm = kind == 0 ? tmCol : tmRow;
                const int krow0 = kind == 0 ? ct * NCE_T : rt * NCE_T;
                for (int t = 0; t < 3; ++t) {
                    const int pb = (t == 1) ? 1 : 0;              // terms (dL_h, B_h), (dL_h, B_m), (dL_m, B_h)
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], NCE_BWD_STAGE_BYTES);
                    uint8_t* dst = ring + stage * NCE_BWD_STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < NCE_ND / 64; ++j)
                        tma_load_2d(dst + j * 16384, tm, &full[stage], pb * a.D + d0 + j * 64, krow0);
                    if (++stage == NCE_BWD_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            nce_issue_logits(tmem_base, ring, NCE_BWD_STAGE_BYTES, NCE_BWD_STAGES, full, empty, a.D, stage, phase);
            umma_commit(sfull);
            mbar_wait(dlready, 0);
            tc_fence_after();
            constexpr uint32_t idesc_k = make_idesc_bf16(NCE_T, NCE_ND, 0, 1);     // A = dL   (K-major),  B MN-major
            constexpr uint32_t idesc_t = make_idesc_bf16(NCE_T, NCE_ND, 1, 1);     // A = dL^T (MN-major), B MN-major
            const uint64_t db_base = make_smem_desc(smem_u32(ring), 16384, 1024);
            for (int g = 0; g < n_groups; ++g) {
                const int kind = g / n_dch;
                const int acc = g & 1;
                mbar_wait(&tempty[acc], ((g >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + NCE_T + acc * NCE_ND;
                for (int t = 0; t < 3; ++t) {
                    const int pl = (t == 2) ? 1 : 0;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t db = smem_desc_advance(db_base, stage * NCE_BWD_STAGE_BYTES);
                    const uint32_t dl_addr = smem_u32(sDL + pl * NCE_DL_BYTES);
                    if (kind == 0) {
                        // K = 128 logits columns: two 64-wide blocks (16 KB apart), four 32-byte steps each
                        const uint64_t da = make_smem_desc(dl_addr, 16, 1024);
#pragma unroll
                        for (int s = 0; s < 8; ++s)
                            umma_bf16(d_tmem, smem_desc_advance(da, (s >> 2) * 16384 + (s & 3) * 32), smem_desc_advance(db, s * 2048), idesc_k,
                                      (t | s) != 0);
                    } else {
                        // K = 128 logits rows: 16 rows (2048 B) per step; M = 128 logits columns = two 64-wide chunks (LBO)
                        const uint64_t da = make_smem_desc(dl_addr, 16384, 1024);
#pragma unroll
                        for (int s = 0; s < 8; ++s)
                            umma_bf16(d_tmem, smem_desc_advance(da, s * 2048), smem_desc_advance(db, s * 2048), idesc_t, (t | s) != 0);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == NCE_BWD_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int rl = q * 32 + lane;              // row of the tile == TMEM lane
        const int r = rt * NCE_T + rl;
        const bool row_valid = r < a.R;
        const int label = a.label0 + r;
        const float g_up = a.gout[dir] ? __ldg(a.gout[dir]) : 0.f;
        const float coef = row_valid ? g_up * (a.row_w ? __ldg(a.row_w + r) : 1.0f / static_cast<float>(a.R)) : 0.f;
        const float lse = row_valid ? __ldg(a.lse + dir * a.R + r) : 0.f;
        const float t_scale = __ldg(a.scale);
        const float hit = 1.0f - a.smoothing, base = a.smoothing / static_cast<float>(a.N);
        float ds = 0.f;
        mbar_wait(sfull, 0);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(rl & ~31) << 16);
#pragma unroll 1
        for (int c = 0; c < NCE_T / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + c * 32, v);
            tmem_ld_wait();
            const int n0 = ct * NCE_T + c * 32;
            uint32_t ph[16], pm[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float d2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float f = __uint_as_float(v[j + u]);
                    float d = 0.f;
                    if (row_valid && n0 + j + u < a.N) {
                        d = coef * (expf(f - lse) - ((n0 + j + u == label) ? hit : 0.f) - base);
                        ds = fmaf(d, f, ds);
                    }
                    d2[u] = d;
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(d2[0], d2[1]);
                const float2 hf = __bfloat1622float2(h2);
                ph[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                pm[j >> 1] = pack_bf16x2(d2[0] - hf.x, d2[1] - hf.y);
            }
            // K-major SWIZZLE_128B: block (c >> 1) of 64 columns, row rl, 16-byte unit u ^ (rl & 7)
            uint8_t* blk = sDL + (c >> 1) * 16384 + rl * 128;
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                const int unit = ((c & 1) * 4 + u4) ^ (rl & 7);
                *reinterpret_cast<uint4*>(blk + unit * 16) = make_uint4(ph[4 * u4], ph[4 * u4 + 1], ph[4 * u4 + 2], ph[4 * u4 + 3]);
                *reinterpret_cast<uint4*>(blk + NCE_DL_BYTES + unit * 16) = make_uint4(pm[4 * u4], pm[4 * u4 + 1], pm[4 * u4 + 2], pm[4 * u4 + 3]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dlready);
        ds = warp_sum(ds);
        if (lane == 0) s_ds[q] = ds;

        // second-stage accumulators -> red.add into the gradient buffers
        for (int g = 0; g < n_groups; ++g) {
            const int kind = g / n_dch, d0 = (g - kind * n_dch) * NCE_ND;
            const int acc = g & 1;
            mbar_wait(&tfull[acc], (g >> 1) & 1);
            tc_fence_after();
            const int orow = (kind == 0 ? rt : ct) * NCE_T + rl;
            const bool ok = orow < (kind == 0 ? a.R : a.N);
            float* out = (kind == 0 ? a.drow[dir] : a.dcol[dir]);
            const float mul = kind == 0 ? t_scale : 1.0f;       // the row operand's split already carries exp(logit_scale)
            float* orow_p = out ? out + static_cast<size_t>(orow) * a.D + d0 : nullptr;
#pragma unroll 1
            for (int c = 0; c < NCE_ND / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + NCE_T + acc * NCE_ND + c * 32, v);
                tmem_ld_wait();
                if (ok && orow_p) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        red_add_v4_f32(orow_p + c * 32 + 4 * j, mul * __uint_as_float(v[4 * j]), mul * __uint_as_float(v[4 * j + 1]),
                                       mul * __uint_as_float(v[4 * j + 2]), mul * __uint_as_float(v[4 * j + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    // d logit_scale = sum dL * L (d exp(s) / d s = exp(s), and L already carries exp(s)); per-CTA partials, last CTA adds them up
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned me = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        a.dscale_part[me] = (s_ds[0] + s_ds[1]) + (s_ds[2] + s_ds[3]);
        __threadfence();
        s_last = (atomicAdd(a.counter, 1u) == total - 1) ? 1 : 0;
        if (s_last) {
            __threadfence();
            float sum = 0.f;
            for (unsigned i = 0; i < total; ++i) sum += __ldcg(a.dscale_part + i);
            if (a.dscale) a.dscale[0] = a.accumulate_dscale ? a.dscale[0] + sum : sum;
            *a.counter = 0u;
        }
    }
}

static int nce_maps(CUtensorMap* row, CUtensorMap* col, const void* ws_row, const void* ws_col, int R, int N, int D, const char* what) {
    int rc = encode_tmap_bf16(row, ws_row, static_cast<uint64_t>(3) * D, static_cast<uint64_t>(R), static_cast<uint64_t>(3) * D, 64, NCE_T, what);
    if (rc) return rc;
    return encode_tmap_bf16(col, ws_col, static_cast<uint64_t>(3) * D, static_cast<uint64_t>(N), static_cast<uint64_t>(3) * D, 64, NCE_T, what);
}

}  // namespace mm

using namespace mm;

// workspace layout (bytes, every block 1 KB aligned):
//   split(t a) [R, 3D] bf16 | split(t b) [R, 3D] | split(all_b) [N, 3D] | split(all_a) [N, 3D] | part [2][R][n_ct] float4 |
//   picked [2][R] | dscale_part [2 n_rt n_ct] | counter
struct NceLayout { size_t row[2], col[2], part, picked, dpart, counter, total; int n_ct, n_rt; };
static NceLayout nce_layout(int R, int N, int D) {
    NceLayout L;
    auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
    L.n_ct = (N + NCE_T - 1) / NCE_T;
    L.n_rt = (R + NCE_T - 1) / NCE_T;
    size_t off = 0;
    const size_t rb = up(static_cast<size_t>(R) * 3 * D * 2), cb = up(static_cast<size_t>(N) * 3 * D * 2);
    L.row[0] = off; off += rb;
    L.row[1] = off; off += rb;
    L.col[0] = off; off += cb;
    L.col[1] = off; off += cb;
    L.part = off; off += up(static_cast<size_t>(2) * R * L.n_ct * 16);
    L.picked = off; off += up(static_cast<size_t>(2) * R * 4);
    L.dpart = off; off += up(static_cast<size_t>(2) * L.n_rt * L.n_ct * 4);
    L.counter = off; off += 1024;
    L.total = off;
    return L;
}

extern "C" int mm_infonce_fused_supported(int R, int N, int D) {
    return (R > 0 && N > 0 && D > 0 && D % 64 == 0 && D % NCE_ND == 0) ? 1 : 0;
}

extern "C" long long mm_infonce_fused_workspace_bytes(int R, int N, int D) {
    return static_cast<long long>(nce_layout(R, N, D).total);
}

// Forward of both directions in two launches (split + fused GEMM / log-sum-exp / label pick / loss).
//   a, b [R, D] fp32 (this rank's embeddings); all_a, all_b [N, D] fp32 (gathered; may alias a / b when N == R);
//   labels are label0 + r;  row_w [R] or NULL (1 / R);  label_smoothing as in F.cross_entropy;
//   logits_a / logits_b [R, N] fp32 or NULL (not materialised);  lse [2, R];  loss [2] = (loss_a, loss_b).
//   `workspace` (mm_infonce_fused_workspace_bytes, 1 KB aligned, contents need not be initialised) keeps the bf16 splits
//   and lse for the backward.
extern "C" int mm_infonce_fused_fwd(const float* a, const float* b, const float* all_a, const float* all_b, int R, int N, int D,
                                    const float* logit_scale_exp, int label0, const float* row_w, float label_smoothing,
                                    void* workspace, float* logits_a, float* logits_b, float* lse, float* loss, void* stream) {
    MM_REQUIRE(a && b && all_a && all_b && logit_scale_exp && workspace && lse && loss, MM_ERR_BAD_SHAPE, "mm_infonce_fused_fwd: null operand");
    MM_REQUIRE(mm_infonce_fused_supported(R, N, D), MM_ERR_UNSUPPORTED, "mm_infonce_fused_fwd: D must be a multiple of 192");
    MM_REQUIRE(label0 >= 0 && label0 + R <= N, MM_ERR_BAD_SHAPE, "mm_infonce_fused_fwd: label offset out of range");
    MM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, MM_ERR_MISALIGNED, "mm_infonce_fused_fwd: workspace must be 1 KB aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const NceLayout L = nce_layout(R, N, D);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    NceSplitArgs sp{};
    sp.D = D; sp.scale = logit_scale_exp;
    sp.job[0] = {a, reinterpret_cast<__nv_bfloat16*>(ws + L.row[0]), R, 1};
    sp.job[1] = {b, reinterpret_cast<__nv_bfloat16*>(ws + L.row[1]), R, 1};
    sp.job[2] = {all_b, reinterpret_cast<__nv_bfloat16*>(ws + L.col[0]), N, 0};
    sp.job[3] = {all_a, reinterpret_cast<__nv_bfloat16*>(ws + L.col[1]), N, 0};
    if (cudaMemsetAsync(ws + L.counter, 0, 4, st) != cudaSuccess) {     // a fresh workspace; the kernels re-arm the counter themselves
        set_error("mm_infonce_fused_fwd: cudaMemsetAsync failed");
        return MM_ERR_CUDA;
    }
    nce_split_kernel<<<dim3((N + 7) / 8, 4), 256, 0, st>>>(sp);
    note_launches(1);
    CUtensorMap row[2], col[2];
    for (int d = 0; d < 2; ++d)
        if (int rc = nce_maps(&row[d], &col[d], ws + L.row[d], ws + L.col[d], R, N, D, "mm_infonce_fused_fwd")) return rc;
    NceArgs g{};
    g.R = R; g.N = N; g.D = D; g.n_ct = L.n_ct; g.n_rt = L.n_rt; g.label0 = label0; g.row_w = row_w; g.smoothing = label_smoothing;
    g.logits[0] = logits_a; g.logits[1] = logits_b;
    g.part = reinterpret_cast<float*>(ws + L.part);
    g.picked = reinterpret_cast<float*>(ws + L.picked);
    g.lse = lse; g.loss = loss;
    g.counter = reinterpret_cast<unsigned*>(ws + L.counter);
    constexpr int SMEM = NCE_FWD_STAGES * NCE_FWD_STAGE_BYTES + 256 + 1024;
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        if (cudaFuncSetAttribute(nce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            set_error("mm_infonce_fused_fwd: cannot opt in to %d B of shared memory", SMEM);
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    nce_fwd_kernel<<<dim3(L.n_ct, L.n_rt, 2), NCE_THREADS, SMEM, st>>>(row[0], row[1], col[0], col[1], g);
    note_launches(1);
    return mm_check_launch("mm_infonce_fused_fwd");
}

// Backward of both directions in one launch (+ one memset of the gradient buffers by the caller-provided pointers):
//   d a, d b [R, D] and d all_a, d all_b [N, D] fp32 are ACCUMULATED (red.add): the caller zero-fills them; d all_a may alias
//   d a (and d all_b alias d b) when N == R, which sums the two contributions of a single-rank run in place.
//   g_a, g_b: device scalars (upstream gradients of loss_a / loss_b), NULL = that direction is skipped.
extern "C" int mm_infonce_fused_bwd(int R, int N, int D, const float* logit_scale_exp, int label0, const float* row_w,
                                    float label_smoothing, void* workspace, const float* lse, const float* g_a, const float* g_b,
                                    float* da, float* db, float* dall_a, float* dall_b, float* dscale, int accumulate_dscale,
                                    void* stream) {
    MM_REQUIRE(workspace && lse && logit_scale_exp, MM_ERR_BAD_SHAPE, "mm_infonce_fused_bwd: null operand");
    MM_REQUIRE(mm_infonce_fused_supported(R, N, D), MM_ERR_UNSUPPORTED, "mm_infonce_fused_bwd: D must be a multiple of 192");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const NceLayout L = nce_layout(R, N, D);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    CUtensorMap row[2], col[2];
    for (int d = 0; d < 2; ++d)
        if (int rc = nce_maps(&row[d], &col[d], ws + L.row[d], ws + L.col[d], R, N, D, "mm_infonce_fused_bwd")) return rc;
    NceArgs g{};
    g.R = R; g.N = N; g.D = D; g.n_ct = L.n_ct; g.n_rt = L.n_rt; g.label0 = label0; g.row_w = row_w; g.smoothing = label_smoothing;
    g.lse = const_cast<float*>(lse);
    g.counter = reinterpret_cast<unsigned*>(ws + L.counter);
    g.gout[0] = g_a; g.gout[1] = g_b;
    g.scale = logit_scale_exp;
    g.drow[0] = da; g.drow[1] = db;
    g.dcol[0] = dall_b; g.dcol[1] = dall_a;
    g.dscale_part = reinterpret_cast<float*>(ws + L.dpart);
    g.dscale = dscale; g.accumulate_dscale = accumulate_dscale;
    constexpr int SMEM = NCE_BWD_STAGES * NCE_BWD_STAGE_BYTES + 2 * NCE_DL_BYTES + 256 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        if (cudaFuncSetAttribute(nce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            set_error("mm_infonce_fused_bwd: cannot opt in to %d B of shared memory", SMEM);
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    nce_bwd_kernel<<<dim3(L.n_ct, L.n_rt, 2), NCE_THREADS, SMEM, st>>>(row[0], row[1], col[0], col[1], g);
    note_launches(1);
    return mm_check_launch("mm_infonce_fused_bwd");
}
