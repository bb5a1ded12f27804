// medmoe_b200 — forward combine on the tensor cores (tcgen05 / TMEM / TMA).
// Included by combine.cu after CombineArgs.
//
//   fused[p, :] = sum_s beta_s(p) interp(Y_s)(p, :)        (reference swin.py:42,78-80)
//
// is a sparse matrix product: for a tile of 128 consecutive tokens (= 128 consecutive rows of the finest scale
// in the expert-sorted row space) every output row is a combination of 7 native rows
//   out[128, D] = C[128, KT] * Yrows[KT, D],   KT = 128 (scale 0, diagonal beta_0) + sum_{s>0} (128 / r_s + 2 -> mult. of 8)
// The native rows a tile touches are ONE contiguous row range per scale (items of an expert are contiguous in
// every region), so TMA stages them as the MN-major B operand (channels contiguous), the coefficient matrix C is
// built in shared memory as the K-major A operand (7 non-zeros per row at fixed positions, everything else stays
// zero), and tcgen05.mma does all the multiply-adds.  CUDA cores only compute 7 coefficients per token and run the
// epilogue (TMEM -> bf16 -> swizzled staging -> TMA store, plus the per-32-token column sums for global_feat).
// HBM traffic = Y once + out once: the kernel is bound by HBM, not by instruction issue like its CUDA-core predecessor.
//
// Roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 coefficient
// builders (thread = token), warps 8-15 epilogue (two per TMEM lane quarter).
// Requirements (host falls back to the CUDA-core kernel otherwise): topk == 1, Ps[0] == P, P % 32 == 0, every
// ratio r_s = P / Ps[s] a power of two <= 128, D % 128 == 0.
#pragma once
#include "gemm.cuh"

namespace mm {

constexpr int CM_BN = 128;              // output columns per accumulator pass
constexpr int CM_STAGES = 3;             // maximum (the IMG / top-k > 1 variant runs 2: its coefficient matrices take the room)
constexpr int CM_ACC = 4;               // TMEM ring: 4 accumulators x 128 columns
constexpr int CM_EPI_WARPS = 8;
constexpr int CM_COEF_WARPS = 4;
constexpr int CM_THREADS = (4 + CM_COEF_WARPS + CM_EPI_WARPS) * 32;

struct CmArgs {
    int n_tiles;                 // 128-row tiles of the finest-scale region
    const int2* tile_info;       // [n_tiles] {expert | -1, valid rows} of those tiles
    int region_row[4];           // first row-space row of region s
    const int* seg_start;        // [4, K]
    const int* offsets;          // [K + 1] first slot of expert e
    int K;
    int cap[4], koff[4], ktot;   // rows staged per scale, their first k index, total (multiple of 16)
    int D, n_pass;               // D / CM_BN
    int out_f32;
    // IMG tiling (top-k > 1): a tile is 128 consecutive tokens of one IMAGE; its n_src = topk expert choices are extra
    // K groups accumulated into the same TMEM tile (one coefficient matrix per choice, gate weights folded in)
    int tiles_per_img, n_src, stages;
};

MM_DEVINL int cm_floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// byte offset of element (m, k) of the K-major SWIZZLE_128B coefficient matrix (blocks of 64 k = 16 KB)
MM_DEVINL uint32_t cm_a_off(int m, int k) {
    const int kb = k >> 6, kin = k & 63;
    return static_cast<uint32_t>(kb * 16384 + m * 128 + ((((kin >> 3) ^ (m & 7))) << 4) + (kin & 7) * 2);
}

// Issue nk tcgen05.mma steps (K = 16 each) of a K-major SWIZZLE_128B A operand laid out in 64-wide k blocks of 16 KB
// (step s at + (s >> 2) * 16384 + (s & 3) * 32) against an MN-major B operand whose 16 k-rows per step are 2048 B apart.
// da0 / db0 are the descriptors of step 0; they advance by constant increments (no per-step descriptor rebuild: the
// single issuing lane is otherwise bound by its own instruction stream, ~200 cycles per MMA instead of the ~100 the pipe needs).
MM_DEVINL void cm_issue_mmas(uint32_t d_tmem, uint64_t da0, uint64_t db0, uint32_t idesc, int nk, bool accumulate_first) {
    int s = 0;
    for (; s + 4 <= nk; s += 4) {
        const uint64_t da = smem_desc_advance(da0, (s >> 2) * 16384);
        const uint64_t db = smem_desc_advance(db0, s * 2048);
        umma_bf16(d_tmem, da, db, idesc, (accumulate_first || s != 0) ? 1u : 0u);
        umma_bf16(d_tmem, smem_desc_advance(da, 32), smem_desc_advance(db, 2048), idesc, 1);
        umma_bf16(d_tmem, smem_desc_advance(da, 64), smem_desc_advance(db, 4096), idesc, 1);
        umma_bf16(d_tmem, smem_desc_advance(da, 96), smem_desc_advance(db, 6144), idesc, 1);
    }
    for (; s < nk; ++s)
        umma_bf16(d_tmem, smem_desc_advance(da0, (s >> 2) * 16384 + (s & 3) * 32), smem_desc_advance(db0, s * 2048), idesc,
                  (accumulate_first || s != 0) ? 1u : 0u);
}

template <bool OUT_F32, bool IMG>
__global__ void __launch_bounds__(CM_THREADS, 1)
cm_out_kernel(const __grid_constant__ CUtensorMap tmY0, const __grid_constant__ CUtensorMap tmY1,
              const __grid_constant__ CUtensorMap tmY2, const __grid_constant__ CUtensorMap tmY3,
              const __grid_constant__ CUtensorMap tmOut, const CombineArgs a, const CmArgs c) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = ((c.ktot + 63) / 64) * 16384;
    const int stage_bytes = c.ktot * 256;                       // 2 column chunks x ktot rows x 128 B
    const int n_src = IMG ? c.n_src : 1;
    const int n_stages = c.stages;
    uint8_t* sA = smem;                                         // [n_src] coefficient matrices
    uint8_t* sB = sA + n_src * a_bytes;
    uint8_t* sOut = sB + n_stages * stage_bytes;                // [CM_EPI_WARPS][2 slots] (bf16 output only)
    uint64_t* full = reinterpret_cast<uint64_t*>(sOut + (OUT_F32 ? 0 : CM_EPI_WARPS * 2 * EPI_SLOT_BYTES));
    uint64_t* empty = full + CM_STAGES;
    uint64_t* tfull = empty + CM_STAGES;
    uint64_t* tempty = tfull + CM_ACC;
    uint64_t* a_full = tempty + CM_ACC;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmY0); tma_prefetch_desc(&tmY1); tma_prefetch_desc(&tmY2); tma_prefetch_desc(&tmY3);
        if (!OUT_F32) tma_prefetch_desc(&tmOut);
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < CM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < CM_ACC; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], CM_EPI_WARPS); }
        mbar_init(a_full, CM_COEF_WARPS);
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // the coefficient matrix starts as all zeros; only the 7 fixed positions of every row are ever rewritten
    for (int i = threadIdx.x * 16; i < n_src * a_bytes; i += CM_THREADS * 16) *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nk = c.ktot >> 4;
    const int n_work = IMG ? a.B * c.tiles_per_img : c.n_tiles;

    if (warp == 0 && elect_one()) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < n_work; t += gridDim.x) {
            int row_s[2][4];        // first staged row of every scale, per source (n_src <= 2 is enforced by the host)
            if constexpr (IMG) {
                const int b = t / c.tiles_per_img, t0 = (t - b * c.tiles_per_img) * TILE_M;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j >= n_src) break;
                    const int slot = a.inv_perm[b * a.topk + j];
                    row_s[j][0] = a.slot_row[slot] + t0;
#pragma unroll
                    for (int s = 1; s < 4; ++s) row_s[j][s] = a.slot_row[s * a.n_items + slot] + t0 / a.ratio[s] - 1;
                }
            } else {
                const int e = c.tile_info[t].x;
                if (e < 0) continue;
                const int row0 = c.region_row[0] + t * TILE_M;
                const int rel0 = row0 - c.seg_start[e];                 // multiple of 128
                row_s[0][0] = row0;
#pragma unroll
                for (int s = 1; s < 4; ++s) row_s[0][s] = c.seg_start[s * c.K + e] + rel0 / a.ratio[s] - 1;
            }
            for (int n = 0; n < c.n_pass; ++n) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j >= n_src) break;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], static_cast<uint32_t>(stage_bytes));
                    uint8_t* dst = sB + stage * stage_bytes;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        uint8_t* d = dst + ch * (c.ktot * 128);
                        const int col = n * CM_BN + ch * 64;
                        tma_load_2d(d + c.koff[0] * 128, &tmY0, &full[stage], col, row_s[j][0]);
                        tma_load_2d(d + c.koff[1] * 128, &tmY1, &full[stage], col, row_s[j][1]);
                        tma_load_2d(d + c.koff[2] * 128, &tmY2, &full[stage], col, row_s[j][2]);
                        tma_load_2d(d + c.koff[3] * 128, &tmY3, &full[stage], col, row_s[j][3]);
                    }
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16(TILE_M, CM_BN, 0, 1);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t a_phase = 0;
        const uint32_t lbo = static_cast<uint32_t>(c.ktot) * 128u;
        const uint64_t da_base = make_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t db_base = make_smem_desc(smem_u32(sB), lbo, 1024);
        for (int t = blockIdx.x; t < n_work; t += gridDim.x) {
            if (!IMG && c.tile_info[t].x < 0) continue;
            mbar_wait(a_full, a_phase);
            a_phase ^= 1;
            tc_fence_after();
            for (int n = 0; n < c.n_pass; ++n) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                const uint32_t d_tmem = tmem_base + acc * CM_BN;
                for (int j = 0; j < n_src; ++j) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    cm_issue_mmas(d_tmem, smem_desc_advance(da_base, j * a_bytes), smem_desc_advance(db_base, stage * stage_bytes),
                                  idesc, nk, j != 0);
                    umma_commit(&empty[stage]);
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == CM_ACC) { acc = 0; acc_phase ^= 1; }
            }
            umma_commit(a_empty);       // every MMA that reads this tile's coefficients has retired
        }
    } else if (warp >= 4 && warp < 4 + CM_COEF_WARPS) {
        // ===================== coefficient builders: thread = token m of the tile =====================
        const int m = (warp - 4) * 32 + lane;
        uint32_t a_phase = 0;
        // fixed positions of this token's non-zeros: k = m (scale 0) and koff[s] + j0 + 1, + 2 (coarse scales)
        uint32_t off0 = cm_a_off(m, c.koff[0] + m), offa[4], offb[4];
#pragma unroll
        for (int s = 1; s < 4; ++s) {
            const int r = a.ratio[s];
            const int j0 = cm_floor_div(2 * m + 1 - r, 2 * r);
            offa[s] = cm_a_off(m, c.koff[s] + j0 + 1);
            offb[s] = cm_a_off(m, c.koff[s] + j0 + 2);
        }
        for (int t = blockIdx.x; t < n_work; t += gridDim.x) {
            float v0[2] = {0.f, 0.f}, va[2][4] = {}, vb[2][4] = {};
            if constexpr (IMG) {
                const int b = t / c.tiles_per_img, t0 = (t - b * c.tiles_per_img) * TILE_M;
                const int p = t0 + m;
                if (p < a.P) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j >= n_src) break;
                        const int item = b * a.topk + j;
                        const int slot = a.inv_perm[item];
                        const float g = a.gate ? a.gate[item] : 1.0f;
                        const float4 bt = *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4);
                        const float bs[4] = {bt.x * g, bt.y * g, bt.z * g, bt.w * g};
                        v0[j] = bs[0];
#pragma unroll
                        for (int s = 1; s < 4; ++s) {
                            const int r = a.ratio[s];
                            const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                            const int qa = L.i0 - (t0 / r - 1), qb = L.i1 - (t0 / r - 1);
                            const int q0 = cm_floor_div(2 * m + 1 - r, 2 * r) + 1;
                            va[j][s] = bs[s] * ((qa == q0 ? 1.0f - L.lam : 0.f) + (qb == q0 ? L.lam : 0.f));
                            vb[j][s] = bs[s] * ((qa == q0 + 1 ? 1.0f - L.lam : 0.f) + (qb == q0 + 1 ? L.lam : 0.f));
                        }
                    }
                }
            } else {
                const int2 ti = c.tile_info[t];
                const int e = ti.x;
                if (e < 0) continue;
                if (m < ti.y) {
                    const int rel0 = c.region_row[0] + t * TILE_M - c.seg_start[e];
                    const int rel = rel0 + m;
                    const int j = rel / a.P, p = rel - j * a.P;
                    const int slot = c.offsets[e] + j;
                    const float g = a.gate ? a.gate[a.perm[slot]] : 1.0f;
                    const float4 bt = *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4);
                    const float bs[4] = {bt.x * g, bt.y * g, bt.z * g, bt.w * g};
                    v0[0] = bs[0];
#pragma unroll
                    for (int s = 1; s < 4; ++s) {
                        const int r = a.ratio[s];
                        const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                        const int fs = c.seg_start[s * c.K + e] + rel0 / r - 1;
                        const int base = a.slot_row[s * a.n_items + slot];
                        const int qa = base + L.i0 - fs, qb = base + L.i1 - fs;
                        const int q0 = cm_floor_div(2 * m + 1 - r, 2 * r) + 1;
                        va[0][s] = bs[s] * ((qa == q0 ? 1.0f - L.lam : 0.f) + (qb == q0 ? L.lam : 0.f));
                        vb[0][s] = bs[s] * ((qa == q0 + 1 ? 1.0f - L.lam : 0.f) + (qb == q0 + 1 ? L.lam : 0.f));
                    }
                }
            }
            mbar_wait(a_empty, a_phase ^ 1);
            a_phase ^= 1;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j >= n_src) break;
                uint8_t* Aj = sA + j * a_bytes;
                *reinterpret_cast<__nv_bfloat16*>(Aj + off0) = __float2bfloat16_rn(v0[j]);
#pragma unroll
                for (int s = 1; s < 4; ++s) {
                    *reinterpret_cast<__nv_bfloat16*>(Aj + offa[s]) = __float2bfloat16_rn(va[j][s]);
                    *reinterpret_cast<__nv_bfloat16*>(Aj + offb[s]) = __float2bfloat16_rn(vb[j][s]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
    } else if (warp >= 4 + CM_COEF_WARPS) {
        // ===================== epilogue =====================
        const int q = warp & 3;                      // TMEM lane quarter (== warp index % 4)
        const int ew = warp - 4 - CM_COEF_WARPS;
        const int h = ew >> 2;                       // this warp takes the 32-column chunks h and h + 2 of every pass
        uint8_t* my_out = sOut + ew * 2 * EPI_SLOT_BYTES;
        int acc = 0; uint32_t acc_phase = 0;
        int oslot = 0;
        for (int t = blockIdx.x; t < n_work; t += gridDim.x) {
            const int m0 = q * 32;
            bool valid;                              // P % 32 == 0: a warp's 32 tokens are all valid or all padding
            long long orow = 0;
            int blk = 0, b = 0;
            if constexpr (IMG) {
                b = t / c.tiles_per_img;
                const int p0 = (t - b * c.tiles_per_img) * TILE_M + m0;
                valid = p0 < a.P;
                orow = static_cast<long long>(b) * a.P + p0;
                blk = p0 >> 5;
            } else {
                const int2 ti = c.tile_info[t];
                const int e = ti.x;
                if (e < 0) continue;
                valid = m0 < ti.y;
                if (valid) {
                    const int rel = c.region_row[0] + t * TILE_M - c.seg_start[e] + m0;
                    const int j = rel / a.P, p0 = rel - j * a.P;
                    b = a.perm[c.offsets[e] + j] / a.topk;
                    orow = static_cast<long long>(b) * a.P + p0;
                    blk = p0 >> 5;
                }
            }
            for (int n = 0; n < c.n_pass; ++n) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                if (valid) {
                    const uint32_t t_row = tmem_base + acc * CM_BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                    for (int cc = h; cc < CM_BN / 32; cc += 2) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_row + cc * 32, v);
                        tmem_ld_wait();
                        const int col0 = n * CM_BN + cc * 32;
                        float f[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
                        if constexpr (OUT_F32) {
                            float4* op = reinterpret_cast<float4*>(static_cast<float*>(a.out) + (orow + lane) * c.D + col0);
#pragma unroll
                            for (int k = 0; k < 8; ++k) op[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
                        } else {
                            if (lane == 0) tma_store_wait_read<1>();
                            __syncwarp();
                            uint8_t* so = my_out + oslot * EPI_SLOT_BYTES;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                uint4 u;
                                u.x = pack_bf16x2(f[8 * k + 0], f[8 * k + 1]);
                                u.y = pack_bf16x2(f[8 * k + 2], f[8 * k + 3]);
                                u.z = pack_bf16x2(f[8 * k + 4], f[8 * k + 5]);
                                u.w = pack_bf16x2(f[8 * k + 6], f[8 * k + 7]);
                                *reinterpret_cast<uint4*>(so + epi_slot_off(lane, k)) = u;
                            }
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmOut, so, col0, static_cast<int>(orow));
                                tma_store_commit();
                            }
                            oslot ^= 1;
                        }
                        // deterministic partial of global_feat: column sums over this warp's 32 tokens
                        const float cs = warp_colsum32(f, lane);
                        a.gpart[(static_cast<size_t>(b) * a.nblk + blk) * c.D + col0 + lane] = cs;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == CM_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
        if (!OUT_F32 && lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static inline bool cm_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// fills the K layout; false when the tensor-core path does not apply
static inline bool cm_geometry(const CombineArgs& a, int D, CmArgs& c) {
    if (a.topk < 1 || a.topk > 2 || a.Ps[0] != a.P || a.P % 32 != 0 || D % CM_BN != 0) return false;
    int off = 0;
    for (int s = 0; s < 4; ++s) {
        if (a.Ps[s] <= 0 || a.P % a.Ps[s] != 0) return false;
        const int r = a.P / a.Ps[s];
        if (!cm_pow2(r) || r > TILE_M || (s > 0 && r < 2)) return false;
        c.cap[s] = (s == 0) ? TILE_M : ((TILE_M / r + 2 + 7) / 8) * 8;
        c.koff[s] = off;
        off += c.cap[s];
    }
    c.ktot = off;        // every staged row comes from a TMA box: no uninitialised k rows
    return off % 16 == 0 && off <= 256;
}

static inline size_t cm_smem_bytes(const CmArgs& c, bool out_f32) {
    return static_cast<size_t>(c.n_src) * ((c.ktot + 63) / 64) * 16384 + static_cast<size_t>(c.stages) * c.ktot * 256 +
           (out_f32 ? 0 : CM_EPI_WARPS * 2 * EPI_SLOT_BYTES) + (2 * CM_STAGES + 2 * CM_ACC + 2) * 8 + 16 + 1024;
}


// =======================================================================================
// Forward pass 1 on the tensor cores: logits over scales + softmax -> beta.
//
//   logit_s(p) = w2 . ReLU(interp(Z_s)(p)) + b2,   beta = softmax_s(logit)      (reference swin.py:63-68, with the
//   first Linear already applied at native resolution: Z = Y W1^T + b1)
//
// The lerp is a matrix product again: U_s[128 tokens, H] = L_s[128, cap_s] * Zrows_s[cap_s, H] with L_s the pure
// lerp-weight matrix (2 non-zeros per row, k / 128 values: exact in bf16; identity for the finest scale).  The scales
// must stay separate (the ReLU sits between the lerp and the dot with w2), so a tile runs 4 MMAs groups per column
// pass into separate TMEM accumulators and the epilogue reduces each accumulator row with max/fma against w2:
// 2 CUDA-core instructions per interpolated element instead of ~6 (unpack, 2 lerp FMAs, max, fma + shuffles).
//
// Roles (512 threads) as in cm_out_kernel.  Stages alternate {finest-scale rows} / {coarse rows of the 3 other scales}
// of one column pass.  Requirements: those of cm_geometry plus H = D / 2 a multiple of 128 or 192.
// =======================================================================================
constexpr int CL_STAGES = 4;             // maximum; 48 KB stages (HB = 192) run 3
constexpr int CL_MAX_ACC = 4;          // TMEM accumulator ring: 512 / HB slots of HB columns

struct ClArgs {
    int n_tiles;
    const int2* tile_info;
    int region_row0;
    const int* seg_start;
    const int* offsets;
    int K;
    int H, HB, n_half, n_acc;    // hidden width, columns per pass (128 or 192), passes per scale, accumulator slots
    int stages;
    int cap[4], koff[4], kc;     // coarse scales 1..3: rows staged (multiple of 16), first k, total
};

MM_DEVINL float4 cl_softmax4(float l0, float l1, float l2, float l3) {
    const float mx = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx), e3 = expf(l3 - mx);
    const float inv = 1.0f / (e0 + e1 + e2 + e3);
    return make_float4(e0 * inv, e1 * inv, e2 * inv, e3 * inv);
}

__global__ void __launch_bounds__(CM_THREADS, 1)
cm_logits_kernel(const __grid_constant__ CUtensorMap tmZ0, const __grid_constant__ CUtensorMap tmZ1,
                 const __grid_constant__ CUtensorMap tmZ2, const __grid_constant__ CUtensorMap tmZ3,
                 const CombineArgs a, const ClArgs c) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = c.HB * 256;                          // 128 rows x HB columns of bf16
    const int ac_bytes = ((c.kc + 63) / 64) * 16384;
    uint8_t* sAi = smem;                                         // identity, 128 x 128 (two 64-wide k blocks)
    uint8_t* sAc = sAi + 32768;                                  // coarse lerp weights, 128 x kc
    uint8_t* sB = sAc + ac_bytes;
    float* s_w2 = reinterpret_cast<float*>(sB + c.stages * stage_bytes);          // [H] of the current expert
    float4* s_x = reinterpret_cast<float4*>(s_w2 + c.H);                           // [2 tile parities][128 tokens]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_x + 2 * TILE_M);
    uint64_t* empty = full + CL_STAGES;
    uint64_t* tfull = empty + CL_STAGES;
    uint64_t* tempty = tfull + CL_MAX_ACC;
    uint64_t* a_full = tempty + CL_MAX_ACC;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tma_prefetch_desc(&tmZ0); tma_prefetch_desc(&tmZ1); tma_prefetch_desc(&tmZ2); tma_prefetch_desc(&tmZ3); }
    if (threadIdx.x == 32) {
        for (int s = 0; s < CL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < CL_MAX_ACC; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], CM_EPI_WARPS); }
        mbar_init(a_full, CM_COEF_WARPS);
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = threadIdx.x * 16; i < 32768 + ac_bytes; i += CM_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < TILE_M) *reinterpret_cast<__nv_bfloat16*>(sAi + cm_a_off(threadIdx.x, threadIdx.x)) = __float2bfloat16_rn(1.0f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_chunk64 = c.HB / 64;

    if (warp == 0 && elect_one()) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int e = c.tile_info[t].x;
            if (e < 0) continue;
            const int row0 = c.region_row0 + t * TILE_M;
            const int rel0 = row0 - c.seg_start[e];
            int row_s[4];
            row_s[0] = row0;
#pragma unroll
            for (int s = 1; s < 4; ++s) row_s[s] = c.seg_start[s * c.K + e] + rel0 / a.ratio[s] - 1;
            for (int h = 0; h < c.n_half; ++h) {
                // finest scale: 128 rows x HB columns
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], static_cast<uint32_t>(TILE_M * c.HB * 2));
                uint8_t* dst = sB + stage * stage_bytes;
                for (int ch = 0; ch < n_chunk64; ++ch)
                    tma_load_2d(dst + ch * (TILE_M * 128), &tmZ0, &full[stage], h * c.HB + ch * 64, row_s[0]);
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
                // coarse scales: kc rows x HB columns
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], static_cast<uint32_t>(c.kc * c.HB * 2));
                dst = sB + stage * stage_bytes;
                for (int ch = 0; ch < n_chunk64; ++ch) {
                    uint8_t* d = dst + ch * (c.kc * 128);
                    const int col = h * c.HB + ch * 64;
                    tma_load_2d(d + c.koff[1] * 128, &tmZ1, &full[stage], col, row_s[1]);
                    tma_load_2d(d + c.koff[2] * 128, &tmZ2, &full[stage], col, row_s[2]);
                    tma_load_2d(d + c.koff[3] * 128, &tmZ3, &full[stage], col, row_s[3]);
                }
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_bf16(TILE_M, c.HB, 0, 1);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t a_phase = 0;
        const uint64_t dai = make_smem_desc(smem_u32(sAi), 16, 1024);
        const uint64_t dac = make_smem_desc(smem_u32(sAc), 16, 1024);
        const uint64_t db0_base = make_smem_desc(smem_u32(sB), TILE_M * 128, 1024);                         // finest-scale stage: chunks 16 KB apart
        const uint64_t dbc_base = make_smem_desc(smem_u32(sB), static_cast<uint32_t>(c.kc) * 128u, 1024);    // coarse stage: chunks kc rows apart
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            if (c.tile_info[t].x < 0) continue;
            for (int h = 0; h < c.n_half; ++h) {
                // ---- finest scale: identity x Z_0 tile (the identity block is static: no need to wait for this tile's
                //      lerp weights yet, so their rebuild hides behind these MMAs) ----
                mbar_wait(&full[stage], phase);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                {
                    const uint32_t d_tmem = tmem_base + acc * c.HB;
                    cm_issue_mmas(d_tmem, dai, smem_desc_advance(db0_base, stage * stage_bytes), idesc, TILE_M / 16, false);
                    umma_commit(&empty[stage]);
                    umma_commit(&tfull[acc]);
                }
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
                if (++acc == c.n_acc) { acc = 0; acc_phase ^= 1; }
                // ---- coarse scales ----
                if (h == 0) {
                    mbar_wait(a_full, a_phase);
                    a_phase ^= 1;
                }
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t dbc = smem_desc_advance(dbc_base, stage * stage_bytes);
                for (int s = 1; s < 4; ++s) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * c.HB;
                    const int j0 = c.koff[s] >> 4, j1 = (c.koff[s] + c.cap[s]) >> 4;
                    for (int j = j0; j < j1; ++j)
                        umma_bf16(d_tmem, smem_desc_advance(dac, (j >> 2) * 16384 + (j & 3) * 32), smem_desc_advance(dbc, j * 2048), idesc,
                                  j != j0);
                    if (s == 3) umma_commit(&empty[stage]);
                    umma_commit(&tfull[acc]);
                    if (++acc == c.n_acc) { acc = 0; acc_phase ^= 1; }
                }
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(a_empty);
        }
    } else if (warp >= 4 && warp < 4 + CM_COEF_WARPS) {
        // ===================== lerp-weight builders: thread = token m of the tile =====================
        const int m = (warp - 4) * 32 + lane;
        uint32_t a_phase = 0;
        uint32_t offa[4], offb[4];
        int q0[4];
#pragma unroll
        for (int s = 1; s < 4; ++s) {
            const int r = a.ratio[s];
            q0[s] = cm_floor_div(2 * m + 1 - r, 2 * r) + 1;
            offa[s] = cm_a_off(m, c.koff[s] + q0[s]);
            offb[s] = cm_a_off(m, c.koff[s] + q0[s] + 1);
        }
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int2 ti = c.tile_info[t];
            const int e = ti.x;
            if (e < 0) continue;
            float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < ti.y) {
                const int rel0 = c.region_row0 + t * TILE_M - c.seg_start[e];
                const int rel = rel0 + m;
                const int j = rel / a.P, p = rel - j * a.P;
                const int slot = c.offsets[e] + j;
#pragma unroll
                for (int s = 1; s < 4; ++s) {
                    const int r = a.ratio[s];
                    const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                    const int fs = c.seg_start[s * c.K + e] + rel0 / r - 1;
                    const int base = a.slot_row[s * a.n_items + slot];
                    const int qa = base + L.i0 - fs, qb = base + L.i1 - fs;
                    va[s] = (qa == q0[s] ? 1.0f - L.lam : 0.f) + (qb == q0[s] ? L.lam : 0.f);
                    vb[s] = (qa == q0[s] + 1 ? 1.0f - L.lam : 0.f) + (qb == q0[s] + 1 ? L.lam : 0.f);
                }
            }
            mbar_wait(a_empty, a_phase ^ 1);
            a_phase ^= 1;
#pragma unroll
            for (int s = 1; s < 4; ++s) {
                *reinterpret_cast<__nv_bfloat16*>(sAc + offa[s]) = __float2bfloat16_rn(va[s]);
                *reinterpret_cast<__nv_bfloat16*>(sAc + offb[s]) = __float2bfloat16_rn(vb[s]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
    } else if (warp >= 4 + CM_COEF_WARPS) {
        // ===================== epilogue: row-wise w2 . ReLU(.) over the accumulator, then the softmax over scales =====================
        const int q = warp & 3;
        const int ew = warp - 4 - CM_COEF_WARPS;
        const int hsel = ew >> 2;                    // column chunks hsel, hsel + 2, ... of every pass
        const int n_ch = c.HB / 32;
        const int epi_tid = threadIdx.x - (4 + CM_COEF_WARPS) * 32;
        int acc = 0; uint32_t acc_phase = 0;
        int cur_e = -1;
        uint32_t parity = 0;
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int2 ti = c.tile_info[t];
            const int e = ti.x;
            if (e < 0) continue;
            if (e != cur_e) {        // (re)load this expert's w2: every epilogue warp sees the same tile sequence
                named_bar_sync(1, CM_EPI_WARPS * 32);
                for (int i = epi_tid; i < c.H; i += CM_EPI_WARPS * 32) s_w2[i] = a.w2[static_cast<size_t>(e) * c.H + i];
                named_bar_sync(1, CM_EPI_WARPS * 32);
                cur_e = e;
            }
            float lg[4] = {0.f, 0.f, 0.f, 0.f};
            for (int h = 0; h < c.n_half; ++h) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    mbar_wait(&tfull[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t t_row = tmem_base + acc * c.HB + (static_cast<uint32_t>(q * 32) << 16);
                    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;      // four independent FMA chains
#pragma unroll 1
                    for (int cc = hsel; cc < n_ch; cc += 2) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_row + cc * 32, v);
                        const float4* wp = reinterpret_cast<const float4*>(s_w2 + h * c.HB + cc * 32);
                        float4 w[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) w[k] = wp[k];
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            p0 = fmaf(fmaxf(__uint_as_float(v[4 * k + 0]), 0.f), w[k].x, p0);
                            p1 = fmaf(fmaxf(__uint_as_float(v[4 * k + 1]), 0.f), w[k].y, p1);
                            p2 = fmaf(fmaxf(__uint_as_float(v[4 * k + 2]), 0.f), w[k].z, p2);
                            p3 = fmaf(fmaxf(__uint_as_float(v[4 * k + 3]), 0.f), w[k].w, p3);
                        }
                    }
                    lg[s] += (p0 + p1) + (p2 + p3);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                    if (++acc == c.n_acc) { acc = 0; acc_phase ^= 1; }
                }
            }
            // the two warps of a lane quarter hold the two column halves of the same 32 tokens
            float4* xs = s_x + parity * TILE_M + q * 32 + lane;
            if (hsel == 1) *xs = make_float4(lg[0], lg[1], lg[2], lg[3]);
            named_bar_sync(1, CM_EPI_WARPS * 32);
            if (hsel == 0) {
                const int m = q * 32 + lane;
                if (m < ti.y) {
                    const float4 o = *xs;
                    const float b2 = a.b2[e];
                    const int rel = c.region_row0 + t * TILE_M - c.seg_start[e] + m;
                    const int j = rel / a.P, p = rel - j * a.P;
                    const int slot = c.offsets[e] + j;
                    *reinterpret_cast<float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4) =
                        cl_softmax4(lg[0] + o.x + b2, lg[1] + o.y + b2, lg[2] + o.z + b2, lg[3] + o.w + b2);
                }
            }
            parity ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// K layout of the coarse scales; false when the tensor-core logits kernel does not apply
static inline bool cl_geometry(const CombineArgs& a, int D, ClArgs& c) {
    CmArgs tmp{};
    CombineArgs one = a;
    one.topk = 1;                 // beta is per (image, choice) slot: the logits kernel does not care about top-k
    if (!cm_geometry(one, D, tmp)) return false;
    const int H = D / 2;
    c.H = H;
    if (H % 192 == 0) c.HB = 192;        // fewer, wider column passes win: the per-pass hand-off costs more than ring depth buys
    else if (H % 128 == 0) c.HB = 128;
    else return false;
    c.n_half = H / c.HB;
    c.n_acc = 512 / c.HB;
    c.stages = c.HB > 128 ? 3 : CL_STAGES;
    int off = 0;
    c.cap[0] = TILE_M; c.koff[0] = 0;
    for (int s = 1; s < 4; ++s) {
        const int r = a.P / a.Ps[s];
        c.cap[s] = ((TILE_M / r + 2 + 15) / 16) * 16;
        c.koff[s] = off;
        off += c.cap[s];
    }
    c.kc = off;
    return off <= TILE_M;       // the coarse rows share a stage sized for 128 rows
}

static inline size_t cl_smem_bytes(const ClArgs& c) {
    return 32768 + static_cast<size_t>((c.kc + 63) / 64) * 16384 + static_cast<size_t>(c.stages) * c.HB * 256 +
           static_cast<size_t>(c.H) * 4 + 2 * TILE_M * 16 + (2 * CL_STAGES + 2 * CL_MAX_ACC + 2) * 8 + 16 + 1024;
}

// =======================================================================================
// Backward pass A on the tensor cores (general path: local_feat has a cotangent).
//
//   dbeta_s(p) = <dF(p), interp(Y_s)(p)>                                  (autograd of swin.py:78-80)
//
// For a tile of 128 tokens this is a dense GEMM followed by a 7-entry gather:
//   G[128 tokens, 192 rows] = dlocal_tile[128, D] * Yrows[192, D]^T        (both operands K-major, K = D)
//   dbeta_s(p) = (1 - lam) G[p, row_a] + lam G[p, row_b]
// — the same staged row ranges as the forward, now as the N dimension.  The constant dglobal / P part of dF is not
// folded in here: it contributes interp(<dglobal, Y[row]> / P), which the rank-1 path already provides (row_dot).
// Roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
// Requirements: cm_geometry (any top-k) and P % 64 == 0 (a 64-token half tile belongs to one image); dlocal bf16.
// =======================================================================================
constexpr int CD_THREADS = 256;

struct CdArgs {
    int n_tiles;
    const int2* tile_info;
    int region_row0;
    const int* seg_start;
    const int* offsets;
    int K;
    int cap[4], koff[4], ktot;
    int D, n_kb, stages;
    float* dbeta_loc;            // [n_items, P, 4]
};

// v[idx] for a thread-dependent idx in [0, 32) without local memory
MM_DEVINL float cm_pick32(const uint32_t (&v)[32], int idx) {
    uint32_t r = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) r = (idx == j) ? v[j] : r;
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(CD_THREADS, 1)
cm_dbeta_kernel(const __grid_constant__ CUtensorMap tmDL, const __grid_constant__ CUtensorMap tmY0,
                const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
                const __grid_constant__ CUtensorMap tmY3, const CombineArgs a, const CdArgs c) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int A_BYTES = TILE_M * 128;                        // 128 tokens x 64 channels
    const int b_bytes = c.ktot * 128;                            // ktot rows x 64 channels
    const int stage_bytes = A_BYTES + b_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + c.stages * stage_bytes);
    uint64_t* empty = full + 8;
    uint64_t* tfull = empty + 8;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDL); tma_prefetch_desc(&tmY0); tma_prefetch_desc(&tmY1); tma_prefetch_desc(&tmY2); tma_prefetch_desc(&tmY3);
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < c.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && elect_one()) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int2 ti = c.tile_info[t];
            const int e = ti.x;
            if (e < 0) continue;
            const int row0 = c.region_row0 + t * TILE_M;
            const int rel0 = row0 - c.seg_start[e];
            int row_s[4], row_dl[2];
            row_s[0] = row0;
#pragma unroll
            for (int s = 1; s < 4; ++s) row_s[s] = c.seg_start[s * c.K + e] + rel0 / a.ratio[s] - 1;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {     // the two 64-token halves may belong to different images
                const int rel = rel0 + hh * 64;
                row_dl[hh] = 0;
                if (hh * 64 < ti.y) {
                    const int j = rel / a.P, p = rel - j * a.P;
                    row_dl[hh] = (a.perm[c.offsets[e] + j] / a.topk) * a.P + p;
                }
            }
            for (int kb = 0; kb < c.n_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], static_cast<uint32_t>(stage_bytes));
                uint8_t* dA = smem + stage * stage_bytes;
                uint8_t* dB = dA + A_BYTES;
                tma_load_2d(dA, &tmDL, &full[stage], kb * 64, row_dl[0]);
                tma_load_2d(dA + 8192, &tmDL, &full[stage], kb * 64, row_dl[1]);
                tma_load_2d(dB + c.koff[0] * 128, &tmY0, &full[stage], kb * 64, row_s[0]);
                tma_load_2d(dB + c.koff[1] * 128, &tmY1, &full[stage], kb * 64, row_s[1]);
                tma_load_2d(dB + c.koff[2] * 128, &tmY2, &full[stage], kb * 64, row_s[2]);
                tma_load_2d(dB + c.koff[3] * 128, &tmY3, &full[stage], kb * 64, row_s[3]);
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_bf16(TILE_M, c.ktot, 0, 0);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        const uint64_t da_base = make_smem_desc(smem_u32(smem), 16, 1024);
        const uint64_t db_base = make_smem_desc(smem_u32(smem) + A_BYTES, 16, 1024);
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            if (c.tile_info[t].x < 0) continue;
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kb = 0; kb < c.n_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t da = smem_desc_advance(da_base, stage * stage_bytes);
                const uint64_t db = smem_desc_advance(db_base, stage * stage_bytes);
                umma_bf16(d_tmem, da, db, idesc, kb != 0);
                umma_bf16(d_tmem, smem_desc_advance(da, 32), smem_desc_advance(db, 32), idesc, 1);
                umma_bf16(d_tmem, smem_desc_advance(da, 64), smem_desc_advance(db, 64), idesc, 1);
                umma_bf16(d_tmem, smem_desc_advance(da, 96), smem_desc_advance(db, 96), idesc, 1);
                umma_commit(&empty[stage]);
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: thread = token, gather its 7 entries of G =====================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int q0[4];
        q0[0] = m;
#pragma unroll
        for (int s = 1; s < 4; ++s) q0[s] = cm_floor_div(2 * m + 1 - a.ratio[s], 2 * a.ratio[s]) + 1;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int2 ti = c.tile_info[t];
            const int e = ti.x;
            if (e < 0) continue;
            // lerp weights of this token (positions q0[s], q0[s] + 1 of scale s), as in the forward
            float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
            size_t tok = 0;
            const bool valid = m < ti.y;
            if (valid) {
                const int rel0 = c.region_row0 + t * TILE_M - c.seg_start[e];
                const int rel = rel0 + m;
                const int j = rel / a.P, p = rel - j * a.P;
                const int slot = c.offsets[e] + j;
                tok = static_cast<size_t>(slot) * a.P + p;
#pragma unroll
                for (int s = 1; s < 4; ++s) {
                    const int r = a.ratio[s];
                    const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                    const int fs = c.seg_start[s * c.K + e] + rel0 / r - 1;
                    const int base = a.slot_row[s * a.n_items + slot];
                    const int qa = base + L.i0 - fs, qb = base + L.i1 - fs;
                    va[s] = (qa == q0[s] ? 1.0f - L.lam : 0.f) + (qb == q0[s] ? L.lam : 0.f);
                    vb[s] = (qa == q0[s] + 1 ? 1.0f - L.lam : 0.f) + (qb == q0[s] + 1 ? L.lam : 0.f);
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            float d[4];
            {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c.koff[0] + q * 32, v);      // diagonal block of the finest scale
                tmem_ld_wait();
                d[0] = cm_pick32(v, lane);
            }
#pragma unroll
            for (int s = 1; s < 4; ++s) {
                const int qbase = __shfl_sync(0xffffffffu, q0[s], 0);      // lane 0 has the warp's smallest position
                uint32_t v[32];
                tmem_ld_32x32(t_row + c.koff[s] + qbase, v);
                tmem_ld_wait();
                const int idx = q0[s] - qbase;
                d[s] = va[s] * cm_pick32(v, idx) + vb[s] * cm_pick32(v, idx + 1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (valid) *reinterpret_cast<float4*>(c.dbeta_loc + tok * 4) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// =======================================================================================
// Backward pass B on the tensor cores (general path): the local part of d fused / d Y,
//
//   dU_s[i, :] = sum_p w_i(p) g beta_s(p) dlocal(p, :)                     (transpose of the forward's C, autograd of swin.py:42,78-80)
//
// A tile owns 128 finest-scale rows and the 128 / r_s coarse rows below them.  The token window of an owned coarse row
// reaches r_s / 2 tokens outside the tile, so the tile stages the 192 tokens [R0 - 32, R0 + 160) of dlocal (MN-major B
// operand, tokens = K) and every owned row is complete: no atomics, no partial sums.  Two MMAs per 64-column pass:
//   D0[128, 64] = diag(g beta_0) * dlocal[R0 .. R0+128)          (A0: 128 x 128 diagonal, B rows 32..160)
//   D1[128, 64] = C1^T * dlocal[R0-32 .. R0+160)                  (A1: rows 0..n1+n2+n3 = the owned coarse rows, K = 192 tokens)
// The coefficient matrices are written by one thread per token at fixed positions (7 values per tile and token).
// Roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-9 coefficient builders,
// warps 10-17 epilogue.  Requirements: those of cm_dbeta_kernel (P % 64 == 0, ratios 2..64 powers of two).
// =======================================================================================
constexpr int CU_TOK = 192;                  // tokens staged per tile (halo 32 on both sides)
constexpr int CU_HALO = 32;
constexpr int CU_BN = 64;                    // columns per pass
constexpr int CU_STAGES = 4;
constexpr int CU_ACC = 4;                    // TMEM ring: 4 x (D0 64 + D1 64) columns
constexpr int CU_BUILD_WARPS = CU_TOK / 32;
constexpr int CU_EPI_WARPS = 8;
constexpr int CU_THREADS = (4 + CU_BUILD_WARPS + CU_EPI_WARPS) * 32;

struct CuArgs {
    int n_tiles;
    const int2* tile_info;
    int region_row0;
    const int* seg_start;
    const int* offsets;
    const int* counts;
    int K;
    int n_own[4], o_row[4];      // owned coarse rows per scale (128 / r) and their first row in D1
    int D, n_pass;
    __nv_bfloat16* dUT;
};

__global__ void __launch_bounds__(CU_THREADS, 1)
cm_dut_kernel(const __grid_constant__ CUtensorMap tmDL, const __grid_constant__ CUtensorMap tmOut, const CombineArgs a,
              const CuArgs c) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int A0_BYTES = 2 * 16384;                          // 128 x 128 (two 64-wide k blocks)
    constexpr int A1_BYTES = 3 * 16384;                          // 128 x 192
    constexpr int STAGE_BYTES = CU_TOK * 128;                    // 192 tokens x 64 columns
    uint8_t* sA0 = smem;
    uint8_t* sA1 = sA0 + A0_BYTES;
    uint8_t* sB = sA1 + A1_BYTES;
    uint8_t* sOut = sB + CU_STAGES * STAGE_BYTES;                // [CU_EPI_WARPS][2 slots]
    uint64_t* full = reinterpret_cast<uint64_t*>(sOut + CU_EPI_WARPS * 2 * EPI_SLOT_BYTES);
    uint64_t* empty = full + CU_STAGES;
    uint64_t* tfull = empty + CU_STAGES;
    uint64_t* tempty = tfull + CU_ACC;
    uint64_t* a_full = tempty + CU_ACC;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) { tma_prefetch_desc(&tmDL); tma_prefetch_desc(&tmOut); }
    if (threadIdx.x == 32) {
        for (int s = 0; s < CU_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < CU_ACC; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], CU_EPI_WARPS); }
        mbar_init(a_full, CU_BUILD_WARPS);
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = threadIdx.x * 16; i < A0_BYTES + A1_BYTES; i += CU_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && elect_one()) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int e = c.tile_info[t].x;
            if (e < 0) continue;
            const int rel0 = c.region_row0 + t * TILE_M - c.seg_start[e];
            const int seg_tokens = c.counts[e] * a.P;
            int row_g[CU_TOK / 32];                 // dlocal row of each 32-token group (a group never straddles images)
#pragma unroll
            for (int g = 0; g < CU_TOK / 32; ++g) {
                const int rel = rel0 - CU_HALO + 32 * g;
                row_g[g] = 0;
                if (rel >= 0 && rel < seg_tokens) {
                    const int j = rel / a.P, p = rel - j * a.P;
                    row_g[g] = (a.perm[c.offsets[e] + j] / a.topk) * a.P + p;
                }
            }
            for (int n = 0; n < c.n_pass; ++n) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], STAGE_BYTES);
                uint8_t* dst = sB + stage * STAGE_BYTES;
#pragma unroll
                for (int g = 0; g < CU_TOK / 32; ++g) tma_load_2d(dst + g * 4096, &tmDL, &full[stage], n * CU_BN, row_g[g]);
                if (++stage == CU_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16(TILE_M, CU_BN, 0, 1);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t a_phase = 0;
        const uint64_t da0 = make_smem_desc(smem_u32(sA0), 16, 1024), da1 = make_smem_desc(smem_u32(sA1), 16, 1024);
        const uint64_t db_base = make_smem_desc(smem_u32(sB), STAGE_BYTES, 1024);
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            if (c.tile_info[t].x < 0) continue;
            mbar_wait(a_full, a_phase);
            a_phase ^= 1;
            tc_fence_after();
            for (int n = 0; n < c.n_pass; ++n) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * 128, d1 = d0 + 64;
                const uint64_t db = smem_desc_advance(db_base, stage * STAGE_BYTES);
                // finest scale: tokens R0 .. R0+128 = staged rows 32 .. 160; owned coarse rows: all 192 staged tokens
                cm_issue_mmas(d0, da0, smem_desc_advance(db, CU_HALO * 128), idesc, TILE_M / 16, false);
                cm_issue_mmas(d1, da1, db, idesc, CU_TOK / 16, false);
                umma_commit(&empty[stage]);
                umma_commit(&tfull[acc]);
                if (++stage == CU_STAGES) { stage = 0; phase ^= 1; }
                if (++acc == CU_ACC) { acc = 0; acc_phase ^= 1; }
            }
            umma_commit(a_empty);
        }
    } else if (warp >= 4 && warp < 4 + CU_BUILD_WARPS) {
        // ===================== coefficient builders: thread = staged token u (tile-relative token u - 32) =====================
        const int u = (warp - 4) * 32 + lane;
        const int tr = u - CU_HALO;
        uint32_t a_phase = 0;
        const bool has0 = tr >= 0 && tr < TILE_M;
        const uint32_t off0 = has0 ? cm_a_off(tr, tr) : 0u;
        uint32_t offa[4], offb[4];
        bool ina[4], inb[4];
        int j0[4];
#pragma unroll
        for (int s = 1; s < 4; ++s) {
            const int r = a.ratio[s];
            j0[s] = cm_floor_div(2 * tr + 1 - r, 2 * r);       // unclamped lower lerp row, relative to the tile's first owned row
            ina[s] = j0[s] >= 0 && j0[s] < c.n_own[s];
            inb[s] = j0[s] + 1 >= 0 && j0[s] + 1 < c.n_own[s];
            offa[s] = ina[s] ? cm_a_off(c.o_row[s] + j0[s], u) : 0u;
            offb[s] = inb[s] ? cm_a_off(c.o_row[s] + j0[s] + 1, u) : 0u;
        }
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int e = c.tile_info[t].x;
            if (e < 0) continue;
            const int rel0 = c.region_row0 + t * TILE_M - c.seg_start[e];
            const int rel = rel0 + tr;
            float v0 = 0.f, va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
            if (rel >= 0 && rel < c.counts[e] * a.P) {
                const int j = rel / a.P, p = rel - j * a.P;
                const int slot = c.offsets[e] + j;
                const float g = a.gate ? a.gate[a.perm[slot]] : 1.0f;
                const float4 bt = *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4);
                const float bs[4] = {bt.x * g, bt.y * g, bt.z * g, bt.w * g};
                v0 = bs[0];
#pragma unroll
                for (int s = 1; s < 4; ++s) {
                    const int r = a.ratio[s];
                    const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                    const int own0 = c.seg_start[s * c.K + e] + rel0 / r;          // row-space row of the first owned row
                    const int base = a.slot_row[s * a.n_items + slot];
                    const int qa = base + L.i0 - own0, qb = base + L.i1 - own0;
                    va[s] = bs[s] * ((qa == j0[s] ? 1.0f - L.lam : 0.f) + (qb == j0[s] ? L.lam : 0.f));
                    vb[s] = bs[s] * ((qa == j0[s] + 1 ? 1.0f - L.lam : 0.f) + (qb == j0[s] + 1 ? L.lam : 0.f));
                }
            }
            mbar_wait(a_empty, a_phase ^ 1);
            a_phase ^= 1;
            if (has0) *reinterpret_cast<__nv_bfloat16*>(sA0 + off0) = __float2bfloat16_rn(v0);
#pragma unroll
            for (int s = 1; s < 4; ++s) {
                if (ina[s]) *reinterpret_cast<__nv_bfloat16*>(sA1 + offa[s]) = __float2bfloat16_rn(va[s]);
                if (inb[s]) *reinterpret_cast<__nv_bfloat16*>(sA1 + offb[s]) = __float2bfloat16_rn(vb[s]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
    } else if (warp >= 4 + CU_BUILD_WARPS) {
        // ===================== epilogue =====================
        const int q = warp & 3;                                  // TMEM lane quarter
        const int ew = warp - 4 - CU_BUILD_WARPS;
        const int hsel = ew >> 2;                                // 32-column chunk of the 64-column pass
        uint8_t* my_out = sOut + ew * 2 * EPI_SLOT_BYTES;
        int acc = 0; uint32_t acc_phase = 0;
        int oslot = 0;
        const int r_in = q * 32 + lane;                          // row of D0 / D1 this thread reads
        // which owned coarse row is D1 row r_in?
        int cs = 0, ci = 0;
        for (int s = 1; s < 4; ++s)
            if (r_in >= c.o_row[s] && r_in < c.o_row[s] + c.n_own[s]) { cs = s; ci = r_in - c.o_row[s]; }
        for (int t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
            const int e = c.tile_info[t].x;
            if (e < 0) continue;
            const int row0 = c.region_row0 + t * TILE_M;
            const int rel0 = row0 - c.seg_start[e];
            // coarse row this thread owns (if any), only inside the expert's 128-row padded segment of that region
            __nv_bfloat16* crow = nullptr;
            if (cs != 0) {
                const int rr = rel0 / a.ratio[cs] + ci;
                const int seg_rows = (c.counts[e] * a.Ps[cs] + TILE_M - 1) / TILE_M * TILE_M;
                if (rr < seg_rows) crow = c.dUT + (static_cast<long long>(c.seg_start[cs * c.K + e]) + rr) * c.D;
            }
            const bool any_coarse = __any_sync(0xffffffffu, crow != nullptr);
            for (int n = 0; n < c.n_pass; ++n) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + acc * 128 + (static_cast<uint32_t>(q * 32) << 16);
                const int col0 = n * CU_BN + hsel * 32;
                {   // D0: 32 finest-scale rows x 32 columns -> staging -> TMA store (padding rows come out as zeros)
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + hsel * 32, v);
                    tmem_ld_wait();
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    uint8_t* so = my_out + oslot * EPI_SLOT_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint4 w;
                        w.x = pack_bf16x2(__uint_as_float(v[8 * k + 0]), __uint_as_float(v[8 * k + 1]));
                        w.y = pack_bf16x2(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3]));
                        w.z = pack_bf16x2(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5]));
                        w.w = pack_bf16x2(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]));
                        *reinterpret_cast<uint4*>(so + epi_slot_off(lane, k)) = w;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, so, col0, row0 + q * 32);
                        tma_store_commit();
                    }
                    oslot ^= 1;
                }
                if (any_coarse) {   // D1: owned coarse rows, one row per thread, 64 bytes per pass chunk
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + 64 + hsel * 32, v);
                    tmem_ld_wait();
                    if (crow) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            stg_v4(crow + col0 + 8 * k,
                                   make_uint4(pack_bf16x2(__uint_as_float(v[8 * k + 0]), __uint_as_float(v[8 * k + 1])),
                                              pack_bf16x2(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3])),
                                              pack_bf16x2(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])),
                                              pack_bf16x2(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]))));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == CU_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

static inline size_t cu_smem_bytes() {
    return 2 * 16384 + 3 * 16384 + static_cast<size_t>(CU_STAGES) * CU_TOK * 128 + CU_EPI_WARPS * 2 * EPI_SLOT_BYTES +
           (2 * CU_STAGES + 2 * CU_ACC + 2) * 8 + 16 + 1024;
}

}  // namespace mm
