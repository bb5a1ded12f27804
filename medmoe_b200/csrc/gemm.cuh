// medmoe_b200 — tcgen05 / TMEM / TMA grouped GEMM kernels (sm_100a).
//
// Two persistent, warp-specialised kernels cover every matrix product on the hot path
// (SURVEY §2.2 rows E1, E4 and their backward; north-star item 3):
//
//   gemm_rows_kernel   C[m, n] = epi( sum_k A[m, k] * W_e[n, k] )       "TN": both operands K-major
//       rows m are expert-sorted, 128-row tiles never straddle experts (tile_info says
//       which expert's weight slab a tile multiplies).  Used for
//         E1  Y_s = ReLU(f_s W_s^T + b_s)      (reference swin.py:41, Conv1d k=1 + ReLU)
//         E4  Z   = Y W1^T + b1                (reference swin.py:63, first Linear of attn_proj,
//                                               evaluated at native resolution — lerp commutes)
//         dY  = (dUT + c[row] dglobal[img] + dZ W1) * [Y > 0]   (backward of E4 + ReLU of E1; tensor and / or rank-1 aux)
//         df_s = dPre_s W_s                    (backward of E1 w.r.t. the Swin stage features)
//
//   gemm_wgrad_kernel  dW_e[i, j] = sum_m A[m, i] * B[m, j]             both operands MN-major
//       reduction over the rows of expert e (split into chunks, fp32 red.add into dW).
//
// Roles per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator, warp 3 = zero fill of
// the tiles no expert owns, warps 4.. = epilogue (TMEM lane quarter = warp_idx % 4).  gemm_rows takes its epilogue warp
// count as a template parameter (8, 12 or 16): the warps of a lane quarter interleave the 32-column chunks, so every
// scheduler has 2-4 epilogue warps to hide tcgen05.ld / LDS / TMA-store latency behind each other.
// Pipelines: smem full/empty ring (TMA <-> MMA) and a 2-deep TMEM accumulator ring
// (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogue data movement (bf16 output): every epilogue warp owns 32 rows.  Per 32-column chunk
// it reads the accumulator with tcgen05.ld (thread = row), optionally combines it with the
// `aux` / `gate` tiles that it TMA-loaded into its private, double-buffered staging slots,
// writes the bf16 result into a 64-byte-swizzled staging slot and hands that slot to a TMA
// store — so global memory only ever sees whole 64-byte row segments of 32 consecutive rows.
#pragma once
#include "mm_common.cuh"

namespace mm {

enum : int {
    EPI_RELU = 1,       // out = max(out, 0)
    EPI_ZERO_PAD = 2,   // kept for ABI compatibility: padding rows of an owned tile are always written as zeros
    EPI_CAP_SOFTMAX = 4,  // every 32-column chunk is one caption: out = exp(cap_temp * softmax over its first cap_len[chunk] columns)
    EPI_PAIR_OK = 8,      // the caller guarantees that tiles (2j, 2j + 1) of the launch never belong to two experts (SEG_ALIGN row
                          // layout of mm_dispatch_build): CTA pairs (tcgen05 cta_group::2) may share the weight tiles
};

struct RowsGemmArgs {
    const int2* tile_info;   // [tile] {expert or -1, valid_rows}; nullptr => single problem of M rows
    int tile_begin;          // first entry of tile_info used by this launch
    int tile_count;          // number of 128-row tiles this launch covers (local tile 0 == row 0 of A/out)
    int M;                   // rows (only when tile_info == nullptr)
    int N, K;
    int n_tiles;             // N / BN
    const float* bias;       // [E, N] fp32 or nullptr
    void* out;               // fp32 [rows, ld_out] (OUT_F32 only; the bf16 path stores through tmOut)
    long long ld_out;
    float* colsum;           // [E, N] fp32, += column sums of the written tile (bias gradients) or nullptr
    float out_scale;         // multiplies the accumulator before bias (1.0 for the MoE path)
    int flags;
    // AUX bit 1 ("rank-1 aux"): aux[row, col] = row_coef[row] * vecs[row_vec[row], col], never materialised
    const float* row_coef;   // [rows] fp32
    const int* row_vec;      // [rows] index of the row's vector
    const float* vecs;       // [n_vecs, ld_vecs] fp32
    long long ld_vecs;
    // EPI_CAP_SOFTMAX (word-patch attention scores, local_loss.cu): words per caption of each 32-column chunk, temperature
    const int* cap_len;
    float cap_temp;
    // single-CTA bf16 epilogue only, or nullptr: the output is in IMAGE order; the 32 x 32 boxes of 64-row group g of the launch's
    // row space go to rows out_g64[g] .. + 64 of it (mm_dispatch_group_map), padding groups (-1) and unowned tiles are not stored
    const int* out_g64;
};

struct WgradArgs {
    const int4* chunks;      // [chunk] {expert, first_tile (global), num_tiles, 0}
    int chunk_begin, chunk_count;
    int tile_base;           // global tile index of row 0 of A/B in this launch
    int N1, N2;              // dW is [E, N1, N2]; A is [rows, N1], B is [rows, N2]
    int n_i, n_j;            // tiles along N1 (128 each) and N2 (BN each)
    float* out;              // [E, N1, N2] fp32, accumulated with red.add
    float* colsum;           // COLSUM only: [E, N1] fp32, += sum over rows of A[row, i] (column sums of A on the tensor cores)
    const int* b_g64;        // or nullptr: B is in IMAGE order; 64-row group g of the launch's row space = rows b_g64[g] .. + 64 of B
                             // (mm_dispatch_group_map; -1 = padding, whose rows of A are zero)
};

constexpr int EPI_SLOT_BYTES = 32 * 64;   // 32 rows x 32 bf16 columns
constexpr int rows_threads(int epi_warps) { return (4 + epi_warps) * 32; }   // gemm_rows_kernel: 4 role warps + epilogue warps

// AUX (bit flags): 0 = plain epilogue; otherwise out = (acc + aux) * [gate > 0] with the gate tile TMA-prefetched and
//      bit 0: aux tile, TMA-prefetched next to the gate;  bit 1: rank-1 aux coef[row] * vec[idx[row]] (never materialised)
template <int BN, int STAGES, int AUX = 0, int EPI_WARPS = 8>
struct GemmSmem {
    static constexpr int A_BYTES = TILE_M * 64 * 2;                 // 16 KB: 128 rows x 64 bf16 (K-major) or 2 x (64 k-rows x 64 mn)
    static constexpr int B_BYTES = ((BN + 63) / 64) * 64 * 64 * 2;   // BN rounded up to 64-wide chunks
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_SLOTS = (AUX || EPI_WARPS > 8) ? 1 : 2;                // output staging slots per epilogue warp
    static constexpr int EPI_OUT_BYTES = EPI_WARPS * OUT_SLOTS * EPI_SLOT_BYTES;
    static constexpr int IN_PER = (AUX & 1) ? 2 : 1;                               // prefetched tiles per buffer: {aux, gate} or {gate}
    static constexpr int EPI_IN_BYTES = AUX ? EPI_WARPS * 2 * IN_PER * EPI_SLOT_BYTES : 0;   // double-buffered
    static constexpr int BAR_BYTES = (2 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_OUT_BYTES + EPI_IN_BYTES + BAR_BYTES + 1024;   // + alignment slack
    static constexpr int WGRAD_TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;      // the wgrad kernel has no staging slots
};

// 0xffff in every 16-bit half of `w` (two bf16) that is > 0
MM_DEVINL uint32_t bf16x2_pos_mask(uint32_t w) {
    return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w), __float2bfloat162_rn(0.0f));
}

// 16-byte chunk j (0..3) of row r inside a 32-row x 64-byte staging slot written/read by TMA with
// CU_TENSOR_MAP_SWIZZLE_64B (address bits [4,6) ^= bits [7,9)).
MM_DEVINL uint32_t epi_slot_off(int r, int j) { return static_cast<uint32_t>(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

// ------------------------------------------------------------------------------------
// gemm_rows_kernel
// ------------------------------------------------------------------------------------
template <int BN, int STAGES, bool OUT_F32, int AUX, int EPI_WARPS>
__global__ void __launch_bounds__(rows_threads(EPI_WARPS), 1)
gemm_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                 const __grid_constant__ CUtensorMap tmGate, const RowsGemmArgs a) {
    using S = GemmSmem<BN, STAGES, AUX, EPI_WARPS>;
    static_assert(EPI_WARPS % 4 == 0, "epilogue warps come in groups of four (one per TMEM lane quarter)");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32, 256]");
    static_assert(BN % 16 == 0, "UMMA N constraint for M = 128");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * S::A_BYTES;
    uint8_t* sOut = smem + STAGES * S::STAGE_BYTES;
    uint8_t* sIn = sOut + S::EPI_OUT_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sIn + S::EPI_IN_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* inbar = tempty + 2;          // [EPI_WARPS][2 slots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + 2 * EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (!OUT_F32) tma_prefetch_desc(&tmOut);
        if (AUX & 1) tma_prefetch_desc(&tmAux);
        if (AUX) tma_prefetch_desc(&tmGate);
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EPI_WARPS); }
        for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&inbar[s], 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_kb = (a.K + 63) / 64;
    const int total_work = a.tile_count * a.n_tiles;

    if (warp == 0) {
        // ===================== TMA producer (one elected lane) =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int lt = w / a.n_tiles, nt = w - lt * a.n_tiles;
                int e = 0;
                if (a.tile_info) { e = a.tile_info[a.tile_begin + lt].x; if (e < 0) continue; }
                const int row_a = lt * TILE_M;
                const int row_b = e * a.N + nt * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], S::A_BYTES + BN * 128);
                    tma_load_2d(sA + stage * S::A_BYTES, &tmA, &full[stage], kb * 64, row_a);
                    tma_load_2d(sB + stage * S::B_BYTES, &tmB, &full[stage], kb * 64, row_b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane; descriptors advance by constant increments) =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN, 0, 0);
            const uint64_t da_base = make_smem_desc(smem_u32(sA), 16, 1024);
            const uint64_t db_base = make_smem_desc(smem_u32(sB), 16, 1024);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int lt = w / a.n_tiles;
                if (a.tile_info && a.tile_info[a.tile_begin + lt].x < 0) continue;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = smem_desc_advance(da_base, stage * S::A_BYTES);
                    const uint64_t db = smem_desc_advance(db_base, stage * S::B_BYTES);
                    const int ksteps = min(4, (a.K - kb * 64 + 15) / 16);
                    if (ksteps == 4) {
                        umma_bf16(d_tmem, da, db, idesc, kb != 0);
                        umma_bf16(d_tmem, smem_desc_advance(da, 32), smem_desc_advance(db, 32), idesc, 1);
                        umma_bf16(d_tmem, smem_desc_advance(da, 64), smem_desc_advance(db, 64), idesc, 1);
                        umma_bf16(d_tmem, smem_desc_advance(da, 96), smem_desc_advance(db, 96), idesc, 1);
                    } else {
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(d_tmem, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ===================== tiles no expert owns =====================
        // The capacity slack behind a region's last segment gets defined contents (zeros), so that consumers which
        // stage whole row ranges with TMA (combine_mma.cuh) never multiply uninitialised memory by a zero coefficient.
        if (!OUT_F32 && a.tile_info && !a.out_g64) {
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int lt = w / a.n_tiles, nt = w - lt * a.n_tiles;
                if (a.tile_info[a.tile_begin + lt].x >= 0) continue;
                __nv_bfloat16* base = static_cast<__nv_bfloat16*>(a.out) + static_cast<long long>(lt) * TILE_M * a.ld_out + nt * BN;
                for (int r = 0; r < TILE_M; ++r)
                    for (int cidx = lane * 8; cidx < BN; cidx += 256)
                        stg_v4(base + r * a.ld_out + cidx, make_uint4(0, 0, 0, 0));
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;                 // TMEM lane quarter
        const int ew = warp - 4;                // epilogue warp index
        const int h = ew >> 2;                  // this warp takes chunks h, h + 2, ...
        constexpr int NCH = BN / 32;
        constexpr int CSTEP = EPI_WARPS / 4;
        uint8_t* my_out = sOut + ew * S::OUT_SLOTS * EPI_SLOT_BYTES;
        constexpr int IN_BUF = S::IN_PER * EPI_SLOT_BYTES;
        uint8_t* my_in = sIn + ew * 2 * IN_BUF;              // buffer b at b * IN_BUF: [aux tile,] gate tile
        uint64_t* my_bar = inbar + ew * 2;
        int acc = 0; uint32_t acc_phase = 0;
        int oslot = 0;                 // staging slot for the next output chunk
        int islot = 0; uint32_t iphase[2] = {0, 0};

        // first valid work item of this CTA (for the aux/gate prefetch chain)
        auto next_valid = [&](int w) {
            for (; w < total_work; w += gridDim.x) {
                if (!a.tile_info || a.tile_info[a.tile_begin + w / a.n_tiles].x >= 0) break;
            }
            return w;
        };
        auto issue_in = [&](int w, int c, int slot) {   // lane 0 only
            const int lt = w / a.n_tiles, nt = w - lt * a.n_tiles;
            const int row0 = lt * TILE_M + q * 32, col0 = nt * BN + c * 32;
            mbar_expect_tx(&my_bar[slot], ((AUX & 1) ? 2 : 1) * EPI_SLOT_BYTES);
            if (AUX & 1) tma_load_2d(my_in + slot * IN_BUF, &tmAux, &my_bar[slot], col0, row0);
            tma_load_2d(my_in + slot * IN_BUF + (S::IN_PER - 1) * EPI_SLOT_BYTES, &tmGate, &my_bar[slot], col0, row0);
        };

        int w = next_valid(blockIdx.x);
        if (AUX && w < total_work && lane == 0 && h < NCH) issue_in(w, h, 0);
        while (w < total_work) {
            const int lt = w / a.n_tiles, nt = w - lt * a.n_tiles;
            int e = 0, valid = TILE_M;
            if (a.tile_info) {
                const int2 ti = a.tile_info[a.tile_begin + lt];
                e = ti.x; valid = ti.y;
            } else {
                valid = min(TILE_M, a.M - lt * TILE_M);
            }
            const int w_next = next_valid(w + gridDim.x);
            const int r_in_tile = q * 32 + lane;
            const long long row = static_cast<long long>(lt) * TILE_M + r_in_tile;
            const bool row_valid = r_in_tile < valid;
            float r1_coef = 0.f;
            const float* r1_vec = a.vecs;
            if ((AUX & 2) && row_valid) {     // issued before the accumulator wait: the two dependent loads hide behind it
                r1_coef = __ldg(a.row_coef + row);
                r1_vec = a.vecs + static_cast<long long>(__ldg(a.row_vec + row)) * a.ld_vecs;
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = h; c < NCH; c += CSTEP) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c * 32, v);
                if (AUX) {   // prefetch this warp's next chunk's aux/gate into the other slot (its readers finished last iteration)
                    if (lane == 0) {
                        if (c + CSTEP < NCH) issue_in(w, c + CSTEP, islot ^ 1);
                        else if (w_next < total_work) issue_in(w_next, h, islot ^ 1);
                    }
                }
                const int col0 = nt * BN + c * 32;
                float4 xv[8];
                if (AUX & 2) {      // rank-1 aux vector: issued before the TMEM wait so that both latencies overlap
                    const float4* vp = reinterpret_cast<const float4*>(r1_vec + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[j] = __ldg(vp + j);
                }
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (a.out_scale != 1.0f) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] *= a.out_scale;
                }
                if (a.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(a.bias + static_cast<size_t>(e) * a.N + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = __ldg(bp + j);
                        f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                    }
                }
                uint32_t gmask[16];      // AUX: 0xffff per bf16 half whose gate (a ReLU output, >= 0) is > 0
                if (AUX) {
                    mbar_wait(&my_bar[islot], iphase[islot]);
                    iphase[islot] ^= 1;
                    const uint8_t* ax = my_in + islot * IN_BUF;
                    const uint8_t* gt = ax + (S::IN_PER - 1) * EPI_SLOT_BYTES;
                    if (AUX & 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            f[4 * j + 0] = fmaf(r1_coef, xv[j].x, f[4 * j + 0]); f[4 * j + 1] = fmaf(r1_coef, xv[j].y, f[4 * j + 1]);
                            f[4 * j + 2] = fmaf(r1_coef, xv[j].z, f[4 * j + 2]); f[4 * j + 3] = fmaf(r1_coef, xv[j].w, f[4 * j + 3]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 g = *reinterpret_cast<const uint4*>(gt + epi_slot_off(lane, j));
                        if (AUX & 1) {
                            const uint4 u = *reinterpret_cast<const uint4*>(ax + epi_slot_off(lane, j));
                            f[8 * j + 0] += bf16lo(u.x); f[8 * j + 1] += bf16hi(u.x);
                            f[8 * j + 2] += bf16lo(u.y); f[8 * j + 3] += bf16hi(u.y);
                            f[8 * j + 4] += bf16lo(u.z); f[8 * j + 5] += bf16hi(u.z);
                            f[8 * j + 6] += bf16lo(u.w); f[8 * j + 7] += bf16hi(u.w);
                        }
                        gmask[4 * j + 0] = bf16x2_pos_mask(g.x); gmask[4 * j + 1] = bf16x2_pos_mask(g.y);
                        gmask[4 * j + 2] = bf16x2_pos_mask(g.z); gmask[4 * j + 3] = bf16x2_pos_mask(g.w);
                    }
                    __syncwarp();       // every lane is done with slot `islot` before lane 0 refills it next iteration
                    islot ^= 1;
                }
                if (a.flags & EPI_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (!AUX && (a.flags & EPI_CAP_SOFTMAX)) {     // a lane holds one row's scores against the 32 word slots of a caption
                    const int len = min(__ldg(a.cap_len + (col0 >> 5)), 32);
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, j < len ? f[j] : -INFINITY);
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        f[j] = j < len ? __expf(f[j] - mx) : 0.f;
                        sum += f[j];
                    }
                    const float sc = len > 0 ? a.cap_temp / sum : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = j < len ? __expf(f[j] * sc) : 0.f;
                }
                if (valid < TILE_M && !row_valid) {      // only a segment's last tile has padding rows
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 0.f;
                }
                if constexpr (OUT_F32) {
                    if (row_valid) {
                        float4* op = reinterpret_cast<float4*>(static_cast<float*>(a.out) + row * a.ld_out + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    }
                } else {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                        if (AUX) pk[j] &= gmask[j];      // the ReLU gate, applied to the packed pair
                    }
                    // staging slot `oslot` is free once the store that last used it has read it
                    if (lane == 0) tma_store_wait_read<S::OUT_SLOTS - 1>();
                    __syncwarp();
                    uint8_t* so = my_out + oslot * EPI_SLOT_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(so + epi_slot_off(lane, j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        int orow = lt * TILE_M + q * 32;
                        if (a.out_g64) {
                            const int g = a.out_g64[2 * lt + (q >> 1)];
                            orow = g < 0 ? -1 : g + (q & 1) * 32;
                        }
                        if (orow >= 0) tma_store_2d(&tmOut, so, col0, orow);
                        tma_store_commit();
                    }
                    if (S::OUT_SLOTS == 2) oslot ^= 1;
                    if (AUX && a.colsum) {               // column sums of what was written (gated, bf16-rounded)
#pragma unroll
                        for (int j = 0; j < 16; ++j) { f[2 * j] = bf16lo(pk[j]); f[2 * j + 1] = bf16hi(pk[j]); }
                    }
                }
                if (a.colsum) {
                    const float cs = warp_colsum32(f, lane);
                    atomicAdd(a.colsum + static_cast<size_t>(e) * a.N + col0 + lane, cs);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            w = w_next;
        }
        if (!OUT_F32 && lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------
// gemm_wgrad_kernel: dW[e][i][j] += sum over rows m of chunk: A[m][i] * B[m][j]
// ------------------------------------------------------------------------------------
// COLSUM: the B tile gets one more 64-column chunk that holds the constant 1.0 (written once, never touched by TMA) and
// the j == 0 work items run their MMAs 16 columns wider, so accumulator column BN of row i is sum_m A[m][i]: the
// column sums of A (the conv-bias gradients, A = dPre) come out of the tensor cores for free instead of a
// 31-shuffle transpose-reduction per 32x32 chunk in the epilogue of the GEMM that produced A.
template <int BN, int STAGES, bool COLSUM = false>
__global__ void __launch_bounds__(256, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradArgs a) {
    using S = GemmSmem<BN, STAGES>;
    static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64 (MN-major SW128 chunks)");
    static_assert(!COLSUM || BN <= 192, "the ones chunk needs 16 accumulator columns behind BN");
    constexpr int NB_CHUNKS = BN / 64;
    constexpr int B_STAGE = (NB_CHUNKS + (COLSUM ? 1 : 0)) * 8192;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * S::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (S::A_BYTES + B_STAGE));
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    if (COLSUM) {   // bf16 1.0 in the extra chunk of every stage
        for (int st = 0; st < STAGES; ++st) {
            uint8_t* ones = sB + st * B_STAGE + NB_CHUNKS * 8192;
            for (int i = threadIdx.x * 16; i < 8192; i += 256 * 16)
                *reinterpret_cast<uint4*>(ones + i) = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int per_chunk = a.n_i * a.n_j;
    const int total_work = a.chunk_count * per_chunk;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int c = w / per_chunk, rem = w - c * per_chunk;
                const int it = rem / a.n_j, jt = rem - it * a.n_j;
                const int4 ch = a.chunks[a.chunk_begin + c];
                if (ch.z <= 0) continue;
                const int row0 = (ch.y - a.tile_base) * TILE_M;
                const int num_kb = ch.z * 2;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], S::A_BYTES + NB_CHUNKS * 8192);
                    uint8_t* dA = sA + stage * S::A_BYTES;
                    uint8_t* dB = sB + stage * B_STAGE;
                    const int r = row0 + kb * 64;
                    const int rb = a.b_g64 ? max(a.b_g64[r >> 6], 0) : r;
                    tma_load_2d(dA, &tmA, &full[stage], it * 128, r);
                    tma_load_2d(dA + 8192, &tmA, &full[stage], it * 128 + 64, r);
#pragma unroll
                    for (int j = 0; j < NB_CHUNKS; ++j)
                        tma_load_2d(dB + j * 8192, &tmB, &full[stage], jt * BN + j * 64, rb);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_plain = make_idesc_bf16(TILE_M, BN, 1, 1);
            constexpr uint32_t idesc_wide = make_idesc_bf16(TILE_M, BN + 16, 1, 1);
            // 16 k-rows per MMA = 2048 B; 64-element MN chunks are 8192 B apart (LBO), 8-row k groups 1024 B apart (SBO)
            const uint64_t da_base = make_smem_desc(smem_u32(sA), 8192, 1024);
            const uint64_t db_base = make_smem_desc(smem_u32(sB), 8192, 1024);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int c = w / per_chunk;
                const int4 ch = a.chunks[a.chunk_begin + c];
                if (ch.z <= 0) continue;
                const int num_kb = ch.z * 2;
                const uint32_t idesc = (COLSUM && (w - c * per_chunk) % a.n_j == 0) ? idesc_wide : idesc_plain;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = smem_desc_advance(da_base, stage * S::A_BYTES);
                    const uint64_t db = smem_desc_advance(db_base, stage * B_STAGE);
                    umma_bf16(d_tmem, da, db, idesc, kb != 0);
                    umma_bf16(d_tmem, smem_desc_advance(da, 2048), smem_desc_advance(db, 2048), idesc, 1);
                    umma_bf16(d_tmem, smem_desc_advance(da, 4096), smem_desc_advance(db, 4096), idesc, 1);
                    umma_bf16(d_tmem, smem_desc_advance(da, 6144), smem_desc_advance(db, 6144), idesc, 1);
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int c = w / per_chunk, rem = w - c * per_chunk;
            const int it = rem / a.n_j, jt = rem - it * a.n_j;
            const int4 ch = a.chunks[a.chunk_begin + c];
            if (ch.z <= 0) continue;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int i = it * 128 + q * 32 + lane;
            const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            float* orow = a.out + (static_cast<size_t>(ch.x) * a.N1 + i) * a.N2;
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + cc * 32, v);
                tmem_ld_wait();
                const int col0 = jt * BN + cc * 32;
                if (i < a.N1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = col0 + 4 * j;
                        if (col + 3 < a.N2) {
                            red_add_v4_f32(orow + col, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        } else {
                            for (int t = 0; t < 4; ++t)
                                if (col + t < a.N2) atomicAdd(orow + col + t, __uint_as_float(v[4 * j + t]));
                        }
                    }
                }
            }
            if (COLSUM && jt == 0) {      // accumulator column BN = sum over the chunk's rows of A[row, i]
                uint32_t v[32];
                tmem_ld_32x32(t_row + BN, v);
                tmem_ld_wait();
                if (i < a.N1) atomicAdd(a.colsum + static_cast<size_t>(ch.x) * a.N1 + i, __uint_as_float(v[0]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace mm
