// medmoe_b200 — back-to-back expert GEMMs: the conv projection and the first attention Linear of one scale in ONE
// kernel, so that Y is written to HBM once and never read back by the GEMM that consumes it.
//
//   Y[m, :] = ReLU(f_s[m, :] Wp_e^T + bp_e)          reference swin.py:41  (Conv1d k=1 + ReLU)          "GEMM 1", K1 = D_s
//   Z[m, :] = Y[m, :] W1_e^T + b1_e                  reference swin.py:63  (first Linear of attn_proj,   "GEMM 2", K2 = D
//                                                    evaluated at native resolution)
//
// Per 128-row tile (one expert), D = 768 is cut into 12 chunks of 64 columns.  Chunk c of Y is a 128x64 accumulator in
// TMEM (two buffers); the epilogue warps turn it into bf16 and write it ONCE into shared memory in the K-major
// 128-byte-swizzled layout — which is at the same time (a) the A operand of GEMM 2's k block c and (b) the source box of
// the TMA store that puts Y into HBM for the combine / backward kernels.  Z accumulates over the 12 chunks in a
// 128x384 TMEM accumulator and leaves as six more 64-column blocks through the same shared-memory ring and store warp
// (stores straight from registers were measured: 4 x 32 partial-sector STG per block cost 0.27 ms per launch).
//
// TMEM: [0,128) two Y-chunk accumulators, [128,512) Z.   Shared memory (K1 <= 128): f tile 32 KB (resident for the
// tile), Wp chunk ring 2 x 16 KB, W1 half-chunk ring 4 x 24 KB, Y / Z chunk ring 3 x 16 KB = 208 KB.
//
// Roles: warp 0 TMA producer (two polled streams), warp 1 MMA issuer of GEMM 2, warp 3 MMA issuer of GEMM 1, warp 2 TMEM
// allocator + Y store (+ zero fill of tiles no expert owns), warps 4..19 epilogue (lane quarter = warp % 4; group g = (warp - 4) / 4: Y chunks of parity g / 2, column
// half g % 2; Z column blocks g, g + 4, g + 8).
// The two MMA issuers are independent instruction streams ordered only by mbarriers: GEMM 1 runs ahead of GEMM 2 as far as
// its two accumulators allow (they are released as soon as the epilogue has read them into registers), so the tensor
// pipe works on G2(c-1), G2(c-2) while the epilogue converts chunk c.
#pragma once
#include "gemm_pair.cuh"

#ifndef MM_B2B_DBG
#define MM_B2B_DBG 0      // tuning experiments ("switch parts off"): 1 no Y store, 2 no Z store, 4 no W1 loads, 8 no Wp loads, 16 no f loads,
                          // 32 epilogue 1 without math / smem write, 64 Z epilogue without math / smem write / store
#endif

namespace mm {

struct B2BFwdArgs {
    const int2* tile_info;   // [tile] {expert or -1, valid rows}
    int tile_begin;          // first entry of tile_info used by this launch (local tile 0 == row 0 of f / Y / Z)
    int tile_count;
    int K1;                  // D_s (multiple of 16, <= 64 * NKB1)
    const float* bias1;      // [E, D]
    const float* bias2;      // [E, H]
    __nv_bfloat16* y;        // [rows, ld_y]: only for the zero fill of unowned tiles
    long long ld_y;
    __nv_bfloat16* z;
    long long ld_z;
    const int* f_g64;        // pair kernel only, or nullptr: f is in IMAGE order and 64-row group g of the launch's (expert-sorted) row
                             // space lives at rows f_g64[g] .. + 64 of it (mm_dispatch_group_map; -1 = padding): no sorted copy of f
};

constexpr int B2B_D = 768, B2B_H = 384, B2B_NC = 64, B2B_NCH = B2B_D / B2B_NC;   // 12 chunks
constexpr int B2B_EPI_WARPS = 16;
constexpr int B2B_THREADS = (4 + B2B_EPI_WARPS) * 32;
#ifndef MM_B2B_S1
#define MM_B2B_S1 2
#endif
#ifndef MM_B2B_STG
#define MM_B2B_STG 0      // 1: the store warp copies the finished chunk smem -> global with LDS.128 + STG.128 (whole 128-byte lines);
#endif                    // 0: one TMA store per chunk (shares the TMA queue with the operand loads)
#ifndef MM_B2B_S2
#define MM_B2B_S2 4
#endif
constexpr int B2B_S1 = MM_B2B_S1, B2B_S2 = MM_B2B_S2, B2B_SA2 = 3;      // ring depths: Wp chunks, W1 half chunks, Y / Z chunks
constexpr int B2B_NZB = B2B_H / B2B_NC;              // Z leaves through the Y-chunk ring as 6 more 64-column blocks
constexpr int B2B_NRING = B2B_NCH + B2B_NZB;         // ring uses per tile
static_assert(B2B_NCH % B2B_S1 == 0 && (B2B_NCH / B2B_S1) % 2 == 0 && (2 * B2B_NCH) % B2B_S2 == 0 &&
                  ((2 * B2B_NCH) / B2B_S2) % 2 == 0 && (B2B_NCH / 2) % 2 == 0 && B2B_NRING % B2B_SA2 == 0 &&
                  (B2B_NRING / B2B_SA2) % 2 == 0,
              "ring stages and phases are compile-time functions of the chunk index: every ring must wrap an even number of times per tile");

template <int NKB1>
struct B2BSmem {
    static constexpr int A1_BYTES = NKB1 * 16384;                  // f tile: NKB1 k blocks of 128 rows x 64 bf16
    static constexpr int B1_STAGE = NKB1 * 8192;                   // Wp chunk: 64 rows x 64 k per k block
    static constexpr int B2_STAGE = (B2B_H / 2) * 128;             // W1 half chunk: 192 rows x 64 k = 24 KB
    static constexpr int A2_BYTES = 16384;                         // Y chunk: 128 rows x 64 bf16
    static constexpr int OFF_B1 = A1_BYTES;
    static constexpr int OFF_B2 = OFF_B1 + B2B_S1 * B1_STAGE;
    static constexpr int OFF_A2 = OFF_B2 + B2B_S2 * B2_STAGE;
    static constexpr int OFF_BAR = OFF_A2 + B2B_SA2 * A2_BYTES;
    static constexpr int N_BARS = 2 + 2 * B2B_S1 + 2 * B2B_S2 + 4 + 2 * B2B_SA2 + 2;
    static constexpr int TOTAL = OFF_BAR + N_BARS * 8 + 16 + 1024;
};

template <int NKB1>
__global__ void __launch_bounds__(B2B_THREADS, 1)
b2b_fwd_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmZ, const B2BFwdArgs a) {
    using S = B2BSmem<NKB1>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA1 = smem;
    uint8_t* sB1 = smem + S::OFF_B1;
    uint8_t* sB2 = smem + S::OFF_B2;
    uint8_t* sA2 = smem + S::OFF_A2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* a1full = bars;                   // 1 (TMA tx)
    uint64_t* a1empty = bars + 1;              // 1 (commit after the tile's last GEMM-1 MMA)
    uint64_t* b1full = bars + 2;               // [S1]
    uint64_t* b1empty = b1full + B2B_S1;
    uint64_t* b2full = b1empty + B2B_S1;       // [S2]
    uint64_t* b2empty = b2full + B2B_S2;
    uint64_t* acc1full = b2empty + B2B_S2;     // [2] commit
    uint64_t* acc1empty = acc1full + 2;        // [2] 8 epilogue warps
    uint64_t* a2full = acc1empty + 2;          // [SA2] 8 epilogue warps
    uint64_t* a2empty = a2full + B2B_SA2;      // [SA2] commit of G2(c) + the Y store warp
    uint64_t* acc2full = a2empty + B2B_SA2;    // commit
    uint64_t* acc2empty = acc2full + 1;        // 16 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmZ);
    }
    if (threadIdx.x == 32) {
        mbar_init(a1full, 1); mbar_init(a1empty, 1);
        for (int s = 0; s < B2B_S1; ++s) { mbar_init(&b1full[s], 1); mbar_init(&b1empty[s], 1); }
        for (int s = 0; s < B2B_S2; ++s) { mbar_init(&b2full[s], 1); mbar_init(&b2empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc1full[s], 1); mbar_init(&acc1empty[s], B2B_EPI_WARPS / 2); }
        for (int s = 0; s < B2B_SA2; ++s) { mbar_init(&a2full[s], B2B_EPI_WARPS / 2); mbar_init(&a2empty[s], 2); }
        mbar_init(acc2full, 1); mbar_init(acc2empty, B2B_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc2 = tmem_base + 2 * B2B_NC;

    if (warp == 0) {
        // ===================== TMA producer: one thread, two independent streams, polled without blocking =====================
        // stream 1 feeds GEMM 1 (the f tile, then the 12 Wp chunks of the tile), stream 2 feeds GEMM 2 (24 W1 half chunks);
        // each runs as far ahead as its own ring allows, whatever the state of the other.
        if (elect_one()) {
            auto next_owned = [&](int lt) {
                while (lt < a.tile_count && a.tile_info[a.tile_begin + lt].x < 0) lt += gridDim.x;
                return lt;
            };
            int lt1 = next_owned(blockIdx.x), c1 = -1;      // c1 == -1: the tile's f load is pending
            int lt2 = lt1, i2 = 0;
            int e1 = lt1 < a.tile_count ? a.tile_info[a.tile_begin + lt1].x : 0, e2 = e1;
            uint32_t a1ph = 0;
            uint32_t idle = 0;
            while (lt1 < a.tile_count || lt2 < a.tile_count) {
                bool progressed = false;
                if (lt1 < a.tile_count) {
                    if (c1 < 0) {
                        if (mbar_test_wait(a1empty, a1ph ^ 1)) {
                            a1ph ^= 1;
                            if (MM_B2B_DBG & 16) { mbar_arrive(a1full); } else {
                            mbar_expect_tx(a1full, NKB1 * 16384);
#pragma unroll
                            for (int kb = 0; kb < NKB1; ++kb) tma_load_2d(sA1 + kb * 16384, &tmA1, a1full, kb * 64, lt1 * TILE_M); }
                            c1 = 0; progressed = true;
                        }
                    } else {
                        const int b = c1 % B2B_S1;
                        if (mbar_test_wait(&b1empty[b], ((c1 / B2B_S1) & 1) ^ 1)) {
                            if (MM_B2B_DBG & 8) { mbar_arrive(&b1full[b]); } else {
                            mbar_expect_tx(&b1full[b], NKB1 * 8192);
#pragma unroll
                            for (int kb = 0; kb < NKB1; ++kb)
                                tma_load_2d(sB1 + b * S::B1_STAGE + kb * 8192, &tmB1, &b1full[b], kb * 64, e1 * B2B_D + c1 * B2B_NC); }
                            progressed = true;
                            if (++c1 == B2B_NCH) {
                                c1 = -1;
                                lt1 = next_owned(lt1 + gridDim.x);
                                if (lt1 < a.tile_count) e1 = a.tile_info[a.tile_begin + lt1].x;
                            }
                        }
                    }
                }
                if (lt2 < a.tile_count) {
                    const int s2 = i2 % B2B_S2;
                    if (mbar_test_wait(&b2empty[s2], ((i2 / B2B_S2) & 1) ^ 1)) {
                        if (MM_B2B_DBG & 4) { mbar_arrive(&b2full[s2]); } else {
                        mbar_expect_tx(&b2full[s2], S::B2_STAGE);
                        tma_load_2d(sB2 + s2 * S::B2_STAGE, &tmB2, &b2full[s2], (i2 >> 1) * B2B_NC, e2 * B2B_H + (i2 & 1) * (B2B_H / 2)); }
                        progressed = true;
                        if (++i2 == 2 * B2B_NCH) {
                            i2 = 0;
                            lt2 = next_owned(lt2 + gridDim.x);
                            if (lt2 < a.tile_count) e2 = a.tile_info[a.tile_begin + lt2].x;
                        }
                    }
                }
                if (progressed) idle = 0;
                else if (++idle > (1u << 27)) __trap();      // a protocol bug must become an error, never a hung GPU
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer, GEMM 2 (Z += Y_chunk W1_chunk^T) =====================
        // Two issuing threads (this one and warp 3 for GEMM 1): the instruction stream of a single elected thread
        // (waits, descriptor moves, commits: ~8 cycles per dependent uniform-datapath instruction) was the bound of the
        // first version.  Everything that depends on the chunk index is a compile-time constant: 12 chunks per tile fill
        // the 2-deep and 4-deep rings a whole (even) number of times, so stage and phase are functions of c alone.
        if (elect_one()) {
            constexpr uint32_t idesc2 = make_idesc_bf16(TILE_M, B2B_H / 2, 0, 0);
            const uint64_t db2 = make_smem_desc(smem_u32(sB2), 16, 1024);
            const uint64_t da2 = make_smem_desc(smem_u32(sA2), 16, 1024);
            uint32_t acc2e_ph = 0;
            for (int lt = blockIdx.x; lt < a.tile_count; lt += gridDim.x) {
                if (a.tile_info[a.tile_begin + lt].x < 0) continue;
                mbar_wait(acc2empty, acc2e_ph ^ 1); acc2e_ph ^= 1;
#pragma unroll
                for (int c = 0; c < B2B_NCH; ++c) {
                    const int b = c % B2B_SA2;
                    mbar_wait(&a2full[b], (c / B2B_SA2) & 1);
                    tc_fence_after();
                    const uint64_t da = smem_desc_advance(da2, b * S::A2_BYTES);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int idx = 2 * c + hf, s2 = idx % B2B_S2;
                        mbar_wait(&b2full[s2], (idx / B2B_S2) & 1);
                        tc_fence_after();
                        const uint64_t db = smem_desc_advance(db2, s2 * S::B2_STAGE);
                        const uint32_t d = tmem_acc2 + hf * (B2B_H / 2);
                        umma_bf16(d, da, db, idesc2, c != 0);
                        umma_bf16(d, smem_desc_advance(da, 32), smem_desc_advance(db, 32), idesc2, 1);
                        umma_bf16(d, smem_desc_advance(da, 64), smem_desc_advance(db, 64), idesc2, 1);
                        umma_bf16(d, smem_desc_advance(da, 96), smem_desc_advance(db, 96), idesc2, 1);
                        umma_commit(&b2empty[s2]);
                    }
                    umma_commit(&a2empty[b]);
                    if (c == B2B_NCH - 1) umma_commit(acc2full);
                }
            }
        }
    } else if (warp == 3) {
        // ===================== MMA issuer, GEMM 1 (acc1[c & 1] = f_tile Wp_chunk^T) =====================
        if (elect_one()) {
            constexpr uint32_t idesc1 = make_idesc_bf16(TILE_M, B2B_NC, 0, 0);
            const uint64_t da1 = make_smem_desc(smem_u32(sA1), 16, 1024);
            const uint64_t db1 = make_smem_desc(smem_u32(sB1), 16, 1024);
            const int ks_last = (a.K1 - (NKB1 - 1) * 64 + 15) / 16;      // k steps of the last k block (1..4)
            uint32_t a1ph = 0;
            for (int lt = blockIdx.x; lt < a.tile_count; lt += gridDim.x) {
                if (a.tile_info[a.tile_begin + lt].x < 0) continue;
                mbar_wait(a1full, a1ph); a1ph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < B2B_NCH; ++c) {
                    const int b = c & 1, sb = c % B2B_S1;
                    mbar_wait(&acc1empty[b], ((c >> 1) & 1) ^ 1);
                    mbar_wait(&b1full[sb], (c / B2B_S1) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + b * B2B_NC;
#pragma unroll
                    for (int kb = 0; kb < NKB1; ++kb) {
                        const uint64_t da = smem_desc_advance(da1, kb * 16384);
                        const uint64_t db = smem_desc_advance(db1, sb * S::B1_STAGE + kb * 8192);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (kb < NKB1 - 1 || k < ks_last)
                                umma_bf16(d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc1, (kb | k) != 0);
                    }
                    umma_commit(&b1empty[sb]);
                    umma_commit(&acc1full[b]);
                    if (c == B2B_NCH - 1) umma_commit(a1empty);
                }
            }
        }
    } else if (warp == 2) {
        // ===================== Y store: one 128 x 64 box per chunk, straight from the GEMM-2 operand buffer =====================
        for (int lt = blockIdx.x; lt < a.tile_count; lt += gridDim.x) {
            if (a.tile_info[a.tile_begin + lt].x < 0) {
                // tiles no expert owns get defined contents (consumers stage whole row ranges with TMA)
                __nv_bfloat16* yb = a.y + static_cast<long long>(lt) * TILE_M * a.ld_y;
                __nv_bfloat16* zb = a.z + static_cast<long long>(lt) * TILE_M * a.ld_z;
                for (int r = 0; r < TILE_M; ++r) {
                    for (int cidx = lane * 8; cidx < B2B_D; cidx += 256) stg_v4(yb + r * a.ld_y + cidx, make_uint4(0, 0, 0, 0));
                    for (int cidx = lane * 8; cidx < B2B_H; cidx += 256) stg_v4(zb + r * a.ld_z + cidx, make_uint4(0, 0, 0, 0));
                }
                continue;
            }
#pragma unroll
            for (int c = 0; c < B2B_NRING; ++c) {       // 12 Y chunks, then 6 Z blocks (no MMA reads those: arrive twice)
                const int b = c % B2B_SA2;
                mbar_wait(&a2full[b], (c / B2B_SA2) & 1);
#if MM_B2B_STG
                {
                    const uint8_t* src = sA2 + b * S::A2_BYTES;
                    const bool is_y = c < B2B_NCH;
                    const long long ld = is_y ? a.ld_y : a.ld_z;
                    __nv_bfloat16* dst = (is_y ? a.y + c * B2B_NC : a.z + (c - B2B_NCH) * B2B_NC) + static_cast<long long>(lt) * TILE_M * ld;
                    const int slot = lane & 7, rsub = lane >> 3;
                    if (!(MM_B2B_DBG & (is_y ? 1 : 2))) {
#pragma unroll 8
                        for (int it = 0; it < TILE_M / 4; ++it) {       // a warp instruction moves four whole 128-byte row segments
                            const int row = it * 4 + rsub;
                            const uint4 v = *reinterpret_cast<const uint4*>(src + row * 128 + ((slot ^ (row & 7)) << 4));
                            stg_v4(dst + row * ld + slot * 8, v);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&a2empty[b]); if (!is_y) mbar_arrive(&a2empty[b]); }
                }
#else
                if (lane == 0) {
                    if (c < B2B_NCH) {
                        if (!(MM_B2B_DBG & 1)) tma_store_2d(&tmY, sA2 + b * S::A2_BYTES, c * B2B_NC, lt * TILE_M);
                    } else {
                        if (!(MM_B2B_DBG & 2)) tma_store_2d(&tmZ, sA2 + b * S::A2_BYTES, (c - B2B_NCH) * B2B_NC, lt * TILE_M);
                    }
                    tma_store_commit();
                    // the buffer may be rewritten once this store has READ it; keep one younger store in flight
                    if (c > 0) {
                        tma_store_wait_read<1>();
                        mbar_arrive(&a2empty[(c - 1) % B2B_SA2]);
                        if (c - 1 >= B2B_NCH) mbar_arrive(&a2empty[(c - 1) % B2B_SA2]);
                    }
                    if (c == B2B_NRING - 1) { tma_store_wait_read<0>(); mbar_arrive(&a2empty[b]); mbar_arrive(&a2empty[b]); }
                }
                __syncwarp();
#endif
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;                 // TMEM lane quarter
        const int ew = warp - 4;
        const int g = ew >> 2;                  // 0..3
        const int p = g >> 1;                   // Y chunks of this parity
        const int hh = g & 1;                   // column half of the chunk
        uint32_t acc1f_ph = 0, acc2f_ph = 0;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const int r = q * 32 + lane;            // row inside the tile
        for (int lt = blockIdx.x; lt < a.tile_count; lt += gridDim.x) {
            const int2 ti = a.tile_info[a.tile_begin + lt];
            const int e = ti.x;
            if (e < 0) continue;
            const bool row_valid = r < ti.y;
            const float* b1p = a.bias1 + static_cast<size_t>(e) * B2B_D + hh * 32;
#pragma unroll 1
            for (int c = p; c < B2B_NCH; c += 2) {
                float4 bv[8];
                const float4* bp = reinterpret_cast<const float4*>(b1p + c * B2B_NC);
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
                mbar_wait(&acc1full[p], acc1f_ph); acc1f_ph ^= 1;
                tc_fence_after();
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_off + p * B2B_NC + hh * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1empty[p]);
                const int b2 = c % B2B_SA2;                      // Y-chunk buffer; its n-th use in the tile is c / SA2
                const uint32_t a2e_par = ((c / B2B_SA2) & 1) ^ 1;
                if (MM_B2B_DBG & 32) {
                    mbar_wait(&a2empty[b2], a2e_par);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a2full[b2]);
                    continue;
                }
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f0 = fmaxf(__uint_as_float(v[4 * j + 0]) + bv[j].x, 0.f);
                    const float f1 = fmaxf(__uint_as_float(v[4 * j + 1]) + bv[j].y, 0.f);
                    const float f2 = fmaxf(__uint_as_float(v[4 * j + 2]) + bv[j].z, 0.f);
                    const float f3 = fmaxf(__uint_as_float(v[4 * j + 3]) + bv[j].w, 0.f);
                    pk[2 * j] = row_valid ? pack_bf16x2(f0, f1) : 0u;
                    pk[2 * j + 1] = row_valid ? pack_bf16x2(f2, f3) : 0u;
                }
                mbar_wait(&a2empty[b2], a2e_par);
                uint8_t* a2row = sA2 + b2 * S::A2_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = hh * 4 + j;      // 16-byte chunk of the 128-byte row, 128B swizzle: chunk ^= row & 7
                    *reinterpret_cast<uint4*>(a2row + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2full[b2]);
            }
            // ---- Z: six 64-column blocks through the same ring; the warps of parity p take blocks p, p + 2, p + 4 ----
            mbar_wait(acc2full, acc2f_ph); acc2f_ph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int zc = p; zc < B2B_NZB; zc += 2) {
                const int col0 = zc * B2B_NC + hh * 32;
                float4 bv[8];
                const float4* bp = reinterpret_cast<const float4*>(a.bias2 + static_cast<size_t>(e) * B2B_H + col0);
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc2 + lane_off + col0, v);
                tmem_ld_wait();
                if (zc + 2 >= B2B_NZB) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc2empty);
                }
                const int c = B2B_NCH + zc;
                const int b2 = c % B2B_SA2;
                const uint32_t a2e_par = ((c / B2B_SA2) & 1) ^ 1;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f0 = __uint_as_float(v[4 * j + 0]) + bv[j].x, f1 = __uint_as_float(v[4 * j + 1]) + bv[j].y;
                    const float f2 = __uint_as_float(v[4 * j + 2]) + bv[j].z, f3 = __uint_as_float(v[4 * j + 3]) + bv[j].w;
                    pk[2 * j] = row_valid ? pack_bf16x2(f0, f1) : 0u;
                    pk[2 * j + 1] = row_valid ? pack_bf16x2(f2, f3) : 0u;
                }
                mbar_wait(&a2empty[b2], a2e_par);
                if (!(MM_B2B_DBG & 64)) {
                    uint8_t* a2row = sA2 + b2 * S::A2_BYTES + r * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ch = hh * 4 + j;
                        *reinterpret_cast<uint4*>(a2row + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    }
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2full[b2]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// =====================================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster (the two SMs of a TPC) work on the tiles (2j, 2j + 1)
// of the region — always one expert's (SEG_ALIGN) — and SHARE the weight tiles: each CTA stages half of every Wp chunk
// (32 of 64 rows) and half of every W1 half chunk (96 of 192 rows), the leader issues one M = 256 MMA per k step that reads
// both halves, and each CTA keeps the accumulators of its own 128 rows in its own TMEM.  Per SM this halves the bytes that
// come out of L2 for the weights (the single-CTA kernel re-fetches 0.72 MB of weights per tile) and, with the same shared
// memory, doubles the latency the operand rings cover (6 Wp chunks and 3 W1 chunks in flight instead of 2 and 2).
//
// Barriers (same shared-memory offsets in both CTAs; "L." = the leader's instance, the only one used):
//   L.a1full, L.b1full[], L.b2full[]   TMA bytes of BOTH CTAs (the leader's producer expects 2x, gemm_pair.cuh)
//   a1empty, b1empty[], b2empty[]      multicast commits -> each CTA's producer
//   acc1full[2], acc2full              multicast commits -> each CTA's epilogue
//   L.acc1empty[2] (2 x 8 warps), L.acc2empty (2 x 16 warps), L.a2pair[] (2 x 8 warps: a Y chunk is in shared memory in both
//                                      CTAs)   epilogue warps arrive locally (leader) or remotely (peer, default .release.cta form)
//   a2full[] (8 warps)                 own epilogue -> own store warp;   a2empty[] (2)   multicast commit of GEMM 2 + own store warp
// A tile no expert owns (second tile of a segment's last pair) runs the protocol with zero valid rows: its epilogue writes zeros.
// =====================================================================================================================
constexpr int B2BP_S2 = 6, B2BP_SA2 = 3;
constexpr int b2bp_s1(int nkb1) { return nkb1 <= 2 ? 6 : 2; }       // Wp ring depth: a 48 KB f tile (K1 up to 192) leaves room for two chunks
static_assert((2 * B2B_NCH) % B2BP_S2 == 0 && ((2 * B2B_NCH) / B2BP_S2) % 2 == 0 && B2B_NRING % B2BP_SA2 == 0 &&
                  (B2B_NRING / B2BP_SA2) % 2 == 0 && (B2B_NCH / B2BP_SA2) % 2 == 0 && (B2B_NCH / 6) % 2 == 0 && (B2B_NCH / 2) % 2 == 0,
              "ring stages and phases are compile-time functions of the chunk index: every ring must wrap an even number of times per tile");

template <int NKB1>
struct B2BPairSmem {
    static constexpr int A1_BYTES = NKB1 * 16384;                  // own f tile
    static constexpr int B1_STAGE = NKB1 * 4096;                   // half of a Wp chunk: 32 rows x 64 k per k block
    static constexpr int B2_STAGE = (B2B_H / 4) * 128;             // half of a W1 half chunk: 96 rows x 64 k = 12 KB
    static constexpr int A2_BYTES = 16384;
    static constexpr int S1 = b2bp_s1(NKB1);
    static constexpr int OFF_B1 = A1_BYTES;
    static constexpr int OFF_B2 = OFF_B1 + S1 * B1_STAGE;
    static constexpr int OFF_A2 = OFF_B2 + B2BP_S2 * B2_STAGE;
    static constexpr int OFF_BAR = OFF_A2 + B2BP_SA2 * A2_BYTES;
    static constexpr int N_BARS = 2 + 2 * S1 + 2 * B2BP_S2 + 4 + 3 * B2BP_SA2 + 2;
    static constexpr int TOTAL = OFF_BAR + N_BARS * 8 + 16 + 1024;
};

template <int NKB1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(B2B_THREADS, 1)
b2b_pair_fwd_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmY,
                    const __grid_constant__ CUtensorMap tmZ, const B2BFwdArgs a) {
    using S = B2BPairSmem<NKB1>;
    constexpr int B2BP_S1 = S::S1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA1 = smem;
    uint8_t* sB1 = smem + S::OFF_B1;
    uint8_t* sB2 = smem + S::OFF_B2;
    uint8_t* sA2 = smem + S::OFF_A2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* a1full = bars;
    uint64_t* a1empty = bars + 1;
    uint64_t* b1full = bars + 2;               // [S1]
    uint64_t* b1empty = b1full + B2BP_S1;
    uint64_t* b2full = b1empty + B2BP_S1;      // [S2]
    uint64_t* b2empty = b2full + B2BP_S2;
    uint64_t* acc1full = b2empty + B2BP_S2;    // [2]
    uint64_t* acc1empty = acc1full + 2;        // [2]
    uint64_t* a2full = acc1empty + 2;          // [SA2] own epilogue -> own store warp
    uint64_t* a2pair = a2full + B2BP_SA2;      // [SA2] both CTAs' epilogues -> leader's GEMM-2 issuer
    uint64_t* a2empty = a2pair + B2BP_SA2;     // [SA2]
    uint64_t* acc2full = a2empty + B2BP_SA2;
    uint64_t* acc2empty = acc2full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
    const int n_pairs = (a.tile_count + 1) >> 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmZ);
    }
    if (threadIdx.x == 32) {
        mbar_init(a1full, 1); mbar_init(a1empty, 1);
        for (int s = 0; s < B2BP_S1; ++s) { mbar_init(&b1full[s], 1); mbar_init(&b1empty[s], 1); }
        for (int s = 0; s < B2BP_S2; ++s) { mbar_init(&b2full[s], 1); mbar_init(&b2empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc1full[s], 1); mbar_init(&acc1empty[s], B2B_EPI_WARPS); }      // 2 CTAs x 8 warps
        for (int s = 0; s < B2BP_SA2; ++s) {
            mbar_init(&a2full[s], B2B_EPI_WARPS / 2); mbar_init(&a2pair[s], B2B_EPI_WARPS); mbar_init(&a2empty[s], 2);
        }
        mbar_init(acc2full, 1); mbar_init(acc2empty, 2 * B2B_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    cluster_sync_all();                        // both CTAs' barriers exist before anyone signals across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc2 = tmem_base + 2 * B2B_NC;

    // expert of a pair = expert of its first owned tile; -1 when neither tile is owned
    auto pair_expert = [&](int pj) {
        const int t0 = 2 * pj, t1 = 2 * pj + 1;
        const int e0 = a.tile_info[a.tile_begin + t0].x;
        if (e0 >= 0) return e0;
        return t1 < a.tile_count ? a.tile_info[a.tile_begin + t1].x : -1;
    };
    auto next_owned = [&](int pj) {
        while (pj < n_pairs && pair_expert(pj) < 0) pj += pair_step;
        return pj;
    };
    // arrive on the LEADER's instance of a barrier
    auto arrive_leader = [&](uint64_t* bar) {
        if (rank == 0) mbar_arrive(bar);
        else mbar_arrive_cluster(mapa_u32(bar, 0));
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs: own f tile, own halves of the weight chunks) =====================
        if (elect_one()) {
            int p1 = next_owned(pair0), c1 = -1;
            int p2 = p1, i2 = 0;
            int e1 = p1 < n_pairs ? pair_expert(p1) : 0, e2 = e1;
            uint32_t a1ph = 0;
            uint32_t idle = 0;
            const uint32_t bar_a1 = mapa_u32(a1full, 0);
            while (p1 < n_pairs || p2 < n_pairs) {
                bool progressed = false;
                if (p1 < n_pairs) {
                    if (c1 < 0) {
                        if (mbar_test_wait(a1empty, a1ph ^ 1)) {
                            a1ph ^= 1;
                            if (rank == 0) mbar_expect_tx(a1full, 2 * NKB1 * 16384);
                            const int lt = min(2 * p1 + static_cast<int>(rank), a.tile_count - 1);
                            if (a.f_g64) {      // two 64-row boxes per k block, addressed through the group map (padding groups read group 0)
                                const int r0 = max(a.f_g64[2 * lt], 0), r1 = max(a.f_g64[2 * lt + 1], 0);
#pragma unroll
                                for (int kb = 0; kb < NKB1; ++kb) {
                                    tma_load_2d_pair(sA1 + kb * 16384, &tmA1, bar_a1, kb * 64, r0);
                                    tma_load_2d_pair(sA1 + kb * 16384 + 8192, &tmA1, bar_a1, kb * 64, r1);
                                }
                            } else {
#pragma unroll
                                for (int kb = 0; kb < NKB1; ++kb) tma_load_2d_pair(sA1 + kb * 16384, &tmA1, bar_a1, kb * 64, lt * TILE_M);
                            }
                            c1 = 0; progressed = true;
                        }
                    } else {
                        const int b = c1 % B2BP_S1;
                        if (mbar_test_wait(&b1empty[b], ((c1 / B2BP_S1) & 1) ^ 1)) {
                            if (rank == 0) mbar_expect_tx(&b1full[b], 2 * S::B1_STAGE);
                            const uint32_t bar = mapa_u32(&b1full[b], 0);
#pragma unroll
                            for (int kb = 0; kb < NKB1; ++kb)
                                tma_load_2d_pair(sB1 + b * S::B1_STAGE + kb * 4096, &tmB1, bar, kb * 64,
                                                 e1 * B2B_D + c1 * B2B_NC + static_cast<int>(rank) * (B2B_NC / 2));
                            progressed = true;
                            if (++c1 == B2B_NCH) {
                                c1 = -1;
                                p1 = next_owned(p1 + pair_step);
                                if (p1 < n_pairs) e1 = pair_expert(p1);
                            }
                        }
                    }
                }
                if (p2 < n_pairs) {
                    const int s2 = i2 % B2BP_S2;
                    if (mbar_test_wait(&b2empty[s2], ((i2 / B2BP_S2) & 1) ^ 1)) {
                        if (rank == 0) mbar_expect_tx(&b2full[s2], 2 * S::B2_STAGE);
                        tma_load_2d_pair(sB2 + s2 * S::B2_STAGE, &tmB2, mapa_u32(&b2full[s2], 0), (i2 >> 1) * B2B_NC,
                                         e2 * B2B_H + (i2 & 1) * (B2B_H / 2) + static_cast<int>(rank) * (B2B_H / 4));
                        progressed = true;
                        if (++i2 == 2 * B2B_NCH) {
                            i2 = 0;
                            p2 = next_owned(p2 + pair_step);
                            if (p2 < n_pairs) e2 = pair_expert(p2);
                        }
                    }
                }
                if (progressed) idle = 0;
                else if (++idle > (1u << 27)) __trap();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer, GEMM 2 (leader CTA only): Z += Y_chunk W1_chunk^T on both CTAs' rows =====================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc2 = make_idesc_bf16(2 * TILE_M, B2B_H / 2, 0, 0);
            const uint64_t db2 = make_smem_desc(smem_u32(sB2), 16, 1024);
            const uint64_t da2 = make_smem_desc(smem_u32(sA2), 16, 1024);
            uint32_t acc2e_ph = 0;
            for (int pj = pair0; pj < n_pairs; pj += pair_step) {
                if (pair_expert(pj) < 0) continue;
                mbar_wait(acc2empty, acc2e_ph ^ 1); acc2e_ph ^= 1;
#pragma unroll
                for (int c = 0; c < B2B_NCH; ++c) {
                    const int b = c % B2BP_SA2;
                    mbar_wait(&a2pair[b], (c / B2BP_SA2) & 1);
                    tc_fence_after();
                    const uint64_t da = smem_desc_advance(da2, b * S::A2_BYTES);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int idx = 2 * c + hf, s2 = idx % B2BP_S2;
                        mbar_wait(&b2full[s2], (idx / B2BP_S2) & 1);
                        tc_fence_after();
                        const uint64_t db = smem_desc_advance(db2, s2 * S::B2_STAGE);
                        const uint32_t d = tmem_acc2 + hf * (B2B_H / 2);
                        umma_bf16_pair(d, da, db, idesc2, c != 0);
                        umma_bf16_pair(d, smem_desc_advance(da, 32), smem_desc_advance(db, 32), idesc2, 1);
                        umma_bf16_pair(d, smem_desc_advance(da, 64), smem_desc_advance(db, 64), idesc2, 1);
                        umma_bf16_pair(d, smem_desc_advance(da, 96), smem_desc_advance(db, 96), idesc2, 1);
                        umma_commit_pair(&b2empty[s2]);
                    }
                    umma_commit_pair(&a2empty[b]);
                    if (c == B2B_NCH - 1) umma_commit_pair(acc2full);
                }
            }
        }
    } else if (warp == 3) {
        // ===================== MMA issuer, GEMM 1 (leader CTA only): acc1[c & 1] = f_tile Wp_chunk^T =====================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc1 = make_idesc_bf16(2 * TILE_M, B2B_NC, 0, 0);
            const uint64_t da1 = make_smem_desc(smem_u32(sA1), 16, 1024);
            const uint64_t db1 = make_smem_desc(smem_u32(sB1), 16, 1024);
            const int ks_last = (a.K1 - (NKB1 - 1) * 64 + 15) / 16;
            uint32_t a1ph = 0;
            for (int pj = pair0; pj < n_pairs; pj += pair_step) {
                if (pair_expert(pj) < 0) continue;
                mbar_wait(a1full, a1ph); a1ph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < B2B_NCH; ++c) {
                    const int b = c & 1, sb = c % B2BP_S1;
                    mbar_wait(&acc1empty[b], ((c >> 1) & 1) ^ 1);
                    mbar_wait(&b1full[sb], (c / B2BP_S1) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + b * B2B_NC;
#pragma unroll
                    for (int kb = 0; kb < NKB1; ++kb) {
                        const uint64_t da = smem_desc_advance(da1, kb * 16384);
                        const uint64_t db = smem_desc_advance(db1, sb * S::B1_STAGE + kb * 4096);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (kb < NKB1 - 1 || k < ks_last)
                                umma_bf16_pair(d, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc1, (kb | k) != 0);
                    }
                    umma_commit_pair(&b1empty[sb]);
                    umma_commit_pair(&acc1full[b]);
                    if (c == B2B_NCH - 1) umma_commit_pair(a1empty);
                }
            }
        }
    } else if (warp == 2) {
        // ===================== Y / Z store of the own tile (zero fill when the whole pair is unowned) =====================
        for (int pj = pair0; pj < n_pairs; pj += pair_step) {
            const int lt = 2 * pj + static_cast<int>(rank);
            if (pair_expert(pj) < 0) {
                if (lt < a.tile_count) {
                    __nv_bfloat16* yb = a.y + static_cast<long long>(lt) * TILE_M * a.ld_y;
                    __nv_bfloat16* zb = a.z + static_cast<long long>(lt) * TILE_M * a.ld_z;
                    for (int r = 0; r < TILE_M; ++r) {
                        for (int cidx = lane * 8; cidx < B2B_D; cidx += 256) stg_v4(yb + r * a.ld_y + cidx, make_uint4(0, 0, 0, 0));
                        for (int cidx = lane * 8; cidx < B2B_H; cidx += 256) stg_v4(zb + r * a.ld_z + cidx, make_uint4(0, 0, 0, 0));
                    }
                }
                continue;
            }
            const bool store = lt < a.tile_count;
#pragma unroll
            for (int c = 0; c < B2B_NRING; ++c) {
                const int b = c % B2BP_SA2;
                mbar_wait(&a2full[b], (c / B2BP_SA2) & 1);
                if (lane == 0) {
                    if (store) {
                        if (c < B2B_NCH) tma_store_2d(&tmY, sA2 + b * S::A2_BYTES, c * B2B_NC, lt * TILE_M);
                        else tma_store_2d(&tmZ, sA2 + b * S::A2_BYTES, (c - B2B_NCH) * B2B_NC, lt * TILE_M);
                    }
                    tma_store_commit();
                    if (c > 0) {
                        tma_store_wait_read<1>();
                        mbar_arrive(&a2empty[(c - 1) % B2BP_SA2]);
                        if (c - 1 >= B2B_NCH) mbar_arrive(&a2empty[(c - 1) % B2BP_SA2]);
                    }
                    if (c == B2B_NRING - 1) { tma_store_wait_read<0>(); mbar_arrive(&a2empty[b]); mbar_arrive(&a2empty[b]); }
                }
                __syncwarp();
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp >= 4) {
        // ===================== epilogue (each CTA: its own 128 rows) =====================
        const int q = warp & 3;
        const int ew = warp - 4;
        const int g = ew >> 2;
        const int p = g >> 1;
        const int hh = g & 1;
        uint32_t acc1f_ph = 0, acc2f_ph = 0;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const int r = q * 32 + lane;
        for (int pj = pair0; pj < n_pairs; pj += pair_step) {
            const int e = pair_expert(pj);
            if (e < 0) continue;
            const int lt = 2 * pj + static_cast<int>(rank);
            int valid = 0;
            if (lt < a.tile_count) {
                const int2 ti = a.tile_info[a.tile_begin + lt];
                valid = ti.x >= 0 ? ti.y : 0;
            }
            const bool row_valid = r < valid;
            const float* b1p = a.bias1 + static_cast<size_t>(e) * B2B_D + hh * 32;
#pragma unroll 1
            for (int c = p; c < B2B_NCH; c += 2) {
                float4 bv[8];
                const float4* bp = reinterpret_cast<const float4*>(b1p + c * B2B_NC);
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
                mbar_wait(&acc1full[p], acc1f_ph); acc1f_ph ^= 1;
                tc_fence_after();
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_off + p * B2B_NC + hh * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader(&acc1empty[p]);
                const int b2 = c % B2BP_SA2;
                const uint32_t a2e_par = ((c / B2BP_SA2) & 1) ^ 1;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f0 = fmaxf(__uint_as_float(v[4 * j + 0]) + bv[j].x, 0.f);
                    const float f1 = fmaxf(__uint_as_float(v[4 * j + 1]) + bv[j].y, 0.f);
                    const float f2 = fmaxf(__uint_as_float(v[4 * j + 2]) + bv[j].z, 0.f);
                    const float f3 = fmaxf(__uint_as_float(v[4 * j + 3]) + bv[j].w, 0.f);
                    pk[2 * j] = row_valid ? pack_bf16x2(f0, f1) : 0u;
                    pk[2 * j + 1] = row_valid ? pack_bf16x2(f2, f3) : 0u;
                }
                mbar_wait(&a2empty[b2], a2e_par);
                uint8_t* a2row = sA2 + b2 * S::A2_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = hh * 4 + j;
                    *reinterpret_cast<uint4*>(a2row + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&a2full[b2]); arrive_leader(&a2pair[b2]); }
            }
            mbar_wait(acc2full, acc2f_ph); acc2f_ph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int zc = p; zc < B2B_NZB; zc += 2) {
                const int col0 = zc * B2B_NC + hh * 32;
                float4 bv[8];
                const float4* bp = reinterpret_cast<const float4*>(a.bias2 + static_cast<size_t>(e) * B2B_H + col0);
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc2 + lane_off + col0, v);
                tmem_ld_wait();
                if (zc + 2 >= B2B_NZB) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_leader(acc2empty);
                }
                const int c = B2B_NCH + zc;
                const int b2 = c % B2BP_SA2;
                const uint32_t a2e_par = ((c / B2BP_SA2) & 1) ^ 1;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f0 = __uint_as_float(v[4 * j + 0]) + bv[j].x, f1 = __uint_as_float(v[4 * j + 1]) + bv[j].y;
                    const float f2 = __uint_as_float(v[4 * j + 2]) + bv[j].z, f3 = __uint_as_float(v[4 * j + 3]) + bv[j].w;
                    pk[2 * j] = row_valid ? pack_bf16x2(f0, f1) : 0u;
                    pk[2 * j + 1] = row_valid ? pack_bf16x2(f2, f3) : 0u;
                }
                mbar_wait(&a2empty[b2], a2e_par);
                uint8_t* a2row = sA2 + b2 * S::A2_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = hh * 4 + j;
                    *reinterpret_cast<uint4*>(a2row + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2full[b2]);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();                        // the peer may still read this CTA's shared memory / signal its barriers
    if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace mm
