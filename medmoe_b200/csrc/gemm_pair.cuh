// medmoe_b200 — row GEMM on CTA pairs (tcgen05.mma.cta_group::2): C[m, n] = epi(sum_k A[m, k] W_e[n, k]).
//
// Two CTAs of a cluster (the two SMs of a TPC) compute a 256 x BN tile: CTA r stages its own 128 rows of A and HALF of
// the W tile (BN / 2 rows), the leader CTA issues one M = 256 MMA per k step that reads both halves, and each CTA keeps
// the accumulator of its 128 rows in its own TMEM.  Per SM this halves the B bytes TMA writes into and the tensor core
// reads out of shared memory, halves the weight bytes that come out of L2 per 128-row tile, and lets the same shared
// memory hold more stages (the single-CTA kernels re-fetch their weight tile from L2 for every tile).
//
// Pairs are (tile 2j, tile 2j + 1) of the launch; both must multiply the same expert's weights, which the 256-row segment
// alignment of the row layout guarantees (plan.py).  A pair whose second tile is not owned runs with zero valid rows there
// (it writes zeros, like the single-CTA kernel's zero fill); a pair with no owned tile is zero-filled by warp 3.
// gemm_rows_pair_kernel: plain epilogue (bias, ReLU, bf16 TMA store) — the E1 / E4 / dX path; gemm_rows_pair_r1_kernel
// (below): rank-1 aux + gate epilogue — the dY GEMM.  Used when the caller passes EPI_PAIR_OK (api_core.cu).
// Barriers: both CTAs' TMA loads complete on the LEADER's `full` (which expects the bytes of the pair); `empty` and `tfull`
// are signalled in both CTAs by multicast commits; the peer's epilogue warps arrive remotely on the leader's `tempty`.
// Remote arrives use the default .release.cta form: a cluster-scope release costs a full memory fence per arrive (measured:
// 2.3x slower kernel when the per-stage hand-shake used it).
#pragma once
#include "gemm.cuh"

namespace mm {

MM_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
MM_DEVINL void cluster_sync_all() {
    __syncwarp();                             // .aligned: the whole warp must arrive together (role lanes rejoin here)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` in CTA `rank` of the cluster
MM_DEVINL uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
MM_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
MM_DEVINL void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
}
MM_DEVINL void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
MM_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile load of a CTA of a pair: data lands in this CTA's shared memory, the byte count is credited to the mbarrier at
// the given shared::cluster address (the leader's).
MM_DEVINL void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
MM_DEVINL void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at the same offset in both CTAs once all previously issued MMAs of this thread retire
MM_DEVINL void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
                 : "memory");
}

template <int BN, int STAGES, int EPI_WARPS>
struct PairSmem {
    static constexpr int A_BYTES = TILE_M * 128;                   // own 128 rows x 64 k
    static constexpr int B_BYTES = (BN / 2) * 128;                 // half of the W tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OUT_BYTES = EPI_WARPS * EPI_SLOT_BYTES;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_OUT_BYTES + BAR_BYTES + 1024;
};

template <int BN, int STAGES, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(rows_threads(EPI_WARPS), 1)
gemm_rows_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const RowsGemmArgs a) {
    using S = PairSmem<BN, STAGES, EPI_WARPS>;
    static_assert(BN % 64 == 0 && BN <= 256, "BN / 2 must be a multiple of 32 rows");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * S::A_BYTES;
    uint8_t* sOut = smem + STAGES * S::STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sOut + S::EPI_OUT_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;

    if (threadIdx.x == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmOut); }
    if (threadIdx.x == 32) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 2 * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    cluster_sync_all();                       // both CTAs' barriers are initialised before anyone signals across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_kb = (a.K + 63) / 64;
    const int n_pairs = (a.tile_count + 1) >> 1;
    const int total_work = n_pairs * a.n_tiles;

    // expert of a pair = expert of its first owned tile; -1 when neither tile is owned
    auto pair_expert = [&](int pt) {
        if (!a.tile_info) return 0;
        const int t0 = 2 * pt, t1 = 2 * pt + 1;
        const int e0 = a.tile_info[a.tile_begin + t0].x;
        if (e0 >= 0) return e0;
        return t1 < a.tile_count ? a.tile_info[a.tile_begin + t1].x : -1;
    };

    if (warp == 0 && elect_one()) {
        // ===================== TMA producer (both CTAs: own A rows, own half of W) =====================
        int stage = 0; uint32_t phase = 0;
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
            const int e = pair_expert(pt);
            if (e < 0) continue;
            const int row_a = (2 * pt + static_cast<int>(rank)) * TILE_M;
            const int row_b = e * a.N + nt * BN + static_cast<int>(rank) * (BN / 2);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                // both CTAs' loads are credited to the leader's barrier, which expects the bytes of the pair
                if (rank == 0) mbar_expect_tx(&full[stage], 2 * S::STAGE_BYTES);
                const uint32_t bar = mapa_u32(&full[stage], 0);
                tma_load_2d_pair(sA + stage * S::A_BYTES, &tmA, bar, kb * 64, row_a);
                tma_load_2d_pair(sB + stage * S::B_BYTES, &tmB, bar, kb * 64, row_b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank == 0 && elect_one()) {
        // ===================== MMA issuer (leader CTA only) =====================
        constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, BN, 0, 0);
        const uint64_t da_base = make_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t db_base = make_smem_desc(smem_u32(sB), 16, 1024);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            if (pair_expert(w / a.n_tiles) < 0) continue;
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t da = smem_desc_advance(da_base, stage * S::A_BYTES);
                const uint64_t db = smem_desc_advance(db_base, stage * S::B_BYTES);
                const int ksteps = min(4, (a.K - kb * 64 + 15) / 16);
                for (int k = 0; k < ksteps; ++k)
                    umma_bf16_pair(d_tmem, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (kb | k) != 0);
                umma_commit_pair(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 3) {
        // ===================== pairs no expert owns: zero fill (see gemm_rows_kernel) =====================
        if (a.tile_info) {
            for (int w = pair; w < total_work; w += n_pairs_grid) {
                const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
                if (pair_expert(pt) >= 0) continue;
                const int lt = 2 * pt + static_cast<int>(rank);
                if (lt >= a.tile_count) continue;
                __nv_bfloat16* base = static_cast<__nv_bfloat16*>(a.out) + static_cast<long long>(lt) * TILE_M * a.ld_out + nt * BN;
                for (int r = 0; r < TILE_M; ++r)
                    for (int cidx = lane * 8; cidx < BN; cidx += 256) stg_v4(base + r * a.ld_out + cidx, make_uint4(0, 0, 0, 0));
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (each CTA: its own 128 rows) =====================
        const int q = warp & 3;
        const int ew = warp - 4;
        const int h = ew >> 2;
        constexpr int NCH = BN / 32;
        constexpr int CSTEP = EPI_WARPS / 4;
        uint8_t* so = sOut + ew * EPI_SLOT_BYTES;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
            const int e = pair_expert(pt);
            if (e < 0) continue;
            const int lt = 2 * pt + static_cast<int>(rank);
            int valid = 0;
            if (lt < a.tile_count) {
                if (a.tile_info) {
                    const int2 ti = a.tile_info[a.tile_begin + lt];
                    valid = ti.x >= 0 ? ti.y : 0;
                } else {
                    valid = max(0, min(TILE_M, a.M - lt * TILE_M));
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int r_in_tile = q * 32 + lane;
            const bool row_valid = r_in_tile < valid;
            const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            if (lt < a.tile_count) {
#pragma unroll 1
                for (int c = h; c < NCH; c += CSTEP) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = nt * BN + c * 32;
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (a.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(a.bias + static_cast<size_t>(e) * a.N + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(bp + j);
                            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                        }
                    }
                    if (a.flags & EPI_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (!row_valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = 0.f;
                    }
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(so + epi_slot_off(lane, j)) =
                            make_uint4(pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                       pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, so, col0, lt * TILE_M + q * 32);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();                       // the peer may still read this CTA's shared memory / signal its barriers
    if (warp == 2) tmem_dealloc2(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// CTA-pair row GEMM with the rank-1 aux + gate epilogue (gemm_rows_kernel's AUX = 2): the dY GEMM of the backward,
//   out = (A W_e^T + row_coef[row] * vecs[row_vec[row], :]) * [gate > 0]
// Same pair plumbing as above; the epilogue is gemm_rows_kernel's (gate tiles TMA-prefetched one chunk ahead into two
// slots per warp, rank-1 vector loaded before the TMEM wait, packed gate mask, swizzled staging, TMA store), with TWO
// output staging slots per warp: the halved W stages leave the room, and the per-chunk wait for the previous TMA store
// to have read its slot was the longest link of the single-CTA kernel's per-warp chain.
// ------------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EPI_WARPS>
struct PairR1Smem {
    static constexpr int A_BYTES = TILE_M * 128;
    static constexpr int B_BYTES = (BN / 2) * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_SLOTS = 2;
    static constexpr int EPI_OUT_BYTES = EPI_WARPS * OUT_SLOTS * EPI_SLOT_BYTES;
    static constexpr int EPI_IN_BYTES = EPI_WARPS * 2 * EPI_SLOT_BYTES;          // gate tiles, double-buffered
    static constexpr int BAR_BYTES = (2 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_OUT_BYTES + EPI_IN_BYTES + BAR_BYTES + 1024;
};

template <int BN, int STAGES, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(rows_threads(EPI_WARPS), 1)
gemm_rows_pair_r1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmGate,
                         const RowsGemmArgs a) {
    using S = PairR1Smem<BN, STAGES, EPI_WARPS>;
    static_assert(BN % 64 == 0 && BN <= 256, "BN / 2 must be a multiple of 32 rows");
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
    uint64_t* inbar = tempty + 2;              // [EPI_WARPS][2 slots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + 2 * EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;

    if (threadIdx.x == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmGate); }
    if (threadIdx.x == 32) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 2 * EPI_WARPS); }
        for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&inbar[s], 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_kb = (a.K + 63) / 64;
    const int n_pairs = (a.tile_count + 1) >> 1;
    const int total_work = n_pairs * a.n_tiles;

    auto pair_expert = [&](int pt) {
        const int t0 = 2 * pt, t1 = 2 * pt + 1;
        const int e0 = a.tile_info[a.tile_begin + t0].x;
        if (e0 >= 0) return e0;
        return t1 < a.tile_count ? a.tile_info[a.tile_begin + t1].x : -1;
    };
    // valid rows of this CTA's tile of pair pt (0: not owned / past the end)
    auto own_valid = [&](int pt) {
        const int lt = 2 * pt + static_cast<int>(rank);
        if (lt >= a.tile_count) return 0;
        const int2 ti = a.tile_info[a.tile_begin + lt];
        return ti.x >= 0 ? ti.y : 0;
    };

    if (warp == 0 && elect_one()) {
        // ===================== TMA producer (both CTAs: own A rows, own half of W) =====================
        int stage = 0; uint32_t phase = 0;
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
            const int e = pair_expert(pt);
            if (e < 0) continue;
            const int row_a = min(2 * pt + static_cast<int>(rank), a.tile_count - 1) * TILE_M;
            const int row_b = e * a.N + nt * BN + static_cast<int>(rank) * (BN / 2);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (rank == 0) mbar_expect_tx(&full[stage], 2 * S::STAGE_BYTES);
                const uint32_t bar = mapa_u32(&full[stage], 0);
                tma_load_2d_pair(sA + stage * S::A_BYTES, &tmA, bar, kb * 64, row_a);
                tma_load_2d_pair(sB + stage * S::B_BYTES, &tmB, bar, kb * 64, row_b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank == 0 && elect_one()) {
        // ===================== MMA issuer (leader CTA only) =====================
        constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, BN, 0, 0);
        const uint64_t da_base = make_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t db_base = make_smem_desc(smem_u32(sB), 16, 1024);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            if (pair_expert(w / a.n_tiles) < 0) continue;
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
                    umma_bf16_pair(d_tmem, da, db, idesc, kb != 0);
                    umma_bf16_pair(d_tmem, smem_desc_advance(da, 32), smem_desc_advance(db, 32), idesc, 1);
                    umma_bf16_pair(d_tmem, smem_desc_advance(da, 64), smem_desc_advance(db, 64), idesc, 1);
                    umma_bf16_pair(d_tmem, smem_desc_advance(da, 96), smem_desc_advance(db, 96), idesc, 1);
                } else {
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_pair(d_tmem, smem_desc_advance(da, k * 32), smem_desc_advance(db, k * 32), idesc, (kb | k) != 0);
                }
                umma_commit_pair(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 3) {
        // ===================== own tiles that are not owned by an expert: zero fill =====================
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
            const int lt = 2 * pt + static_cast<int>(rank);
            if (lt >= a.tile_count || a.tile_info[a.tile_begin + lt].x >= 0) continue;
            __nv_bfloat16* base = static_cast<__nv_bfloat16*>(a.out) + static_cast<long long>(lt) * TILE_M * a.ld_out + nt * BN;
            for (int r = 0; r < TILE_M; ++r)
                for (int cidx = lane * 8; cidx < BN; cidx += 256) stg_v4(base + r * a.ld_out + cidx, make_uint4(0, 0, 0, 0));
        }
    } else if (warp >= 4) {
        // ===================== epilogue (each CTA: its own 128 rows) =====================
        const int q = warp & 3;
        const int ew = warp - 4;
        const int h = ew >> 2;
        constexpr int NCH = BN / 32;
        constexpr int CSTEP = EPI_WARPS / 4;
        uint8_t* my_out = sOut + ew * S::OUT_SLOTS * EPI_SLOT_BYTES;
        uint8_t* my_in = sIn + ew * 2 * EPI_SLOT_BYTES;
        uint64_t* my_bar = inbar + ew * 2;
        int acc = 0; uint32_t acc_phase = 0;
        int oslot = 0;
        int islot = 0; uint32_t iphase[2] = {0, 0};

        // next work item of this pair sequence in which this CTA's own tile is owned (gate prefetch chain)
        auto next_active = [&](int w) {
            for (; w < total_work; w += n_pairs_grid)
                if (own_valid(w / a.n_tiles) > 0) break;
            return w;
        };
        auto issue_in = [&](int w, int c, int slot) {   // lane 0 only
            const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
            const int row0 = (2 * pt + static_cast<int>(rank)) * TILE_M + q * 32, col0 = nt * BN + c * 32;
            mbar_expect_tx(&my_bar[slot], EPI_SLOT_BYTES);
            tma_load_2d(my_in + slot * EPI_SLOT_BYTES, &tmGate, &my_bar[slot], col0, row0);
        };

        int w_act = next_active(pair);            // the work item whose gate tiles are (being) prefetched
        if (w_act < total_work && lane == 0 && h < NCH) issue_in(w_act, h, 0);
        for (int w = pair; w < total_work; w += n_pairs_grid) {
            const int pt = w / a.n_tiles, nt = w - pt * a.n_tiles;
            if (pair_expert(pt) < 0) continue;
            const int lt = 2 * pt + static_cast<int>(rank);
            const int valid = own_valid(pt);
            const int r_in_tile = q * 32 + lane;
            const long long row = static_cast<long long>(lt) * TILE_M + r_in_tile;
            const bool row_valid = r_in_tile < valid;
            float r1_coef = 0.f;
            const float* r1_vec = a.vecs;
            if (row_valid) {
                r1_coef = __ldg(a.row_coef + row);
                r1_vec = a.vecs + static_cast<long long>(__ldg(a.row_vec + row)) * a.ld_vecs;
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (valid > 0) {          // w == w_act: this warp's first gate tile of the item is in flight in slot `islot`
                const int w_next = next_active(w + n_pairs_grid);
                const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                for (int c = h; c < NCH; c += CSTEP) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + c * 32, v);
                    if (lane == 0) {
                        if (c + CSTEP < NCH) issue_in(w, c + CSTEP, islot ^ 1);
                        else if (w_next < total_work) issue_in(w_next, h, islot ^ 1);
                    }
                    const int col0 = nt * BN + c * 32;
                    float4 xv[8];
                    const float4* vp = reinterpret_cast<const float4*>(r1_vec + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[j] = __ldg(vp + j);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[4 * j + 0] = fmaf(r1_coef, xv[j].x, __uint_as_float(v[4 * j + 0]));
                        f[4 * j + 1] = fmaf(r1_coef, xv[j].y, __uint_as_float(v[4 * j + 1]));
                        f[4 * j + 2] = fmaf(r1_coef, xv[j].z, __uint_as_float(v[4 * j + 2]));
                        f[4 * j + 3] = fmaf(r1_coef, xv[j].w, __uint_as_float(v[4 * j + 3]));
                    }
                    mbar_wait(&my_bar[islot], iphase[islot]);
                    iphase[islot] ^= 1;
                    const uint8_t* gt = my_in + islot * EPI_SLOT_BYTES;
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 g = *reinterpret_cast<const uint4*>(gt + epi_slot_off(lane, j));
                        pk[4 * j + 0] = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]) & bf16x2_pos_mask(g.x);
                        pk[4 * j + 1] = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]) & bf16x2_pos_mask(g.y);
                        pk[4 * j + 2] = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]) & bf16x2_pos_mask(g.z);
                        pk[4 * j + 3] = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]) & bf16x2_pos_mask(g.w);
                    }
                    __syncwarp();       // every lane is done with gate slot `islot` before lane 0 refills it next iteration
                    islot ^= 1;
                    if (!row_valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = 0u;
                    }
                    if (lane == 0) tma_store_wait_read<S::OUT_SLOTS - 1>();
                    __syncwarp();
                    uint8_t* so = my_out + oslot * EPI_SLOT_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(so + epi_slot_off(lane, j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, so, col0, lt * TILE_M + q * 32);
                        tma_store_commit();
                    }
                    oslot ^= 1;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tempty[acc]);
                else mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace mm
