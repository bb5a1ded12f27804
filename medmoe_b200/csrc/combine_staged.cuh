// medmoe_b200 — TMA-staged persistent combine kernels (forward logits, forward combine, backward dbeta).
// Included by combine.cu after CombineArgs and the row load/store helpers.
//
// One CTA per SM walks over (item, token-tile) work items.  A producer thread stages, per scale,
// the contiguous range of native rows a tile of TT tokens touches with ONE bulk copy
// (cp.async.bulk global -> shared, completion on an mbarrier), one or more tiles ahead; sixteen
// consumer warps read the rows back with conflict-free LDS and never issue a dependent global load.
// Loads in flight per SM are set by the stage depth (~75 KB), not by registers or occupancy.
// The consumers split the channel width in two (warp = (token group, column half)) so that a thread
// needs ~80 registers and all sixteen warps (4 per scheduler) stay resident.
#pragma once

namespace mm {

constexpr int SG_CONSUMER_WARPS = 16;
constexpr int SG_THREADS = (SG_CONSUMER_WARPS + 1) * 32;
constexpr int SG_TOKEN_GROUPS = SG_CONSUMER_WARPS / 2;      // for the column-split kernels

struct SgTileRows { int i_lo[4]; int n[4]; };
MM_DEVINL SgTileRows sg_tile_rows(const CombineArgs& a, int t0, int t1) {
    SgTileRows r;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const LerpSrc A = lerp_src(t0, a.scale[s], a.Ps[s]);
        const LerpSrc B = lerp_src(t1 - 1, a.scale[s], a.Ps[s]);
        r.i_lo[s] = A.i0;
        r.n[s] = B.i1 - A.i0 + 1;
    }
    return r;
}

// producer: stage the rows of tile (slot, [t0, t1)) of the [rows, W] bf16 matrix `mat`
template <int W>
MM_DEVINL void sg_stage_tile(const CombineArgs& a, const __nv_bfloat16* mat, int slot, int t0, int t1, uint8_t* dst, uint64_t* bar,
                             uint32_t extra_bytes) {
    const SgTileRows r = sg_tile_rows(a, t0, t1);
    uint32_t bytes = extra_bytes;
#pragma unroll
    for (int s = 0; s < 4; ++s) bytes += static_cast<uint32_t>(r.n[s]) * W * 2;
    mbar_expect_tx(bar, bytes);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const long long row = static_cast<long long>(a.slot_row[s * a.n_items + slot]) + r.i_lo[s];
        bulk_load_1d(dst + static_cast<size_t>(a.cap_off[s]) * W * 2, mat + row * W, static_cast<uint32_t>(r.n[s]) * W * 2, bar);
    }
}

// 4*NE bf16 values of a staged row segment: lane owns the 8-byte chunks lane, lane + 32, ...
template <int NE>
MM_DEVINL void sg_lds_x4(const uint8_t* seg, int lane, float (&f)[NE * 4]) {
#pragma unroll
    for (int t = 0; t < NE; ++t) {
        const uint2 u = *reinterpret_cast<const uint2*>(seg + 8 * (lane + 32 * t));
        f[4 * t + 0] = bf16lo(u.x); f[4 * t + 1] = bf16hi(u.x); f[4 * t + 2] = bf16lo(u.y); f[4 * t + 3] = bf16hi(u.y);
    }
}

// ---- forward pass 1: logits over scales + softmax -> beta.  Work item = (item, tile); Z rows staged. ----
template <int D, int STAGES>
__global__ void __launch_bounds__(SG_THREADS, 1)
sg_logits_kernel(const CombineArgs a) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const size_t stage_bytes = static_cast<size_t>(a.cap_total) * H * 2;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SG_CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    const int total = a.n_items * a.tiles_per_img;
    const int TT = a.tile_tokens;
    if (warp == SG_CONSUMER_WARPS) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                const int slot = w / a.tiles_per_img, tt = w - slot * a.tiles_per_img;
                const int t0 = tt * TT, t1 = min(a.P, t0 + TT);
                mbar_wait(&empty[stage], phase ^ 1);
                sg_stage_tile<H>(a, a.Z, slot, t0, t1, smem + stage * stage_bytes, &full[stage], 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    int stage = 0; uint32_t phase = 0;
    const int tok_per_warp = TT / SG_CONSUMER_WARPS;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / a.tiles_per_img, tt = w - slot * a.tiles_per_img;
        const int t0 = tt * TT, t1 = min(a.P, t0 + TT);
        const int e = a.slot_expert[slot];
        float w2[E];
        load_row_f32x4<NE>(a.w2 + static_cast<size_t>(e) * H, lane, w2);
        const float b2 = a.b2[e];
        const SgTileRows r = sg_tile_rows(a, t0, t1);
        mbar_wait(&full[stage], phase);
        const uint8_t* st = smem + stage * stage_bytes;
        for (int k = 0; k < tok_per_warp; ++k) {
            const int p = t0 + warp * tok_per_warp + k;
            if (p >= t1) break;
            float lg[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                const uint8_t* ra = st + static_cast<size_t>(a.cap_off[s] + L.i0 - r.i_lo[s]) * H * 2;
                float za[E];
                sg_lds_x4<NE>(ra, lane, za);
                float acc = 0.f;
                if (L.i1 != L.i0 && L.lam != 0.f) {
                    float zb[E];
                    sg_lds_x4<NE>(ra + H * 2, lane, zb);
                    const float l0 = 1.0f - L.lam;
#pragma unroll
                    for (int i = 0; i < E; ++i) acc = fmaf(fmaxf(l0 * za[i] + L.lam * zb[i], 0.f), w2[i], acc);
                } else {
#pragma unroll
                    for (int i = 0; i < E; ++i) acc = fmaf(fmaxf(za[i], 0.f), w2[i], acc);
                }
                lg[s] = warp_sum(acc) + b2;
            }
            const float mx = fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3]));
            const float e0 = expf(lg[0] - mx), e1 = expf(lg[1] - mx), e2 = expf(lg[2] - mx), e3 = expf(lg[3] - mx);
            const float inv = 1.0f / (e0 + e1 + e2 + e3);
            if (lane == 0)
                *reinterpret_cast<float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4) =
                    make_float4(e0 * inv, e1 * inv, e2 * inv, e3 * inv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
}

// ---- forward pass 2: out = sum_s beta_s interp(Y_s).  Work item = (image, tile, top-k choice); Y rows staged;
// consumer warp = (token group, column half). ----
template <int D, typename OutT, int STAGES>
__global__ void __launch_bounds__(SG_THREADS, 1)
sg_out_kernel(const CombineArgs a) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const size_t stage_bytes = static_cast<size_t>(a.cap_total) * D * 2;
    float* s_g = reinterpret_cast<float*>(smem + STAGES * stage_bytes);          // [token groups][D]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_g + SG_TOKEN_GROUPS * D);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SG_CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    const int total = a.B * a.tiles_per_img;       // (image, tile); the top-k choices are the inner pipeline items
    const int TT = a.tile_tokens;
    if (warp == SG_CONSUMER_WARPS) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                const int b = w / a.tiles_per_img, tt = w - b * a.tiles_per_img;
                const int t0 = tt * TT, t1 = min(a.P, t0 + TT);
                for (int jk = 0; jk < a.topk; ++jk) {
                    const int slot = a.inv_perm[b * a.topk + jk];
                    mbar_wait(&empty[stage], phase ^ 1);
                    sg_stage_tile<D>(a, a.Y, slot, t0, t1, smem + stage * stage_bytes, &full[stage], 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }
    int stage = 0; uint32_t phase = 0;
    const int grp = warp >> 1;
    const int col0 = (warp & 1) * H;
    const int tok_per_warp = TT / SG_TOKEN_GROUPS;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / a.tiles_per_img, tt = w - b * a.tiles_per_img;
        const int t0 = tt * TT, t1 = min(a.P, t0 + TT);
        const SgTileRows r = sg_tile_rows(a, t0, t1);
        float gsum[E];
#pragma unroll
        for (int i = 0; i < E; ++i) gsum[i] = 0.f;
        for (int jk = 0; jk < a.topk; ++jk) {
            const int item = b * a.topk + jk;
            const int slot = a.inv_perm[item];
            const float g = a.gate ? a.gate[item] : 1.0f;
            // beta of this warp's tokens: issued before the stage wait so the latency overlaps it
            float4 btv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = t0 + grp * tok_per_warp + k;
                btv[k] = (k < tok_per_warp && p < t1) ? *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&full[stage], phase);
            const uint8_t* st = smem + stage * stage_bytes + col0 * 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = t0 + grp * tok_per_warp + k;
                if (k >= tok_per_warp || p >= t1) break;
                const float bt[4] = {btv[k].x * g, btv[k].y * g, btv[k].z * g, btv[k].w * g};
                float o[E];
#pragma unroll
                for (int i = 0; i < E; ++i) o[i] = 0.f;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                    const uint8_t* ra = st + static_cast<size_t>(a.cap_off[s] + L.i0 - r.i_lo[s]) * D * 2;
                    float ya[E];
                    sg_lds_x4<NE>(ra, lane, ya);
                    if (L.i1 != L.i0 && L.lam != 0.f) {
                        float yb[E];
                        sg_lds_x4<NE>(ra + D * 2, lane, yb);
                        const float c0 = bt[s] * (1.0f - L.lam), c1 = bt[s] * L.lam;
#pragma unroll
                        for (int i = 0; i < E; ++i) o[i] = fmaf(c0, ya[i], fmaf(c1, yb[i], o[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < E; ++i) o[i] = fmaf(bt[s], ya[i], o[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < E; ++i) gsum[i] += o[i];
                OutT* orow = static_cast<OutT*>(a.out) + (static_cast<size_t>(b) * a.P + p) * D + col0;
                if (jk > 0) {   // top-k extension: add onto the previous choice's contribution (same thread wrote it)
                    float prev[E];
                    load_slab<NE, OutT>(orow, lane, prev);
#pragma unroll
                    for (int i = 0; i < E; ++i) o[i] += prev[i];
                }
                if constexpr (sizeof(OutT) == 2) store_slab_bf16<NE>(reinterpret_cast<__nv_bfloat16*>(orow), lane, o);
                else store_slab_f32<NE>(reinterpret_cast<float*>(orow), lane, o);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // deterministic partial of the global mean: one [D] vector per (image, tile)
#pragma unroll
        for (int t = 0; t < NE; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) s_g[grp * D + col0 + 4 * (lane + 32 * t) + i] = gsum[4 * t + i];
        named_bar_sync(1, SG_CONSUMER_WARPS * 32);
        for (int d = threadIdx.x; d < D; d += SG_CONSUMER_WARPS * 32) {
            float acc = 0.f;
#pragma unroll
            for (int ww = 0; ww < SG_TOKEN_GROUPS; ++ww) acc += s_g[ww * D + d];
            a.gpart[(static_cast<size_t>(b) * a.nblk + tt) * D + d] = acc;
        }
        named_bar_sync(1, SG_CONSUMER_WARPS * 32);
    }
}

// ---- backward pass A: per column half, dbeta_s = <dF, interp(Y_s)> -> dlogit buffer as [n_items, P, 2 halves, 4]
// (the dZ kernel adds the halves and applies the softmax-over-scales backward).
// Work item = (image, tile, top-k choice); stage = Y rows of the tile [+ the dlocal rows of the tile].
template <int D, typename OutT, int STAGES>
__global__ void __launch_bounds__(SG_THREADS, 1)
sg_bwd_dbeta_kernel(const CombineArgs a) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int TT = a.tile_tokens;
    const size_t y_bytes = static_cast<size_t>(a.cap_total) * D * 2;
    const size_t df_bytes = a.dlocal ? static_cast<size_t>(TT) * D * sizeof(OutT) : 0;
    const size_t stage_bytes = y_bytes + df_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SG_CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    const int total = a.B * a.tiles_per_img;
    if (warp == SG_CONSUMER_WARPS) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                const int b = w / a.tiles_per_img, tt = w - b * a.tiles_per_img;
                const int t0 = tt * TT, t1 = min(a.P, t0 + TT);
                for (int jk = 0; jk < a.topk; ++jk) {
                    const int slot = a.inv_perm[b * a.topk + jk];
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* dst = smem + stage * stage_bytes;
                    const uint32_t extra = a.dlocal ? static_cast<uint32_t>(t1 - t0) * D * sizeof(OutT) : 0u;
                    sg_stage_tile<D>(a, a.Y, slot, t0, t1, dst, &full[stage], extra);
                    if (a.dlocal)
                        bulk_load_1d(dst + y_bytes, static_cast<const OutT*>(a.dlocal) + (static_cast<size_t>(b) * a.P + t0) * D, extra,
                                     &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }
    int stage = 0; uint32_t phase = 0;
    const int grp = warp >> 1;
    const int half = warp & 1;
    const int col0 = half * H;
    const int tok_per_warp = TT / SG_TOKEN_GROUPS;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / a.tiles_per_img, tt = w - b * a.tiles_per_img;
        const int t0 = tt * TT, t1 = min(a.P, t0 + TT);
        const SgTileRows r = sg_tile_rows(a, t0, t1);
        float dg[E];
        if (a.dglobal) {
            load_slab<NE, float>(a.dglobal + static_cast<size_t>(b) * D + col0, lane, dg);
            const float inv_p = 1.0f / static_cast<float>(a.P);
#pragma unroll
            for (int i = 0; i < E; ++i) dg[i] *= inv_p;
        } else {
#pragma unroll
            for (int i = 0; i < E; ++i) dg[i] = 0.f;
        }
        for (int jk = 0; jk < a.topk; ++jk) {
            const int slot = a.inv_perm[b * a.topk + jk];
            mbar_wait(&full[stage], phase);
            const uint8_t* st = smem + stage * stage_bytes;
            for (int k = 0; k < tok_per_warp; ++k) {
                const int p = t0 + grp * tok_per_warp + k;
                if (p >= t1) break;
                float df[E];
                if (a.dlocal) {
                    const OutT* drow = reinterpret_cast<const OutT*>(st + y_bytes) + static_cast<size_t>(p - t0) * D + col0;
                    load_slab<NE, OutT>(drow, lane, df);      // generic load from shared memory
#pragma unroll
                    for (int i = 0; i < E; ++i) df[i] += dg[i];
                } else {
#pragma unroll
                    for (int i = 0; i < E; ++i) df[i] = dg[i];
                }
                float dbeta[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                    const uint8_t* ra = st + static_cast<size_t>(a.cap_off[s] + L.i0 - r.i_lo[s]) * D * 2 + col0 * 2;
                    float ya[E];
                    sg_lds_x4<NE>(ra, lane, ya);
                    float acc = 0.f;
                    if (L.i1 != L.i0 && L.lam != 0.f) {
                        float yb[E];
                        sg_lds_x4<NE>(ra + D * 2, lane, yb);
                        const float l0 = 1.0f - L.lam;
#pragma unroll
                        for (int i = 0; i < E; ++i) acc = fmaf(df[i], l0 * ya[i] + L.lam * yb[i], acc);
                    } else {
#pragma unroll
                        for (int i = 0; i < E; ++i) acc = fmaf(df[i], ya[i], acc);
                    }
                    dbeta[s] = warp_sum(acc);
                }
                if (lane == 0)
                    *reinterpret_cast<float4*>(a.dlogit + ((static_cast<size_t>(slot) * a.P + p) * 2 + half) * 4) =
                        make_float4(dbeta[0], dbeta[1], dbeta[2], dbeta[3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr size_t SG_SMEM_LIMIT = 227 * 1024;

// tile geometry: rows reserved per scale in a stage = ceil(TT * Ps / P) + 3 (covers both lerp neighbours)
static inline void sg_setup_tiles(CombineArgs& a, int TT) {
    a.tile_tokens = TT;
    a.tiles_per_img = (a.P + TT - 1) / TT;
    int off = 0;
    for (int s = 0; s < 4; ++s) {
        a.cap[s] = static_cast<int>((static_cast<long long>(TT) * a.Ps[s] + a.P - 1) / a.P) + 3;
        a.cap_off[s] = off;
        off += a.cap[s];
    }
    a.cap_total = off;
}
template <typename K>
static int sg_opt_in(K kern, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) {
        set_error("%s: cannot opt in to %zu B of shared memory (%s)", what, bytes, cudaGetErrorString(e));
        return MM_ERR_CUDA;
    }
    return MM_OK;
}
// each launcher returns 0 = launched, 1 = the tile does not fit in shared memory (caller falls back), < 0 = error
template <int D>
static int sg_launch_logits(const CombineArgs& a, cudaStream_t st) {
    constexpr int STAGES = 4;
    const size_t smem = STAGES * static_cast<size_t>(a.cap_total) * (D / 2) * 2 + 2 * STAGES * 8 + 128;
    if (smem > SG_SMEM_LIMIT || a.tile_tokens % SG_CONSUMER_WARPS != 0) return 1;
    auto kern = sg_logits_kernel<D, STAGES>;
    if (int rc = sg_opt_in(kern, smem, "combine_logits")) return rc;
    const int total = a.n_items * a.tiles_per_img;
    kern<<<total < sm_count() ? total : sm_count(), SG_THREADS, smem, st>>>(a);
    note_launches(1);
    return MM_OK;
}
template <int D, typename OutT>
static int sg_launch_out(const CombineArgs& a, cudaStream_t st) {
    constexpr int STAGES = 2;
    const size_t smem = STAGES * static_cast<size_t>(a.cap_total) * D * 2 + SG_TOKEN_GROUPS * D * 4 + 2 * STAGES * 8 + 128;
    if (smem > SG_SMEM_LIMIT || a.tile_tokens % SG_TOKEN_GROUPS != 0 || a.tile_tokens / SG_TOKEN_GROUPS > 4) return 1;
    auto kern = sg_out_kernel<D, OutT, STAGES>;
    if (int rc = sg_opt_in(kern, smem, "combine_out")) return rc;
    const int total = a.B * a.tiles_per_img;
    kern<<<total < sm_count() ? total : sm_count(), SG_THREADS, smem, st>>>(a);
    note_launches(1);
    return MM_OK;
}
template <int D, typename OutT>
static int sg_launch_bwd_dbeta(const CombineArgs& a, cudaStream_t st) {
    constexpr int STAGES = 2;
    const size_t stage = static_cast<size_t>(a.cap_total) * D * 2 + (a.dlocal ? static_cast<size_t>(a.tile_tokens) * D * sizeof(OutT) : 0);
    const size_t smem = STAGES * stage + 2 * STAGES * 8 + 128;
    if (smem > SG_SMEM_LIMIT || a.tile_tokens % SG_TOKEN_GROUPS != 0) return 1;
    auto kern = sg_bwd_dbeta_kernel<D, OutT, STAGES>;
    if (int rc = sg_opt_in(kern, smem, "combine_bwd_dbeta")) return rc;
    const int total = a.B * a.tiles_per_img;
    kern<<<total < sm_count() ? total : sm_count(), SG_THREADS, smem, st>>>(a);
    note_launches(1);
    return MM_OK;
}

}  // namespace mm
