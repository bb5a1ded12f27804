// medmoe_b200 — dZ (gradient w.r.t. the native-resolution attention hidden Z) by interval prefix sums.
// Included by combine.cu after CombineArgs, the row helpers and token_dlogit().
//
//   dZ_s[i, k] = w2[k] * sum_p w_i(p) dl_s(p) [interp(Z_s)(p, k) > 0]          (autograd of swin.py:63-68)
//
// For an integer scale ratio r the tokens between two adjacent native rows (a = Z_s[m], b = Z_s[m+1]) form an
// interval j = 0..r-1 with lambda_j = (j + 0.5) / r, and h_j = a + lambda_j (b - a) changes sign at most once.
// So the set of tokens whose ReLU gate is open is a prefix or a suffix of the interval, and per element
//   A-side (to row m)   = sum_open (1 - lambda_j) dl_j = S0 - S1
//   B-side (to row m+1) = sum_open lambda_j dl_j       = S1
//   dw2[k]             += sum_open dl_j h_j            = a S0 + (b - a) S1
// with S0 / S1 read from exclusive prefix sums of dl_j and lambda_j dl_j over the interval: O(1) per
// (row, element) instead of O(r).  The r/2 clamped tokens before the first and after the last native row
// ("head" / "tail") see h = Z_s[0] resp. Z_s[P_s - 1] with their whole weight on that row.
//
// Kernels (all deterministic, no atomics):
//   bwd_z_dlogit_kernel : dlogit[slot, p, 0..3] from beta and the two column-half partial dbeta (+ gate gradient)
//   bwd_z_prefix_kernel : per (slot, coarse scale, interval) the r+1 exclusive prefix sums, head / tail sums
//   bwd_z_ident_kernel  : finest scale (r = 1): dZ_0[p] = dl_0(p) w2 [Z_0[p] > 0]  (+ dw2, db1, db2 partials)
//   bwd_z_rows_kernel   : coarse scales, a warp walks 16 consecutive native rows (+ dw2, db1 partials); the ratio-4 scale
//                         (three quarters of the coarse rows) evaluates its four tokens per interval directly instead
//                         of reading prefix sums (0.26 -> 0.21 ms at cfg2)
#pragma once

namespace mm {

constexpr int ZR_ROWS_PER_WARP = 16;
constexpr int ZR_WARPS = 8;

// scratch layout per slot (floats): dlog [P][4] | for s = 1..3: PA [(Ps-1)(r+1)] x {PA0, PA1} interleaved | HT [3][2]
struct ZScratch {
    long long per_slot;
    long long pa_off[4];       // offset of the {PA0, PA1} pairs of scale s (8-byte aligned)
    long long pa_len[4];
    long long ht_off;
};
__host__ __device__ inline ZScratch z_scratch_layout(int P, const int* Ps) {
    ZScratch z{};
    long long off = 4LL * P;
    for (int s = 1; s < 4; ++s) {
        const int r = P / Ps[s];
        z.pa_len[s] = static_cast<long long>(Ps[s] - 1 > 0 ? Ps[s] - 1 : 0) * (r + 1);
        z.pa_off[s] = off;
        off += 2 * z.pa_len[s];
    }
    z.ht_off = off;
    off += 8;
    z.per_slot = (off + 3) / 4 * 4;      // dlog is accessed as float4: keep every slot 16-byte aligned
    return z;
}

// grid = (ceil(P / 256), n_items)
__global__ void __launch_bounds__(256)
bwd_z_dlogit_kernel(const CombineArgs a, const ZScratch zs) {
    const int slot = blockIdx.y;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const int item = a.perm[slot];
    const float g = a.gate ? a.gate[item] : 1.0f;
    float dot = 0.f;
    if (p < a.P) {
        const float4 dl = token_dlogit(a, slot, p, g, dot);
        *reinterpret_cast<float4*>(a.zscr + slot * zs.per_slot + 4LL * p) = dl;
    }
    if (a.dgate) {   // top-k extension: d gate = sum_p <beta, dbeta>
        __shared__ float sm[8];
        dot = warp_sum(p < a.P ? dot : 0.f);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += sm[w];
            atomicAdd(a.dgate + item, t);
        }
    }
}

// grid = (ceil(max_s (Ps - 1) / 256), n_items, 3); thread = interval m of scale s = blockIdx.z + 1
__global__ void __launch_bounds__(256)
bwd_z_prefix_kernel(const CombineArgs a, const ZScratch zs) {
    const int slot = blockIdx.y;
    const int s = blockIdx.z + 1;
    const int m = blockIdx.x * 256 + threadIdx.x;
    const int Ps = a.Ps[s], r = a.ratio[s];
    float* base = a.zscr + slot * zs.per_slot;
    const float* dlog = base;
    if (m < Ps - 1 && r != 4) {      // ratio 4 is evaluated directly by the rows kernel (z_open_sums_r4): no prefix table
        float2* pa = reinterpret_cast<float2*>(base + zs.pa_off[s]) + static_cast<long long>(m) * (r + 1);
        const int p0 = r * m + (r >> 1);
        const float inv_r = 1.0f / static_cast<float>(r);
        float run0 = 0.f, run1 = 0.f;
        for (int j = 0; j < r; ++j) {
            pa[j] = make_float2(run0, run1);
            const float dl = dlog[4LL * (p0 + j) + s];
            run0 += dl;
            run1 = fmaf((static_cast<float>(j) + 0.5f) * inv_r, dl, run1);
        }
        pa[r] = make_float2(run0, run1);
    }
    if (m == 0) {   // clamped tokens: head [0, r/2) -> row 0, tail [P - r/2, P) -> row Ps - 1
        float hs = 0.f, ts = 0.f;
        for (int p = 0; p < (r >> 1); ++p) hs += dlog[4LL * p + s];
        for (int p = a.P - (r >> 1); p < a.P; ++p) ts += dlog[4LL * p + s];
        base[zs.ht_off + (s - 1) * 2 + 0] = hs;
        base[zs.ht_off + (s - 1) * 2 + 1] = ts;
    }
}

// per-CTA partials [dw2 (H) | db1 (H) | db2] -> a.part[(slot * nrb + blk) * PART ...]
template <int D, int NWARPS>
MM_DEVINL void z_write_partials(const CombineArgs& a, int slot, int blk, float (*s_part)[2 * (D / 2) + 1], int warp, int lane,
                                const float (&dw2)[(D / 256) * 4], const float (&db1)[(D / 256) * 4], float db2) {
    constexpr int NE = D / 256;
    constexpr int H = D / 2;
    constexpr int PART = 2 * H + 1;
#pragma unroll
    for (int t = 0; t < NE; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s_part[warp][4 * (lane + 32 * t) + k] = dw2[4 * t + k];
            s_part[warp][H + 4 * (lane + 32 * t) + k] = db1[4 * t + k];
        }
    if (lane == 0) s_part[warp][2 * H] = db2;
    __syncthreads();
    float* dst = a.part + (static_cast<size_t>(slot) * a.nrb + blk) * PART;
    for (int cc = threadIdx.x; cc < PART; cc += blockDim.x) {
        float accv = 0.f;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) accv += s_part[w][cc];
        dst[cc] = accv;
    }
}

// finest scale; grid = (ceil(nruns / 8), n_items + K); warp = run of 32 tokens; blocks y >= n_items zero dZ's padding
template <int D>
__global__ void __launch_bounds__(256)
bwd_z_ident_kernel(const CombineArgs a, const ZScratch zs) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    __shared__ float s_part[8][2 * H + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.y >= a.n_items) {
        const int e = blockIdx.y - a.n_items;
        for (int s = 0; s < 4; ++s) {
            const long long rows = static_cast<long long>(a.counts[e]) * a.Ps[s];
            const long long pad = (rows + TILE_M - 1) / TILE_M * TILE_M - rows;
            __nv_bfloat16* dz = a.dZ + (static_cast<long long>(a.seg_start[s * a.K + e]) + rows) * H;
            for (long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 4; i < pad * H;
                 i += static_cast<long long>(gridDim.x) * 256 * 4)
                *reinterpret_cast<uint2*>(dz + i) = make_uint2(0, 0);
        }
        return;
    }
    const int slot = blockIdx.y;
    const int c = blockIdx.x * 8 + warp;
    const int t0 = c * RUN_TOKENS;
    float dw2[E], db1[E], db2 = 0.f;
#pragma unroll
    for (int k = 0; k < E; ++k) { dw2[k] = 0.f; db1[k] = 0.f; }
    if (c < a.nruns) {
        const int e = a.slot_expert[slot];
        float w2[E];
        load_row_f32x4<NE>(a.w2 + static_cast<size_t>(e) * H, lane, w2);
        const long long base = a.slot_row[slot];                 // scale 0
        const float* dlog = a.zscr + slot * zs.per_slot;
        const int n_tok = min(RUN_TOKENS, a.P - t0);
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < n_tok) mine = *reinterpret_cast<const float4*>(dlog + 4LL * (t0 + lane));
        db2 = mine.x + mine.y + mine.z + mine.w;                 // summed over lanes by the partial reduction below
        // eight rows per batch: all of a batch's loads are issued before its first use, so a warp keeps 8 x D bytes in flight
        constexpr int ZB = 8;
        for (int tb = 0; tb < n_tok; tb += ZB) {
            uint2 raw[ZB][NE];
#pragma unroll
            for (int u = 0; u < ZB; ++u) {
                const int t = min(tb + u, n_tok - 1);
#pragma unroll
                for (int c = 0; c < NE; ++c)
                    raw[u][c] = *reinterpret_cast<const uint2*>(a.Z + (base + t0 + t) * H + 4 * (lane + 32 * c));
            }
#pragma unroll
            for (int u = 0; u < ZB; ++u) {
                const int t = tb + u;
                if (t >= n_tok) break;
                const float dl = __shfl_sync(0xffffffffu, mine.x, t);
                float gk[E];
#pragma unroll
                for (int c = 0; c < NE; ++c) {
                    const float zv[4] = {bf16lo(raw[u][c].x), bf16hi(raw[u][c].x), bf16lo(raw[u][c].y), bf16hi(raw[u][c].y)};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = 4 * c + i;
                        const bool open = zv[i] > 0.f;
                        gk[k] = open ? dl * w2[k] : 0.f;
                        if (open) dw2[k] = fmaf(dl, zv[i], dw2[k]);
                        db1[k] += gk[k];
                    }
                }
                store_slab_bf16<NE>(a.dZ + (base + t0 + t) * H, lane, gk);
            }
        }
        db2 = warp_sum(db2);
    }
    z_write_partials<D, 8>(a, slot, blockIdx.x, s_part, warp, lane, dw2, db1, lane == 0 ? db2 : 0.f);
}

// S0 / S1 of the open-gate token set of interval m for one element with end values (za, zb): the open set is the
// token range [lo, hi) of the interval (a prefix, a suffix, everything or nothing), read from the prefix sums.
// Branch-free: lanes of a warp see all four cases.  The crossing point uses the approximate reciprocal: a gate can only
// differ from the forward's when interp(Z) is within ~1e-6 relative of zero, far below the bf16 storage noise of Z
// (za == zb gives inf/NaN here, but then oa == ob and the crossing is not used; NaN converts to 0).
MM_DEVINL void z_open_sums(float za, float zb, int r, const float2* __restrict__ pa, float& S0, float& S1) {
    const bool oa = za > 0.f, ob = zb > 0.f;
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(za - zb));
    const float x = za * rc * static_cast<float>(r) - 0.5f;                      // crossing, in token units
    const int n_pre = min(max(__float2int_ru(x), 0), r);                        // open for lambda_j < t: tokens j < x
    const int n_suf = min(max(__float2int_rd(x) + 1, 0), r);                    // open for lambda_j > t: tokens j > x
    const int hi = oa ? (ob ? r : n_pre) : (ob ? r : 0);
    const int lo = (ob && !oa) ? n_suf : 0;
    const float2 h = __ldg(pa + hi), l = __ldg(pa + lo);
    S0 = h.x - l.x;
    S1 = h.y - l.y;
}

// The same sums for r = 4 (the scale with three quarters of the coarse rows) evaluated directly: four lerp FMAs and four
// predicated adds per element instead of the crossing point, two conversions and two gathers — about half the instructions.
// dl[j] are the interval's four token gradients (warp-uniform), lambda_j = (j + 0.5) / 4.
MM_DEVINL void z_open_sums_r4(float za, float zb, const float (&dl)[4], float& S0, float& S1) {
    const float d = zb - za;
    S0 = 0.f; S1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float lam = (static_cast<float>(j) + 0.5f) * 0.25f;
        if (fmaf(lam, d, za) > 0.f) { S0 += dl[j]; S1 = fmaf(lam, dl[j], S1); }
    }
}
// token gradients of interval m of a ratio-4 scale s (tokens 4 m + 2 .. 4 m + 5), read from the per-slot dlogit table
MM_DEVINL void z_load_dl4(const float* __restrict__ dlog, int m, int s, float (&dl)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dl[j] = __ldg(dlog + 4LL * (4 * m + 2 + j) + s);
}

template <int N>
MM_DEVINL void z_load_row_raw(const __nv_bfloat16* row, int lane, uint2 (&q)[N]) {
#pragma unroll
    for (int t = 0; t < N; ++t) q[t] = *reinterpret_cast<const uint2*>(row + 4 * (lane + 32 * t));
}

// coarse scales; grid = (ceil(chunks / 8), n_items); warp = ZR_ROWS_PER_WARP consecutive native rows of one scale
template <int D>
__global__ void __launch_bounds__(ZR_WARPS * 32, 2)   // <= 128 registers: two blocks (16 warps) per SM hide the row-load latency
bwd_z_rows_kernel(const CombineArgs a, const ZScratch zs, int chunks1, int chunks2, int chunks3, int blk_base) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    __shared__ float s_part[ZR_WARPS][2 * H + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.y;
    int ch = blockIdx.x * ZR_WARPS + warp;
    float dw2[E], db1[E];
#pragma unroll
    for (int k = 0; k < E; ++k) { dw2[k] = 0.f; db1[k] = 0.f; }
    int s = 0;
    if (ch < chunks1) s = 1;
    else if ((ch -= chunks1) < chunks2) s = 2;
    else if ((ch -= chunks2) < chunks3) s = 3;
    if (s != 0) {
        const int Ps = a.Ps[s], r = a.ratio[s];
        const int e = a.slot_expert[slot];
        float w2[E];
        load_row_f32x4<NE>(a.w2 + static_cast<size_t>(e) * H, lane, w2);
        const long long base = a.slot_row[s * a.n_items + slot];
        const float* sbase = a.zscr + slot * zs.per_slot;
        const float2* PA = reinterpret_cast<const float2*>(sbase + zs.pa_off[s]);
        const float hs = sbase[zs.ht_off + (s - 1) * 2 + 0], ts = sbase[zs.ht_off + (s - 1) * 2 + 1];
        const int i_a = ch * ZR_ROWS_PER_WARP, i_b = min(Ps, i_a + ZR_ROWS_PER_WARP);
        float za[E], zb[E], carry[E];
        uint2 n1[NE], n2[NE];     // Z[i + 2], Z[i + 3] in flight, still packed bf16 (two rows of prefetch in the registers of one)
#pragma unroll
        for (int k = 0; k < E; ++k) carry[k] = 0.f;
#pragma unroll
        for (int t = 0; t < NE; ++t) { n1[t] = make_uint2(0u, 0u); n2[t] = make_uint2(0u, 0u); }
        load_row_bf16x4<NE>(a.Z + (base + i_a) * H, lane, zb);              // zb = Z[i_a]
        if (i_a + 1 < Ps) z_load_row_raw<NE>(a.Z + (base + i_a + 1) * H, lane, n1);
        if (i_a + 2 < Ps && i_a + 1 < i_b) z_load_row_raw<NE>(a.Z + (base + i_a + 2) * H, lane, n2);
        const bool direct4 = r == 4;
        if (i_a >= 1) {   // B-side of the interval that ends in the first row of this chunk
            load_row_bf16x4<NE>(a.Z + (base + i_a - 1) * H, lane, za);
            if (direct4) {
                float dl[4];
                z_load_dl4(sbase, i_a - 1, s, dl);
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    float S0, S1;
                    z_open_sums_r4(za[k], zb[k], dl, S0, S1);
                    carry[k] = S1;
                }
            } else {
                const float2* pa = PA + static_cast<long long>(i_a - 1) * (r + 1);
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    float S0, S1;
                    z_open_sums(za[k], zb[k], r, pa, S0, S1);
                    carry[k] = S1;
                }
            }
        }
        for (int i = i_a; i < i_b; ++i) {
#pragma unroll
            for (int k = 0; k < E; ++k) za[k] = zb[k];                        // za = Z[i]
#pragma unroll
            for (int t = 0; t < NE; ++t) {                                      // zb = Z[i + 1]
                zb[4 * t + 0] = bf16lo(n1[t].x); zb[4 * t + 1] = bf16hi(n1[t].x);
                zb[4 * t + 2] = bf16lo(n1[t].y); zb[4 * t + 3] = bf16hi(n1[t].y);
                n1[t] = n2[t];
            }
            if (i + 3 < Ps && i + 2 < i_b) z_load_row_raw<NE>(a.Z + (base + i + 3) * H, lane, n2);     // prefetch Z[i + 3]
            float row[E];
#pragma unroll
            for (int k = 0; k < E; ++k) row[k] = carry[k];
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < E; ++k)
                    if (za[k] > 0.f) { row[k] += hs; dw2[k] = fmaf(hs, za[k], dw2[k]); }
            }
            if (i == Ps - 1) {
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    if (za[k] > 0.f) { row[k] += ts; dw2[k] = fmaf(ts, za[k], dw2[k]); }
                    carry[k] = 0.f;
                }
            } else if (direct4) {
                float dl[4];
                z_load_dl4(sbase, i, s, dl);
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    float S0, S1;
                    z_open_sums_r4(za[k], zb[k], dl, S0, S1);
                    row[k] += S0 - S1;
                    carry[k] = S1;
                    dw2[k] += za[k] * S0 + (zb[k] - za[k]) * S1;
                }
            } else {
                const float2* pa = PA + static_cast<long long>(i) * (r + 1);
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    float S0, S1;
                    z_open_sums(za[k], zb[k], r, pa, S0, S1);
                    row[k] += S0 - S1;
                    carry[k] = S1;
                    dw2[k] += za[k] * S0 + (zb[k] - za[k]) * S1;
                }
            }
            float o[E];
#pragma unroll
            for (int k = 0; k < E; ++k) { o[k] = w2[k] * row[k]; db1[k] += o[k]; }
            store_slab_bf16<NE>(a.dZ + (base + i) * H, lane, o);
        }
    }
    z_write_partials<D, ZR_WARPS>(a, slot, blk_base + blockIdx.x, s_part, warp, lane, dw2, db1, 0.f);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// the interval formulation needs the finest scale at full resolution and even integer ratios for the others
static inline bool z_rows_path_ok(const CombineArgs& a) {
    if (a.Ps[0] != a.P) return false;
    for (int s = 1; s < 4; ++s) {
        if (a.Ps[s] <= 0 || a.P % a.Ps[s] != 0) return false;
        const int r = a.P / a.Ps[s];
        if (r < 2 || (r & 1)) return false;
    }
    return true;
}
static inline int z_ident_blocks(const CombineArgs& a) { return (a.nruns + 7) / 8; }
static inline int z_row_chunks(const CombineArgs& a, int s) { return (a.Ps[s] + ZR_ROWS_PER_WARP - 1) / ZR_ROWS_PER_WARP; }
static inline int z_rows_blocks(const CombineArgs& a) {
    return (z_row_chunks(a, 1) + z_row_chunks(a, 2) + z_row_chunks(a, 3) + ZR_WARPS - 1) / ZR_WARPS;
}

// launches the four kernels; a.nrb must be z_ident_blocks + z_rows_blocks, a.ratio[] filled
template <int D>
static int launch_bwd_z_rows_path(const CombineArgs& a, cudaStream_t st) {
    int Ps[4] = {a.Ps[0], a.Ps[1], a.Ps[2], a.Ps[3]};
    const ZScratch zs = z_scratch_layout(a.P, Ps);
    bwd_z_dlogit_kernel<<<dim3((a.P + 255) / 256, a.n_items), 256, 0, st>>>(a, zs);
    note_launches(1);
    trace_mark("combine_bwd.dZ.dlogit", st);
    int max_iv = 1;
    for (int s = 1; s < 4; ++s) max_iv = a.Ps[s] - 1 > max_iv ? a.Ps[s] - 1 : max_iv;
    bwd_z_prefix_kernel<<<dim3((max_iv + 255) / 256, a.n_items, 3), 256, 0, st>>>(a, zs);
    note_launches(1);
    trace_mark("combine_bwd.dZ.prefix", st);
    bwd_z_ident_kernel<D><<<dim3(z_ident_blocks(a), a.n_items + a.K), 256, 0, st>>>(a, zs);
    note_launches(1);
    trace_mark("combine_bwd.dZ.ident", st);
    bwd_z_rows_kernel<D><<<dim3(z_rows_blocks(a), a.n_items), ZR_WARPS * 32, 0, st>>>(
        a, zs, z_row_chunks(a, 1), z_row_chunks(a, 2), z_row_chunks(a, 3), z_ident_blocks(a));
    note_launches(1);
    trace_mark("combine_bwd.dZ.rows", st);
    return check_launch("combine_bwd(dZ interval path)");
}

}  // namespace mm
