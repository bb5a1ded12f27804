// medmoe_b200 — interpolate + scale-softmax + weighted combine / scatter-back
// (north-star kernel 4; SURVEY §2.2 rows E2,E3,E5,E6,E7,M2,M3 fused into one pass).
//
// Reference arithmetic being restated (swin.py:38-80, 105-113), per image and selected expert:
//     U_s   = interp_linear(Y_s -> P tokens)                      Y_s = ReLU(conv1x1(f_s)), native resolution
//     logit = w2 . ReLU(W1 U_s + b1) + b2  = w2 . ReLU(interp(Z_s)) + b2   with Z_s = Y_s W1^T + b1
//             (the first Linear commutes with the lerp because lerp weights sum to 1 — SURVEY §8a a6)
//     beta  = softmax_s(logit)
//     fused = sum_s beta_s U_s           -> local_feat (token-major [B, P, D]; the reference's
//                                           [B, D, sqrt(P), sqrt(P)] result is a stride-view of it)
//     global_feat = mean_p fused
// The store goes to the image's ORIGINAL batch slot (scatter-back folded into the store).
//
// Backward (no reference code — it is autograd of the above), dF(p) = dlocal(p) + dglobal / P:
//   tensor-core / rank-1 path (even integer scale ratios; mm_interp_softmax_combine_bwd_tc):
//        dbeta_s = <dF, U_s>          GEMM against the staged Y rows (cm_dbeta_kernel) + interp of <dglobal, Y[row]> / P (rank-1)
//        dUT_s[i] = sum_p w_i(p) beta_s(p) dlocal(p)      cm_dut_kernel; the dglobal part stays rank-1 (never written)
//        dZ_s[i]  = sum_p w_i(p) dlogit_s(p) w2 * [interp(Z_s)(p) > 0]     interval prefix sums (combine_bwd_z.cuh)
//   CUDA-core fallbacks (any ratio): token-centric kernels below, and the generic gather kernels
//        pass A dlogit[slot, p, 4];  pass B (native-row-centric, the transpose of the lerp as a gather, no atomics) dUT, dZ
//   dw2, db2, db1 partial sums per CTA (reduced per expert by expert_reduce_kernel)
// The forward / backward kernels that run on the tensor cores live in combine_mma.cuh, the rank-1 ones in combine_rank1.cuh.
#include "mm_common.cuh"
#include "api_internal.h"

namespace mm {

#ifndef MM_COMBINE_MINBLOCKS
#define MM_COMBINE_MINBLOCKS 1   // resident CTAs/SM the run-based kernels are compiled for (register cap = 65536 / (256 * n))
#endif
constexpr int CB_TOKENS_PER_WARP = 8;
constexpr int CB_TOKENS_PER_BLOCK = 64;   // 8 warps x 8 tokens
constexpr int CB_ROWS_PER_WARP = 8;
constexpr int CB_ROWS_PER_BLOCK = 64;

struct CombineArgs {
    int B, topk, P, n_items;
    int Ps[4];
    float scale[4];                 // Ps[s] / P as float (ATen's area_pixel_compute_scale)
    const int* inv_perm;            // [n_items] item -> slot
    const int* perm;                // [n_items] slot -> item
    const int* slot_expert;         // [n_items]
    const int* slot_row;            // [4, n_items] global row of the slot's first native row
    const int* counts;              // [K]
    const int* seg_start;           // [4, K]
    int K;
    const float* gate;              // [n_items] gate weight of item, nullptr => 1
    const __nv_bfloat16* Y;         // [rows, D]
    const __nv_bfloat16* Z;         // [rows, D/2]
    const float* w2;                // [E, D/2]
    const float* b2;                // [E]
    float* beta;                    // [n_items, P, 4]
    void* out;                      // [B, P, D]
    float* gpart;                   // [B, nblk, D]
    int nblk;
    // backward
    const void* dlocal;             // [B, P, D] (same dtype as out) or nullptr
    const float* dglobal;           // [B, D] fp32 or nullptr
    float* dlogit;                  // [n_items, P, 4]
    float* dgate;                   // [n_items] (zeroed by caller) or nullptr
    __nv_bfloat16* dUT;             // [rows, D]
    __nv_bfloat16* dZ;              // [rows, D/2]
    float* part;                    // [n_items, nrb, 2*(D/2) + 1] per-CTA partials: dw2 | db1 | db2
    int nrb;
    // token-centric ("fast") backward: integer scale ratios only
    int ratio[4];                   // P / Ps[s]
    int mode[4];                    // SCALE_IDENT / SCALE_DIRECT / SCALE_MOMENT
    int halo;                       // max over DIRECT scales of ratio / 2
    int nruns;                      // ceil(P / 32)
    // TMA-staged persistent kernels: a tile is TT consecutive tokens of one item; per scale the native
    // rows it touches are one contiguous row range, staged in shared memory by one bulk copy each
    int tile_tokens;                // TT
    int tiles_per_img;              // ceil(P / TT)
    int cap[4];                     // rows reserved per scale in a stage: ceil(TT * Ps / P) + 2
    int cap_off[4];                 // prefix sums of cap
    int cap_total;
    float* zscr;                    // scratch of the interval-prefix dZ path (combine_bwd_z.cuh), aliases mom_z
    int z_rows_path;                // dZ comes from the interval-prefix kernels: finalize only writes dUT
    int dlogit_is_halves;           // dlogit holds two column-half partial dbeta ([.., 2, 4]) instead of finished dlogit
    const float* row_dot;           // [rows] <dglobal[b], Y[row]> / P: dbeta is its lerp (rank-1 path, combine_rank1.cuh) or nullptr
    const float* dbeta_loc;         // [n_items, P, 4] <dlocal(p), interp(Y_s)(p)> from the tensor-core kernel (cm_dbeta_kernel) or nullptr
    float* mom_u;                   // [n_items, nruns, 2, D]   zeroth / first moments of beta_s * dF per 32-token run
    float* mom_z;                   // [n_items, nruns, 2, D/2] same for dlogit_s * w2 * gate
};

enum : int { SCALE_IDENT = 0, SCALE_DIRECT = 1, SCALE_MOMENT = 2, SCALE_UNSUPPORTED = 3 };
constexpr int RUN_TOKENS = 32;      // tokens per warp-run in the token-centric backward
constexpr int RUNS_PER_BLOCK = 8;

template <int N>
MM_DEVINL void load_row_bf16x8(const __nv_bfloat16* row, int lane, float (&f)[N * 8]) {
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * (lane + 32 * t));
        f[8 * t + 0] = bf16lo(u.x); f[8 * t + 1] = bf16hi(u.x); f[8 * t + 2] = bf16lo(u.y); f[8 * t + 3] = bf16hi(u.y);
        f[8 * t + 4] = bf16lo(u.z); f[8 * t + 5] = bf16hi(u.z); f[8 * t + 6] = bf16lo(u.w); f[8 * t + 7] = bf16hi(u.w);
    }
}
template <int N>
MM_DEVINL void load_row_bf16x4(const __nv_bfloat16* row, int lane, float (&f)[N * 4]) {
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const uint2 u = *reinterpret_cast<const uint2*>(row + 4 * (lane + 32 * t));
        f[4 * t + 0] = bf16lo(u.x); f[4 * t + 1] = bf16hi(u.x); f[4 * t + 2] = bf16lo(u.y); f[4 * t + 3] = bf16hi(u.y);
    }
}
template <int N>
MM_DEVINL void load_row_f32x4(const float* row, int lane, float (&f)[N * 4]) {
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const float4 u = *reinterpret_cast<const float4*>(row + 4 * (lane + 32 * t));
        f[4 * t + 0] = u.x; f[4 * t + 1] = u.y; f[4 * t + 2] = u.z; f[4 * t + 3] = u.w;
    }
}
// One [D] row of the output-typed tensors, in the x8 lane layout.
template <int N, typename T>
MM_DEVINL void load_row_x8(const T* row, int lane, float (&f)[N * 8]) {
    if constexpr (sizeof(T) == 2) {
        load_row_bf16x8<N>(reinterpret_cast<const __nv_bfloat16*>(row), lane, f);
    } else {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const float4 a = *reinterpret_cast<const float4*>(row + 8 * (lane + 32 * t));
            const float4 b = *reinterpret_cast<const float4*>(row + 8 * (lane + 32 * t) + 4);
            f[8 * t + 0] = a.x; f[8 * t + 1] = a.y; f[8 * t + 2] = a.z; f[8 * t + 3] = a.w;
            f[8 * t + 4] = b.x; f[8 * t + 5] = b.y; f[8 * t + 6] = b.z; f[8 * t + 7] = b.w;
        }
    }
}
template <int N, typename T>
MM_DEVINL void store_row_x8(T* row, int lane, const float (&f)[N * 8]) {
    if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int t = 0; t < N; ++t)
            stg_v4(row + 8 * (lane + 32 * t), make_uint4(pack_bf16x2(f[8 * t], f[8 * t + 1]), pack_bf16x2(f[8 * t + 2], f[8 * t + 3]),
                                                         pack_bf16x2(f[8 * t + 4], f[8 * t + 5]), pack_bf16x2(f[8 * t + 6], f[8 * t + 7])));
    } else {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            *reinterpret_cast<float4*>(row + 8 * (lane + 32 * t)) = make_float4(f[8 * t], f[8 * t + 1], f[8 * t + 2], f[8 * t + 3]);
            *reinterpret_cast<float4*>(row + 8 * (lane + 32 * t) + 4) = make_float4(f[8 * t + 4], f[8 * t + 5], f[8 * t + 6], f[8 * t + 7]);
        }
    }
}

// logit of one (token, scale): w2 . ReLU(lerp(Z rows)) (without b2), full warp reduction.
template <int N>
MM_DEVINL float scale_logit(const __nv_bfloat16* Z, int H, long long r0, long long r1, float lam, const float (&w2)[N * 4], int lane) {
    float z0[N * 4], z1[N * 4];
    load_row_bf16x4<N>(Z + r0 * H, lane, z0);
    if (r1 != r0) load_row_bf16x4<N>(Z + r1 * H, lane, z1);
    float acc = 0.f;
    const float l0 = 1.0f - lam;
#pragma unroll
    for (int i = 0; i < N * 4; ++i) {
        const float h = (r1 != r0) ? (l0 * z0[i] + lam * z1[i]) : z0[i];
        acc = fmaf(fmaxf(h, 0.f), w2[i], acc);
    }
    return warp_sum(acc);
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int D, typename OutT>
__global__ void __launch_bounds__(256)
combine_fwd_kernel(const CombineArgs a) {
    constexpr int N = D / 256;      // 16-byte chunks per lane for a [D] bf16 row; 8-byte chunks for a [D/2] row
    constexpr int H = D / 2;
    __shared__ float s_g[8][D];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p_begin = blockIdx.x * CB_TOKENS_PER_BLOCK + warp * CB_TOKENS_PER_WARP;
    float gsum[N * 8];
#pragma unroll
    for (int i = 0; i < N * 8; ++i) gsum[i] = 0.f;

    for (int tp = 0; tp < CB_TOKENS_PER_WARP; ++tp) {
        const int p = p_begin + tp;
        if (p >= a.P) break;
        float o[N * 8];
#pragma unroll
        for (int i = 0; i < N * 8; ++i) o[i] = 0.f;
        for (int j = 0; j < a.topk; ++j) {
            const int item = b * a.topk + j;
            const int slot = a.inv_perm[item];
            const int e = a.slot_expert[slot];
            const float g = a.gate ? a.gate[item] : 1.0f;
            float w2[N * 4];
            load_row_f32x4<N>(a.w2 + static_cast<size_t>(e) * H, lane, w2);
            const float b2 = a.b2[e];
            float logit[4];
            long long r0[4], r1[4];
            float lam[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                const long long base = a.slot_row[s * a.n_items + slot];
                r0[s] = base + L.i0; r1[s] = base + L.i1; lam[s] = L.lam;
                if (L.lam == 0.f) r1[s] = r0[s];
                logit[s] = scale_logit<N>(a.Z, H, r0[s], r1[s], lam[s], w2, lane) + b2;
            }
            const float mx = fmaxf(fmaxf(logit[0], logit[1]), fmaxf(logit[2], logit[3]));
            float ex[4], sum = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) { ex[s] = expf(logit[s] - mx); sum += ex[s]; }
            const float inv = 1.0f / sum;
            float beta[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) beta[s] = ex[s] * inv;
            if (lane == 0)
                *reinterpret_cast<float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4) = make_float4(beta[0], beta[1], beta[2], beta[3]);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                float y0[N * 8];
                load_row_bf16x8<N>(a.Y + r0[s] * D, lane, y0);
                const float c0 = g * beta[s] * (1.0f - lam[s]);
                if (r1[s] != r0[s]) {
                    float y1[N * 8];
                    load_row_bf16x8<N>(a.Y + r1[s] * D, lane, y1);
                    const float c1 = g * beta[s] * lam[s];
#pragma unroll
                    for (int i = 0; i < N * 8; ++i) o[i] = fmaf(c0, y0[i], fmaf(c1, y1[i], o[i]));
                } else {
                    const float c = g * beta[s];
#pragma unroll
                    for (int i = 0; i < N * 8; ++i) o[i] = fmaf(c, y0[i], o[i]);
                }
            }
        }
        store_row_x8<N, OutT>(static_cast<OutT*>(a.out) + (static_cast<size_t>(b) * a.P + p) * D, lane, o);
#pragma unroll
        for (int i = 0; i < N * 8; ++i) gsum[i] += o[i];
    }
    // deterministic per-block partial of the global mean
#pragma unroll
    for (int t = 0; t < N; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) s_g[warp][8 * (lane + 32 * t) + i] = gsum[8 * t + i];
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_g[w][d];
        a.gpart[(static_cast<size_t>(b) * a.nblk + blockIdx.x) * D + d] = acc;
    }
}

// global_feat[b, d] = (sum over blocks of gpart) / P
__global__ void __launch_bounds__(256) global_mean_kernel(const float* __restrict__ gpart, int nblk, int D, float inv_p, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float acc = 0.f;
    for (int k = 0; k < nblk; ++k) acc += gpart[(static_cast<size_t>(b) * nblk + k) * D + d];
    out[static_cast<size_t>(b) * D + d] = acc * inv_p;
}

// dF(p) of image b in the x8 lane layout: dlocal[b, p, :] + dglobal[b, :] / P
template <int N, int D, typename OutT>
MM_DEVINL void load_dfused(const CombineArgs& a, int b, int p, const float (&dg)[N * 8], int lane, float (&df)[N * 8]) {
    if (a.dlocal) {
        load_row_x8<N, OutT>(static_cast<const OutT*>(a.dlocal) + (static_cast<size_t>(b) * a.P + p) * D, lane, df);
#pragma unroll
        for (int i = 0; i < N * 8; ++i) df[i] += dg[i];
    } else {
#pragma unroll
        for (int i = 0; i < N * 8; ++i) df[i] = dg[i];
    }
}
template <int N, int D>
MM_DEVINL void load_dglobal(const CombineArgs& a, int b, int lane, float (&dg)[N * 8]) {
    if (a.dglobal) {
        const float inv_p = 1.0f / static_cast<float>(a.P);
        load_row_x8<N, float>(a.dglobal + static_cast<size_t>(b) * D, lane, dg);
#pragma unroll
        for (int i = 0; i < N * 8; ++i) dg[i] *= inv_p;
    } else {
#pragma unroll
        for (int i = 0; i < N * 8; ++i) dg[i] = 0.f;
    }
}

// ------------------------------------------------------------------------------------
// backward pass A: dlogit (and the gate gradient for the top-k extension)
// ------------------------------------------------------------------------------------
template <int D, typename OutT>
__global__ void __launch_bounds__(256)
combine_bwd_logit_kernel(const CombineArgs a) {
    constexpr int N = D / 256;
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p_begin = blockIdx.x * CB_TOKENS_PER_BLOCK + warp * CB_TOKENS_PER_WARP;
    float dg[N * 8];
    load_dglobal<N, D>(a, b, lane, dg);
    for (int j = 0; j < a.topk; ++j) {
        const int item = b * a.topk + j;
        const int slot = a.inv_perm[item];
        const float g = a.gate ? a.gate[item] : 1.0f;
        float dgate_acc = 0.f;
        for (int tp = 0; tp < CB_TOKENS_PER_WARP; ++tp) {
            const int p = p_begin + tp;
            if (p >= a.P) break;
            float df[N * 8];
            load_dfused<N, D, OutT>(a, b, p, dg, lane, df);
            float dbeta[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                const long long base = a.slot_row[s * a.n_items + slot];
                float y0[N * 8];
                load_row_bf16x8<N>(a.Y + (base + L.i0) * D, lane, y0);
                float acc = 0.f;
                if (L.lam != 0.f && L.i1 != L.i0) {
                    float y1[N * 8];
                    load_row_bf16x8<N>(a.Y + (base + L.i1) * D, lane, y1);
                    const float l0 = 1.0f - L.lam;
#pragma unroll
                    for (int i = 0; i < N * 8; ++i) acc = fmaf(df[i], l0 * y0[i] + L.lam * y1[i], acc);
                } else {
#pragma unroll
                    for (int i = 0; i < N * 8; ++i) acc = fmaf(df[i], y0[i], acc);
                }
                dbeta[s] = warp_sum(acc);
            }
            const float4 bt = *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + p) * 4);
            const float dot = bt.x * dbeta[0] + bt.y * dbeta[1] + bt.z * dbeta[2] + bt.w * dbeta[3];
            dgate_acc += dot;
            if (lane == 0) {
                float4 dl;
                dl.x = g * bt.x * (dbeta[0] - dot); dl.y = g * bt.y * (dbeta[1] - dot);
                dl.z = g * bt.z * (dbeta[2] - dot); dl.w = g * bt.w * (dbeta[3] - dot);
                *reinterpret_cast<float4*>(a.dlogit + (static_cast<size_t>(slot) * a.P + p) * 4) = dl;
            }
        }
        if (a.dgate && lane == 0) atomicAdd(a.dgate + item, dgate_acc);
    }
}

// ------------------------------------------------------------------------------------
// backward pass B: transposed lerp as a gather over each native row's token window
// ------------------------------------------------------------------------------------
template <int D, typename OutT>
__global__ void __launch_bounds__(256)
combine_bwd_rows_kernel(const CombineArgs a) {
    constexpr int N = D / 256;
    constexpr int H = D / 2;
    constexpr int PART = 2 * H + 1;
    __shared__ float s_part[8][PART];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (blockIdx.y >= a.n_items) {
        // zero the 128-row padding behind every expert segment of dZ (wgrad reduces over whole tiles)
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
    const int item = a.perm[slot];
    const int b = item / a.topk;
    const int e = a.slot_expert[slot];
    const float g = a.gate ? a.gate[item] : 1.0f;
    float w2[N * 4];
    load_row_f32x4<N>(a.w2 + static_cast<size_t>(e) * H, lane, w2);
    float dg[N * 8];
    load_dglobal<N, D>(a, b, lane, dg);

    float dw2_acc[N * 4], db1_acc[N * 4], db2_acc = 0.f;
#pragma unroll
    for (int i = 0; i < N * 4; ++i) { dw2_acc[i] = 0.f; db1_acc[i] = 0.f; }

    const int total_rows = a.Ps[0] + a.Ps[1] + a.Ps[2] + a.Ps[3];
    const int u_begin = blockIdx.x * CB_ROWS_PER_BLOCK + warp * CB_ROWS_PER_WARP;
    for (int tr = 0; tr < CB_ROWS_PER_WARP; ++tr) {
        int u = u_begin + tr;
        if (u >= total_rows) break;
        int s = 0;
        while (u >= a.Ps[s]) { u -= a.Ps[s]; ++s; }
        const int i = u;
        const int Ps = a.Ps[s];
        const float scale = a.scale[s];
        const long long base = a.slot_row[s * a.n_items + slot];
        // token window of native row i: tokens whose i0 is i-1 or i (plus a one-token safety margin)
        const float inv_scale = static_cast<float>(a.P) / static_cast<float>(Ps);
        int p_lo = static_cast<int>(floorf((static_cast<float>(i) - 0.5f) * inv_scale - 0.5f)) - 1;
        int p_hi = static_cast<int>(ceilf((static_cast<float>(i) + 1.5f) * inv_scale - 0.5f)) + 1;
        if (i == 0) p_lo = 0;
        if (i == Ps - 1) p_hi = a.P;
        p_lo = max(p_lo, 0);
        p_hi = min(p_hi, a.P);

        float zc[N * 4], zm[N * 4], zp[N * 4];
        load_row_bf16x4<N>(a.Z + (base + i) * H, lane, zc);
        load_row_bf16x4<N>(a.Z + (base + max(i - 1, 0)) * H, lane, zm);
        load_row_bf16x4<N>(a.Z + (base + min(i + 1, Ps - 1)) * H, lane, zp);

        float acc_u[N * 8], acc_z[N * 4];
#pragma unroll
        for (int k = 0; k < N * 8; ++k) acc_u[k] = 0.f;
#pragma unroll
        for (int k = 0; k < N * 4; ++k) acc_z[k] = 0.f;

        for (int p = p_lo; p < p_hi; ++p) {
            const LerpSrc L = lerp_src(p, scale, Ps);
            const float l0 = 1.0f - L.lam;
            float w = 0.f;
            if (L.i0 == i) w += l0;
            if (L.i1 == i) w += L.lam;
            if (L.i0 != i && L.i1 != i) continue;
            const size_t tok = static_cast<size_t>(slot) * a.P + p;
            const float bt = a.beta[tok * 4 + s];
            const float dl = a.dlogit[tok * 4 + s];
            if (w != 0.f) {
                float df[N * 8];
                load_dfused<N, D, OutT>(a, b, p, dg, lane, df);
                const float c = w * bt * g;
#pragma unroll
                for (int k = 0; k < N * 8; ++k) acc_u[k] = fmaf(c, df[k], acc_u[k]);
            }
            // interp(Z)(p) from the three cached native rows
            const bool lo_is_i = (L.i0 == i);
            const float cz = w * dl;
#pragma unroll
            for (int k = 0; k < N * 4; ++k) {
                const float za = lo_is_i ? zc[k] : zm[k];
                const float zb = lo_is_i ? (L.i1 == i ? zc[k] : zp[k]) : zc[k];
                const float h = (L.i1 != L.i0 && L.lam != 0.f) ? (l0 * za + L.lam * zb) : za;
                if (h > 0.f) {
                    acc_z[k] = fmaf(cz, w2[k], acc_z[k]);
                    if (lo_is_i) dw2_acc[k] = fmaf(dl, h, dw2_acc[k]);   // each (token, scale) counted once
                }
            }
            if (lo_is_i) db2_acc += dl;
        }
        // write dUT (bf16) and dZ (bf16); db1 = column sums of dZ
        store_row_x8<N, __nv_bfloat16>(a.dUT + (base + i) * D, lane, acc_u);
#pragma unroll
        for (int t = 0; t < N; ++t) {
            *reinterpret_cast<uint2*>(a.dZ + (base + i) * H + 4 * (lane + 32 * t)) =
                make_uint2(pack_bf16x2(acc_z[4 * t], acc_z[4 * t + 1]), pack_bf16x2(acc_z[4 * t + 2], acc_z[4 * t + 3]));
        }
#pragma unroll
        for (int k = 0; k < N * 4; ++k) db1_acc[k] += acc_z[k];
    }

    // per-CTA partials: [dw2 (H) | db1 (H) | db2 (1)]
#pragma unroll
    for (int t = 0; t < N; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s_part[warp][4 * (lane + 32 * t) + k] = dw2_acc[4 * t + k];
            s_part[warp][H + 4 * (lane + 32 * t) + k] = db1_acc[4 * t + k];
        }
    if (lane == 0) s_part[warp][2 * H] = db2_acc;
    __syncthreads();
    float* dst = a.part + (static_cast<size_t>(slot) * a.nrb + blockIdx.x) * PART;
    for (int c = threadIdx.x; c < PART; c += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_part[w][c];
        dst[c] = acc;
    }
}

// ------------------------------------------------------------------------------------
// Token-centric backward (integer scale ratios r_s = P / P_s).  Each warp owns a run of 32
// consecutive tokens of one slot and streams over them once:
//   * r = 1      (IDENT)  : the transposed lerp is the identity;
//   * r <= 32    (DIRECT) : the warp also owns the native rows i with r*i inside its run; their
//                           token windows reach r/2 tokens outside the run (halo), contributions of
//                           halo tokens to rows owned by the neighbour are dropped, so no atomics;
//   * r/2 % 32 == 0 (MOMENT, the coarsest scale): inside a run the lerp weight of a native row is
//                           affine in the token index, so the run only emits M0 = sum g(p) and
//                           M1 = sum (p - p0) g(p); a finalize kernel combines the <= 4 runs of a
//                           row's window:  dRow = sum_runs w(p0) M0 + slope M1.
// g(p) = beta_s(p) dF(p) for dUT and g(p) = dlogit_s(p) w2 [interp(Z_s)(p) > 0] for dZ.
// ------------------------------------------------------------------------------------
template <int NE, typename T>
MM_DEVINL void load_slab(const T* p, int lane, float (&f)[NE * 4]) {
    if constexpr (sizeof(T) == 2) load_row_bf16x4<NE>(reinterpret_cast<const __nv_bfloat16*>(p), lane, f);
    else load_row_f32x4<NE>(reinterpret_cast<const float*>(p), lane, f);
}
template <int NE>
MM_DEVINL void store_slab_bf16(__nv_bfloat16* p, int lane, const float (&f)[NE * 4]) {
#pragma unroll
    for (int t = 0; t < NE; ++t)
        *reinterpret_cast<uint2*>(p + 4 * (lane + 32 * t)) =
            make_uint2(pack_bf16x2(f[4 * t], f[4 * t + 1]), pack_bf16x2(f[4 * t + 2], f[4 * t + 3]));
}
template <int NE>
MM_DEVINL void store_slab_f32(float* p, int lane, const float (&f)[NE * 4]) {
#pragma unroll
    for (int t = 0; t < NE; ++t)
        *reinterpret_cast<float4*>(p + 4 * (lane + 32 * t)) = make_float4(f[4 * t], f[4 * t + 1], f[4 * t + 2], f[4 * t + 3]);
}

// Row cache: the two native rows (i0, i1) a token interpolates between change only every
// r = P / P_s tokens, so a warp that walks consecutive tokens keeps them in registers.
template <int NE>
struct RowPair {
    float a[NE * 4], b[NE * 4];
    int row_a = -1, row_b = -1;
};
// make `c.a` / `c.b` hold rows L.i0 / L.i1 of the [rows, ld] bf16 matrix (column offset already applied)
template <int NE>
MM_DEVINL bool row_pair_update(RowPair<NE>& c, const __nv_bfloat16* mat, long long base, long long ld, const LerpSrc& L, int lane) {
    if (L.i0 != c.row_a) {
        if (L.i0 == c.row_b) {
#pragma unroll
            for (int k = 0; k < NE * 4; ++k) c.a[k] = c.b[k];
        } else {
            load_row_bf16x4<NE>(mat + (base + L.i0) * ld, lane, c.a);
        }
        c.row_a = L.i0;
    }
    const bool two = (L.i1 != L.i0) && (L.lam != 0.f);
    if (two && L.i1 != c.row_b) {
        load_row_bf16x4<NE>(mat + (base + L.i1) * ld, lane, c.b);
        c.row_b = L.i1;
    }
    return two;
}

constexpr int OUT_RUNS_PER_BLOCK = 4;   // run-based backward kernels: 8 warps = 4 runs x 2 column halves

}  // namespace mm
#include "combine_staged.cuh"
namespace mm {

// Per-lane token scalars for the token-centric backward: lane t keeps the value of tokens
// p_lo + t and p_lo + 32 + t (a run plus its halo spans at most 64 tokens); the loop reads
// them back with one shuffle instead of a dependent global load per token.
MM_DEVINL float lane_pick(float va, float vb, int idx) {
    const float x = __shfl_sync(0xffffffffu, va, idx & 31);
    const float y = __shfl_sync(0xffffffffu, vb, idx & 31);
    return idx < 32 ? x : y;
}

// backward pass A, run-based: partial dbeta_s = <dF, interp(Y_s)> over one column half.
// grid = (ceil(nruns / 4), B); warp = (run, column half).  dbeta [n_items, P, 2, 4].
template <int D, typename OutT>
__global__ void __launch_bounds__(256, MM_COMBINE_MINBLOCKS)
combine_bwd_dbeta_kernel(const CombineArgs a) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * OUT_RUNS_PER_BLOCK + (warp >> 1);
    const int half = warp & 1;
    const int col0 = half * H;
    const int b = blockIdx.y;
    const int t0 = c * RUN_TOKENS;
    if (c >= a.nruns) return;
    const int n_tok = min(RUN_TOKENS, a.P - t0);
    float dg[E];
    if (a.dglobal) {
        load_slab<NE, float>(a.dglobal + static_cast<size_t>(b) * D + col0, lane, dg);
        const float inv_p = 1.0f / static_cast<float>(a.P);
#pragma unroll
        for (int k = 0; k < E; ++k) dg[k] *= inv_p;
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) dg[k] = 0.f;
    }
    const OutT* dl_base = a.dlocal ? static_cast<const OutT*>(a.dlocal) + (static_cast<size_t>(b) * a.P + t0) * D + col0 : nullptr;
    for (int jk = 0; jk < a.topk; ++jk) {
        const int item = b * a.topk + jk;
        const int slot = a.inv_perm[item];
        long long base[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) base[s] = a.slot_row[s * a.n_items + slot];
        RowPair<NE> y1, y2, y3;
        float y0[E], df[E];
        load_row_bf16x4<NE>(a.Y + (base[0] + t0) * D + col0, lane, y0);
        if (dl_base) {
            load_slab<NE, OutT>(dl_base, lane, df);
#pragma unroll
            for (int k = 0; k < E; ++k) df[k] += dg[k];
        } else {
#pragma unroll
            for (int k = 0; k < E; ++k) df[k] = dg[k];
        }
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < n_tok; ++t) {
            const int p = t0 + t;
            float yn[E], dfn[E];
#pragma unroll
            for (int k = 0; k < E; ++k) { yn[k] = 0.f; dfn[k] = dg[k]; }
            if (t + 1 < n_tok) {
                load_row_bf16x4<NE>(a.Y + (base[0] + p + 1) * D + col0, lane, yn);
                if (dl_base) {
                    load_slab<NE, OutT>(dl_base + static_cast<size_t>(t + 1) * D, lane, dfn);
#pragma unroll
                    for (int k = 0; k < E; ++k) dfn[k] += dg[k];
                }
            }
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int k = 0; k < E; ++k) d0 = fmaf(df[k], y0[k], d0);
            {
                const LerpSrc L = lerp_src(p, a.scale[1], a.Ps[1]);
                const bool two = row_pair_update<NE>(y1, a.Y + col0, base[1], D, L, lane);
                const float l0 = 1.0f - L.lam;
#pragma unroll
                for (int k = 0; k < E; ++k) d1 = fmaf(df[k], two ? (l0 * y1.a[k] + L.lam * y1.b[k]) : y1.a[k], d1);
            }
            {
                const LerpSrc L = lerp_src(p, a.scale[2], a.Ps[2]);
                const bool two = row_pair_update<NE>(y2, a.Y + col0, base[2], D, L, lane);
                const float l0 = 1.0f - L.lam;
#pragma unroll
                for (int k = 0; k < E; ++k) d2 = fmaf(df[k], two ? (l0 * y2.a[k] + L.lam * y2.b[k]) : y2.a[k], d2);
            }
            {
                const LerpSrc L = lerp_src(p, a.scale[3], a.Ps[3]);
                const bool two = row_pair_update<NE>(y3, a.Y + col0, base[3], D, L, lane);
                const float l0 = 1.0f - L.lam;
#pragma unroll
                for (int k = 0; k < E; ++k) d3 = fmaf(df[k], two ? (l0 * y3.a[k] + L.lam * y3.b[k]) : y3.a[k], d3);
            }
            d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3);
            if (lane == t) mine = make_float4(d0, d1, d2, d3);
#pragma unroll
            for (int k = 0; k < E; ++k) { y0[k] = yn[k]; df[k] = dfn[k]; }
        }
        if (lane < n_tok)
            *reinterpret_cast<float4*>(a.dlogit + ((static_cast<size_t>(slot) * a.P + t0 + lane) * 2 + half) * 4) = mine;
    }
}

// dlogit of one token from beta and the two dbeta halves (softmax-over-scales backward), scaled by the gate.
MM_DEVINL float4 token_dlogit(const CombineArgs& a, int slot, int p, float g, float& dot_out) {
    dot_out = 0.f;
    if (p < 0 || p >= a.P) return make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t tok = static_cast<size_t>(slot) * a.P + p;
    if (a.row_dot || a.dbeta_loc) {   // tensor-core / rank-1 path: dbeta = local part (GEMM) + interp(G_s)(p) (dF's per-image constant)
        const float4 bt = *reinterpret_cast<const float4*>(a.beta + tok * 4);
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.dbeta_loc) {
            const float4 dl = *reinterpret_cast<const float4*>(a.dbeta_loc + tok * 4);
            d[0] = dl.x; d[1] = dl.y; d[2] = dl.z; d[3] = dl.w;
        }
        if (a.row_dot) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                const float* G = a.row_dot + a.slot_row[s * a.n_items + slot];
                d[s] += (1.0f - L.lam) * G[L.i0] + L.lam * G[L.i1];
            }
        }
        const float dot = bt.x * d[0] + bt.y * d[1] + bt.z * d[2] + bt.w * d[3];
        dot_out = dot;
        return make_float4(g * bt.x * (d[0] - dot), g * bt.y * (d[1] - dot), g * bt.z * (d[2] - dot), g * bt.w * (d[3] - dot));
    }
    if (!a.dlogit_is_halves)     // finished dlogit (gate already applied) written by combine_bwd_logit_staged_kernel
        return *reinterpret_cast<const float4*>(a.dlogit + tok * 4);
    const float4 bt = *reinterpret_cast<const float4*>(a.beta + tok * 4);
    const float4 h0 = *reinterpret_cast<const float4*>(a.dlogit + tok * 8);
    const float4 h1 = *reinterpret_cast<const float4*>(a.dlogit + tok * 8 + 4);
    const float d0 = h0.x + h1.x, d1 = h0.y + h1.y, d2 = h0.z + h1.z, d3 = h0.w + h1.w;
    const float dot = bt.x * d0 + bt.y * d1 + bt.z * d2 + bt.w * d3;
    dot_out = dot;
    return make_float4(g * bt.x * (d0 - dot), g * bt.y * (d1 - dot), g * bt.z * (d2 - dot), g * bt.w * (d3 - dot));
}

// dUT: grid = (ceil(2 * nruns / 8), n_items); warp = (run, column half).
template <int D, typename OutT>
__global__ void __launch_bounds__(256, MM_COMBINE_MINBLOCKS)
combine_bwd_u_kernel(const CombineArgs a) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * 8 + warp;
    const int c = unit >> 1, j = unit & 1;
    if (c >= a.nruns) return;
    const int slot = blockIdx.y;
    const int item = a.perm[slot];
    const int b = item / a.topk;
    const float g = a.gate ? a.gate[item] : 1.0f;
    const int t0 = c * RUN_TOKENS;
    const int col0 = j * H;

    float dg[E];
    if (a.dglobal) {
        load_slab<NE, float>(a.dglobal + static_cast<size_t>(b) * D + col0, lane, dg);
        const float inv_p = 1.0f / static_cast<float>(a.P);
#pragma unroll
        for (int k = 0; k < E; ++k) dg[k] *= inv_p;
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) dg[k] = 0.f;
    }
    float acc[3][2][E];
    int cur[3] = {-1, -1, -1};
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int k = 0; k < E; ++k) { acc[s][0][k] = 0.f; acc[s][1][k] = 0.f; }

    long long base[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) base[s] = a.slot_row[s * a.n_items + slot];

    const int p_lo = t0 - a.halo, p_hi = min(a.P, t0 + RUN_TOKENS + a.halo);
    const int p_first = max(0, p_lo);
    // beta (x gate) of tokens p_lo + lane and p_lo + 32 + lane
    float4 bA = make_float4(0.f, 0.f, 0.f, 0.f), bB = bA;
    {
        const int pa = p_lo + lane, pb = p_lo + 32 + lane;
        if (pa >= 0 && pa < a.P) bA = *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + pa) * 4);
        if (pb >= 0 && pb < a.P) bB = *reinterpret_cast<const float4*>(a.beta + (static_cast<size_t>(slot) * a.P + pb) * 4);
        bA.x *= g; bA.y *= g; bA.z *= g; bA.w *= g;
        bB.x *= g; bB.y *= g; bB.z *= g; bB.w *= g;
    }
    const OutT* dl_base = a.dlocal ? static_cast<const OutT*>(a.dlocal) + static_cast<size_t>(b) * a.P * D + col0 : nullptr;
    float df[E];
    if (dl_base) {
        load_slab<NE, OutT>(dl_base + static_cast<size_t>(p_first) * D, lane, df);
#pragma unroll
        for (int k = 0; k < E; ++k) df[k] += dg[k];
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) df[k] = dg[k];
    }
    for (int p = p_first; p < p_hi; ++p) {
        const bool in_run = (p >= t0) && (p < t0 + RUN_TOKENS);
        float dfn[E];
#pragma unroll
        for (int k = 0; k < E; ++k) dfn[k] = dg[k];
        if (dl_base && p + 1 < p_hi) {
            load_slab<NE, OutT>(dl_base + static_cast<size_t>(p + 1) * D, lane, dfn);
#pragma unroll
            for (int k = 0; k < E; ++k) dfn[k] += dg[k];
        }
        const int idx = p - p_lo;
        const float bt[4] = {lane_pick(bA.x, bB.x, idx), lane_pick(bA.y, bB.y, idx), lane_pick(bA.z, bB.z, idx),
                             lane_pick(bA.w, bB.w, idx)};
        if (in_run) {   // scale 0: identity
            float o[E];
#pragma unroll
            for (int k = 0; k < E; ++k) o[k] = bt[0] * df[k];
            store_slab_bf16<NE>(a.dUT + (base[0] + p) * D + col0, lane, o);
        }
#pragma unroll
        for (int s = 1; s < 4; ++s) {
            const int r = a.ratio[s];
            if (a.mode[s] == SCALE_DIRECT) {
                const int hs = r >> 1;
                if (p < t0 - hs || p >= t0 + RUN_TOKENS + hs) continue;
                const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                if (L.i0 != cur[s - 1]) {
                    const int done = cur[s - 1];
                    if (done >= 0 && done * r >= t0 && done * r < t0 + RUN_TOKENS)
                        store_slab_bf16<NE>(a.dUT + (base[s] + done) * D + col0, lane, acc[s - 1][0]);
#pragma unroll
                    for (int k = 0; k < E; ++k) { acc[s - 1][0][k] = acc[s - 1][1][k]; acc[s - 1][1][k] = 0.f; }
                    cur[s - 1] = L.i0;
                }
                const float c0 = bt[s] * (1.0f - L.lam), c1 = bt[s] * L.lam;
                if (L.i1 == L.i0) {
#pragma unroll
                    for (int k = 0; k < E; ++k) acc[s - 1][0][k] = fmaf(c0 + c1, df[k], acc[s - 1][0][k]);
                } else {
#pragma unroll
                    for (int k = 0; k < E; ++k) {
                        acc[s - 1][0][k] = fmaf(c0, df[k], acc[s - 1][0][k]);
                        acc[s - 1][1][k] = fmaf(c1, df[k], acc[s - 1][1][k]);
                    }
                }
            } else if (a.mode[s] == SCALE_MOMENT) {
                if (!in_run) continue;
                const float c0 = bt[s], c1 = bt[s] * static_cast<float>(p - t0);
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    acc[s - 1][0][k] = fmaf(c0, df[k], acc[s - 1][0][k]);
                    acc[s - 1][1][k] = fmaf(c1, df[k], acc[s - 1][1][k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < E; ++k) df[k] = dfn[k];
    }
#pragma unroll
    for (int s = 1; s < 4; ++s) {
        const int r = a.ratio[s];
        if (a.mode[s] == SCALE_DIRECT) {
            const int i = cur[s - 1];
            if (i >= 0 && i * r >= t0 && i * r < t0 + RUN_TOKENS)
                store_slab_bf16<NE>(a.dUT + (base[s] + i) * D + col0, lane, acc[s - 1][0]);
            if (i >= 0 && i + 1 < a.Ps[s] && (i + 1) * r >= t0 && (i + 1) * r < t0 + RUN_TOKENS)
                store_slab_bf16<NE>(a.dUT + (base[s] + i + 1) * D + col0, lane, acc[s - 1][1]);
        } else if (a.mode[s] == SCALE_MOMENT) {
            float* m = a.mom_u + ((static_cast<size_t>(slot) * a.nruns + c) * 2) * D + col0;
            store_slab_f32<NE>(m, lane, acc[s - 1][0]);
            store_slab_f32<NE>(m + D, lane, acc[s - 1][1]);
        }
    }
}

// dZ + the per-token parameter gradients (dw2, db1, db2): grid = (ceil(nruns / 8), n_items + K); warp = run.
template <int D>
__global__ void __launch_bounds__(256, MM_COMBINE_MINBLOCKS)
combine_bwd_z_kernel(const CombineArgs a) {
    constexpr int NE = D / 256;
    constexpr int E = NE * 4;
    constexpr int H = D / 2;
    constexpr int PART = 2 * H + 1;
    __shared__ float s_part[RUNS_PER_BLOCK][PART];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (blockIdx.y >= a.n_items) {   // zero the 128-row padding behind every expert segment of dZ
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

    const int c = blockIdx.x * RUNS_PER_BLOCK + warp;
    const int slot = blockIdx.y;
    const int item = a.perm[slot];
    const int e = a.slot_expert[slot];
    const float g = a.gate ? a.gate[item] : 1.0f;
    const int t0 = c * RUN_TOKENS;
    float dw2[E], db1[E], db2 = 0.f;
#pragma unroll
    for (int k = 0; k < E; ++k) { dw2[k] = 0.f; db1[k] = 0.f; }

    if (c < a.nruns) {
        float w2[E];
        load_row_f32x4<NE>(a.w2 + static_cast<size_t>(e) * H, lane, w2);
        // dlogit of tokens (t0 - halo) + lane and + 32 + lane; the gate gradient falls out of the same dot product
        const int p_base = t0 - a.halo;
        float dotA, dotB;
        const float4 dlA = token_dlogit(a, slot, p_base + lane, g, dotA);
        const float4 dlB = token_dlogit(a, slot, p_base + 32 + lane, g, dotB);
        if (a.dgate && a.dlogit_is_halves) {
            const int pa = p_base + lane, pb = p_base + 32 + lane;
            float dsum = ((pa >= t0 && pa < t0 + RUN_TOKENS) ? dotA : 0.f) + ((pb >= t0 && pb < t0 + RUN_TOKENS) ? dotB : 0.f);
            dsum = warp_sum(dsum);
            if (lane == 0) atomicAdd(a.dgate + item, dsum);
        }
        for (int s = 0; s < 4; ++s) {
            const int mode = a.mode[s];
            const int r = a.ratio[s];
            const long long base = a.slot_row[s * a.n_items + slot];
            const int hs = (mode == SCALE_DIRECT) ? (r >> 1) : 0;
            const int p_lo = max(0, t0 - hs), p_hi = min(a.P, t0 + RUN_TOKENS + hs);
            const float cA = s == 0 ? dlA.x : (s == 1 ? dlA.y : (s == 2 ? dlA.z : dlA.w));
            const float cB = s == 0 ? dlB.x : (s == 1 ? dlB.y : (s == 2 ? dlB.z : dlB.w));
            float lo[E], hi[E];
            RowPair<NE> z;
            int cur = -1;
#pragma unroll
            for (int k = 0; k < E; ++k) { lo[k] = 0.f; hi[k] = 0.f; }
            if (mode == SCALE_IDENT) {   // every token has its own row: keep the next row's load in flight
                load_row_bf16x4<NE>(a.Z + (base + p_lo) * H, lane, z.a);
            }
            for (int p = p_lo; p < p_hi; ++p) {
                const bool in_run = (p >= t0) && (p < t0 + RUN_TOKENS);
                const float dl = lane_pick(cA, cB, p - p_base);
                const LerpSrc L = lerp_src(p, a.scale[s], a.Ps[s]);
                bool two = false;
                float zn[E];
                if (mode == SCALE_IDENT) {
#pragma unroll
                    for (int k = 0; k < E; ++k) zn[k] = 0.f;
                    if (p + 1 < p_hi) load_row_bf16x4<NE>(a.Z + (base + p + 1) * H, lane, zn);
                } else {
                    two = row_pair_update<NE>(z, a.Z, base, H, L, lane);
                }
                const float l0 = 1.0f - L.lam;
                float gk[E];
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    const float h = two ? (l0 * z.a[k] + L.lam * z.b[k]) : z.a[k];
                    gk[k] = h > 0.f ? dl * w2[k] : 0.f;
                    if (in_run && h > 0.f) dw2[k] = fmaf(dl, h, dw2[k]);
                }
                if (in_run) {
                    db2 += dl;
#pragma unroll
                    for (int k = 0; k < E; ++k) db1[k] += gk[k];
                }
                if (mode == SCALE_IDENT) {
                    if (in_run) store_slab_bf16<NE>(a.dZ + (base + p) * H, lane, gk);
#pragma unroll
                    for (int k = 0; k < E; ++k) z.a[k] = zn[k];
                } else if (mode == SCALE_DIRECT) {
                    if (L.i0 != cur) {
                        if (cur >= 0 && cur * r >= t0 && cur * r < t0 + RUN_TOKENS)
                            store_slab_bf16<NE>(a.dZ + (base + cur) * H, lane, lo);
#pragma unroll
                        for (int k = 0; k < E; ++k) { lo[k] = hi[k]; hi[k] = 0.f; }
                        cur = L.i0;
                    }
                    if (L.i1 == L.i0) {
#pragma unroll
                        for (int k = 0; k < E; ++k) lo[k] += gk[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < E; ++k) { lo[k] = fmaf(l0, gk[k], lo[k]); hi[k] = fmaf(L.lam, gk[k], hi[k]); }
                    }
                } else {   // SCALE_MOMENT
                    const float t = static_cast<float>(p - t0);
#pragma unroll
                    for (int k = 0; k < E; ++k) { lo[k] += gk[k]; hi[k] = fmaf(t, gk[k], hi[k]); }
                }
            }
            if (mode == SCALE_DIRECT) {
                if (cur >= 0 && cur * r >= t0 && cur * r < t0 + RUN_TOKENS)
                    store_slab_bf16<NE>(a.dZ + (base + cur) * H, lane, lo);
                if (cur >= 0 && cur + 1 < a.Ps[s] && (cur + 1) * r >= t0 && (cur + 1) * r < t0 + RUN_TOKENS)
                    store_slab_bf16<NE>(a.dZ + (base + cur + 1) * H, lane, hi);
            } else if (mode == SCALE_MOMENT) {
                float* m = a.mom_z + ((static_cast<size_t>(slot) * a.nruns + c) * 2) * H;
                store_slab_f32<NE>(m, lane, lo);
                store_slab_f32<NE>(m + H, lane, hi);
            }
        }
    }
    // per-CTA partials: [dw2 (H) | db1 (H) | db2 (1)]
#pragma unroll
    for (int t = 0; t < NE; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s_part[warp][4 * (lane + 32 * t) + k] = dw2[4 * t + k];
            s_part[warp][H + 4 * (lane + 32 * t) + k] = db1[4 * t + k];
        }
    if (lane == 0) s_part[warp][2 * H] = db2;
    __syncthreads();
    float* dst = a.part + (static_cast<size_t>(slot) * a.nrb + blockIdx.x) * PART;
    for (int cc = threadIdx.x; cc < PART; cc += blockDim.x) {
        float accv = 0.f;
#pragma unroll
        for (int w = 0; w < RUNS_PER_BLOCK; ++w) accv += s_part[w][cc];
        dst[cc] = accv;
    }
}

}  // namespace mm
#include "combine_bwd_z.cuh"
#include "combine_rank1.cuh"
#include "combine_mma.cuh"
namespace mm {

// MOMENT scale: combine the runs of each native row's window.  grid = (ceil(Ps[s] / 8), n_items); warp = native row.
template <int D>
__global__ void __launch_bounds__(256)
combine_bwd_finalize_kernel(const CombineArgs a, int s) {
    constexpr int N = D / 256;
    constexpr int NE = D / 256;
    constexpr int H = D / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    const int Ps = a.Ps[s];
    if (i >= Ps) return;
    const int slot = blockIdx.y;
    const int r = a.ratio[s];
    const long long base = a.slot_row[s * a.n_items + slot];
    float u[N * 8], z[NE * 4];
#pragma unroll
    for (int k = 0; k < N * 8; ++k) u[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NE * 4; ++k) z[k] = 0.f;
    const int p_first = max(0, r * i - (r >> 1)), p_last = min(a.P, r * i + r + (r >> 1));   // [first, last)
    for (int c = p_first / RUN_TOKENS; c * RUN_TOKENS < p_last; ++c) {
        const int pa = c * RUN_TOKENS, pb = min(pa + RUN_TOKENS - 1, a.P - 1);
        float wa, wb;
        {
            const LerpSrc L = lerp_src(pa, a.scale[s], Ps);
            wa = (L.i0 == i ? 1.0f - L.lam : 0.f) + (L.i1 == i ? L.lam : 0.f);
            const LerpSrc M = lerp_src(pb, a.scale[s], Ps);
            wb = (M.i0 == i ? 1.0f - M.lam : 0.f) + (M.i1 == i ? M.lam : 0.f);
        }
        if (wa == 0.f && wb == 0.f) continue;
        const float slope = pb > pa ? (wb - wa) / static_cast<float>(pb - pa) : 0.f;
        const float* mu = a.mom_u + ((static_cast<size_t>(slot) * a.nruns + c) * 2) * D;
        const float* mz = a.mom_z + ((static_cast<size_t>(slot) * a.nruns + c) * 2) * H;
        float m0[N * 8], m1[N * 8];
        load_row_x8<N, float>(mu, lane, m0);
        load_row_x8<N, float>(mu + D, lane, m1);
#pragma unroll
        for (int k = 0; k < N * 8; ++k) u[k] = fmaf(wa, m0[k], fmaf(slope, m1[k], u[k]));
        if (!a.z_rows_path) {
            float n0[NE * 4], n1[NE * 4];
            load_row_f32x4<NE>(mz, lane, n0);
            load_row_f32x4<NE>(mz + H, lane, n1);
#pragma unroll
            for (int k = 0; k < NE * 4; ++k) z[k] = fmaf(wa, n0[k], fmaf(slope, n1[k], z[k]));
        }
    }
    store_row_x8<N, __nv_bfloat16>(a.dUT + (base + i) * D, lane, u);
    if (!a.z_rows_path) store_slab_bf16<NE>(a.dZ + (base + i) * H, lane, z);
}

// out[e, c] = sum over the slots of expert e and their nrb blocks of part[slot, blk, c].
// grid = (ceil(C / 32), K); block = 32 columns x 32 row groups (the 100 blocks of cfg2 are latency-bound: the more row groups,
// the shorter each thread's serial chain of loads), fixed summation order (deterministic).
constexpr int ER_GROUPS = 32;
__global__ void __launch_bounds__(32 * ER_GROUPS)
expert_reduce_kernel(const float* __restrict__ part, const int* __restrict__ offsets, int nrb, int C, float* __restrict__ out) {
    __shared__ float sm[ER_GROUPS][33];
    const int e = blockIdx.y;
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const long long lo = static_cast<long long>(offsets[e]) * nrb, hi = static_cast<long long>(offsets[e + 1]) * nrb;
    float acc = 0.f;
    if (c < C) {
#pragma unroll 8
        for (long long r = lo + grp; r < hi; r += ER_GROUPS) acc += part[r * C + c];
    }
    sm[grp][cl] = acc;
    __syncthreads();
    if (grp == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < ER_GROUPS; ++g) t += sm[g][cl];
        out[static_cast<size_t>(e) * C + c] = t;
    }
}

}  // namespace mm

using namespace mm;

static int fill_common(CombineArgs& a, int B, int topk, int P, const int32_t* Ps, int D, const char* who) {
    if (!(B > 0 && topk >= 1 && P > 0)) { mm::set_error("%s: bad shape", who); return MM_ERR_BAD_SHAPE; }
    if (!(D == 256 || D == 512 || D == 768 || D == 1024)) {
        mm::set_error("%s: output_dim must be one of 256/512/768/1024 (got %d)", who, D);
        return MM_ERR_UNSUPPORTED;
    }
    a.B = B; a.topk = topk; a.P = P; a.n_items = B * topk;
    for (int s = 0; s < 4; ++s) {
        a.Ps[s] = Ps[s];
        a.scale[s] = static_cast<float>(Ps[s]) / static_cast<float>(P);
    }
    return MM_OK;
}

#define MM_DISPATCH_D(D, OUT_F32, KERNEL, GRID, ST, ARGS)                                             \
    switch (D) {                                                                                      \
        case 256: if (OUT_F32) KERNEL<256, float><<<GRID, 256, 0, ST>>>(ARGS); else KERNEL<256, __nv_bfloat16><<<GRID, 256, 0, ST>>>(ARGS); break;   \
        case 512: if (OUT_F32) KERNEL<512, float><<<GRID, 256, 0, ST>>>(ARGS); else KERNEL<512, __nv_bfloat16><<<GRID, 256, 0, ST>>>(ARGS); break;   \
        case 768: if (OUT_F32) KERNEL<768, float><<<GRID, 256, 0, ST>>>(ARGS); else KERNEL<768, __nv_bfloat16><<<GRID, 256, 0, ST>>>(ARGS); break;   \
        case 1024: if (OUT_F32) KERNEL<1024, float><<<GRID, 256, 0, ST>>>(ARGS); else KERNEL<1024, __nv_bfloat16><<<GRID, 256, 0, ST>>>(ARGS); break; \
    }

// global-mean partial blocks per image the forward needs room for (one per 32-token tile)
extern "C" int mm_combine_num_token_blocks(int P) { return (P + RUN_TOKENS - 1) / RUN_TOKENS; }

template <typename K>
static int opt_in_smem(K kern, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) {
        mm::set_error("%s: cannot opt in to %zu B of shared memory (%s)", what, bytes, cudaGetErrorString(e));
        return MM_ERR_CUDA;
    }
    return MM_OK;
}
#define MM_STAGED_D(D, F32, FN, ARGS, ST, RC)                                                            \
    switch (D) {                                                                                         \
        case 256: RC = F32 ? FN<256, float>(ARGS, ST) : FN<256, __nv_bfloat16>(ARGS, ST); break;          \
        case 512: RC = F32 ? FN<512, float>(ARGS, ST) : FN<512, __nv_bfloat16>(ARGS, ST); break;          \
        case 768: RC = F32 ? FN<768, float>(ARGS, ST) : FN<768, __nv_bfloat16>(ARGS, ST); break;          \
        case 1024: RC = F32 ? FN<1024, float>(ARGS, ST) : FN<1024, __nv_bfloat16>(ARGS, ST); break;       \
    }
extern "C" int mm_combine_num_row_blocks(const int32_t* Ps) {
    return (Ps[0] + Ps[1] + Ps[2] + Ps[3] + CB_ROWS_PER_BLOCK - 1) / CB_ROWS_PER_BLOCK;
}

// tensor-core forward combine (combine_mma.cuh): 0 = launched, 1 = does not apply (caller falls back), < 0 = error
static int cm_launch_out(CombineArgs& a, CmArgs& c, int D, long long total_rows, int out_f32, cudaStream_t st) {
    if (!c.tile_info || !c.seg_start || !c.offsets || !a.perm || !cm_geometry(a, D, c)) return 1;
    for (int s = 0; s < 4; ++s) a.ratio[s] = a.P / a.Ps[s];
    c.D = D;
    c.n_pass = D / CM_BN;
    c.out_f32 = out_f32;
    const bool img = a.topk > 1;        // top-k > 1: image-centric tiles, the k choices are extra K groups of the same accumulator
    c.n_src = img ? a.topk : 1;
    c.stages = img ? 2 : CM_STAGES;
    c.tiles_per_img = (a.P + TILE_M - 1) / TILE_M;
    if (img && !a.inv_perm) return 1;
    const size_t smem = cm_smem_bytes(c, out_f32 != 0);
    if (smem > 227 * 1024) return 1;
    CUtensorMap tmY[4], tmOut;
    for (int s = 0; s < 4; ++s) {
        int rc = mm::encode_tmap_bf16(&tmY[s], a.Y, static_cast<uint64_t>(D), static_cast<uint64_t>(total_rows),
                                      static_cast<uint64_t>(D), 64, static_cast<uint32_t>(c.cap[s]), "combine_out(Y)");
        if (rc) return rc;
    }
    tmOut = tmY[0];
    if (!out_f32) {
        int rc = mm::encode_tmap_bf16_swz(&tmOut, a.out, static_cast<uint64_t>(D), static_cast<uint64_t>(a.B) * a.P,
                                          static_cast<uint64_t>(D), 32, 32, 64, "combine_out(out)");
        if (rc) return rc;
    }
    const int n_work = img ? a.B * c.tiles_per_img : c.n_tiles;
    const int grid = n_work < mm::sm_count() ? n_work : mm::sm_count();
#define MM_CM_LAUNCH(F32, IMGV)                                                                          \
    {                                                                                                    \
        auto kern = cm_out_kernel<F32, IMGV>;                                                            \
        if (int rc = opt_in_smem(kern, smem, "combine_out(mma)")) return rc;                             \
        kern<<<grid, CM_THREADS, smem, st>>>(tmY[0], tmY[1], tmY[2], tmY[3], tmOut, a, c);               \
    }
    if (out_f32) { if (img) MM_CM_LAUNCH(true, true) else MM_CM_LAUNCH(true, false) }
    else         { if (img) MM_CM_LAUNCH(false, true) else MM_CM_LAUNCH(false, false) }
#undef MM_CM_LAUNCH
    mm::note_launches(1);
    return MM_OK;
}

// tensor-core logits + softmax (combine_mma.cuh): 0 = launched, 1 = does not apply, < 0 = error
static int cm_launch_logits(CombineArgs& a, ClArgs& c, int D, long long total_rows, cudaStream_t st) {
    if (!c.tile_info || !c.seg_start || !c.offsets || !cl_geometry(a, D, c)) return 1;
    for (int s = 0; s < 4; ++s) a.ratio[s] = a.P / a.Ps[s];
    const size_t smem = cl_smem_bytes(c);
    if (smem > 227 * 1024) return 1;
    CUtensorMap tmZ[4];
    for (int s = 0; s < 4; ++s) {
        int rc = mm::encode_tmap_bf16(&tmZ[s], a.Z, static_cast<uint64_t>(c.H), static_cast<uint64_t>(total_rows),
                                      static_cast<uint64_t>(c.H), 64, static_cast<uint32_t>(c.cap[s]), "combine_logits(Z)");
        if (rc) return rc;
    }
    auto kern = cm_logits_kernel;
    if (int rc = opt_in_smem(kern, smem, "combine_logits(mma)")) return rc;
    const int grid = c.n_tiles < mm::sm_count() ? c.n_tiles : mm::sm_count();
    kern<<<grid, CM_THREADS, smem, st>>>(tmZ[0], tmZ[1], tmZ[2], tmZ[3], a, c);
    mm::note_launches(1);
    return MM_OK;
}

extern "C" int mm_interp_softmax_combine_fwd(const void* Y, const void* Z, const float* w2, const float* b2, int B, int topk,
                                             int P, const int32_t* Ps, int D, const int32_t* inv_perm,
                                             const int32_t* slot_expert, const int32_t* slot_row, const float* gate,
                                             float* beta, void* out, int out_f32, float* gpart, float* global_feat,
                                             const int32_t* perm, const int32_t* seg_start, const int32_t* offsets,
                                             const int32_t* tile_info0, int n_tiles0, int region0_row, int K,
                                             long long total_rows, int flags, void* stream) {
    CombineArgs a{};
    int rc = fill_common(a, B, topk, P, Ps, D, "mm_interp_softmax_combine_fwd");
    if (rc) return rc;
    a.inv_perm = inv_perm; a.slot_expert = slot_expert; a.slot_row = slot_row; a.gate = gate;
    a.Y = static_cast<const __nv_bfloat16*>(Y); a.Z = static_cast<const __nv_bfloat16*>(Z);
    a.w2 = w2; a.b2 = b2; a.beta = beta; a.out = out; a.gpart = gpart;
    a.nblk = mm_combine_num_token_blocks(P);
    a.nruns = (P + RUN_TOKENS - 1) / RUN_TOKENS;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // TMA-staged persistent kernels (any scale ratio); the direct-load kernel remains for shapes whose tile does not fit
    sg_setup_tiles(a, 32);
    mm::trace_mark("begin", st);
    int staged = 1;
    {   // logits + softmax on the tensor cores when the geometry allows it
        ClArgs cl{};
        cl.n_tiles = n_tiles0;
        cl.tile_info = reinterpret_cast<const int2*>(tile_info0);
        cl.region_row0 = region0_row;
        cl.seg_start = seg_start; cl.offsets = offsets; cl.K = K;
        const int mma = (flags & 1) ? 1 : cm_launch_logits(a, cl, D, total_rows, st);
        if (mma < 0) return mma;
        if (mma == 0) staged = 0;
    }
    if (staged != 0) switch (D) {
        case 256: staged = sg_launch_logits<256>(a, st); break;
        case 512: staged = sg_launch_logits<512>(a, st); break;
        case 768: staged = sg_launch_logits<768>(a, st); break;
        case 1024: staged = sg_launch_logits<1024>(a, st); break;
    }
    if (staged < 0) return staged;
    if (staged == 0) {
        mm::trace_mark("combine_fwd.logits", st);
        // out = C Yrows on the tensor cores when the geometry allows it, else the CUDA-core staged kernel
        CmArgs c{};
        c.n_tiles = n_tiles0;
        c.tile_info = reinterpret_cast<const int2*>(tile_info0);
        c.region_row[0] = region0_row;
        c.seg_start = seg_start; c.offsets = offsets; c.K = K;
        a.perm = perm;
        a.nblk = mm_combine_num_token_blocks(P);
        int mma = (flags & 1) ? 1 : cm_launch_out(a, c, D, total_rows, out_f32, st);
        if (mma < 0) return mma;
        if (mma == 0) {
            mm::trace_mark("combine_fwd.out", st);
        } else {
            a.nblk = a.tiles_per_img;
            MM_STAGED_D(D, out_f32, sg_launch_out, a, st, staged)
            if (staged < 0) return staged;
            if (staged == 0) mm::trace_mark("combine_fwd.out", st);
        }
    }
    if (staged != 0) {
        a.nblk = (P + CB_TOKENS_PER_BLOCK - 1) / CB_TOKENS_PER_BLOCK;
        dim3 grid(a.nblk, B);
        MM_DISPATCH_D(D, out_f32, combine_fwd_kernel, grid, st, a)
        mm::note_launches(1);
    }
    rc = mm_check_launch("mm_interp_softmax_combine_fwd");
    if (rc) return rc;
    global_mean_kernel<<<dim3((D + 255) / 256, B), 256, 0, st>>>(gpart, a.nblk, D, 1.0f / static_cast<float>(P), global_feat);
    mm::note_launches(1);
    mm::trace_mark("combine_fwd.global_mean", st);
    return mm_check_launch("mm_interp_softmax_combine_fwd(global mean)");
}

extern "C" int mm_combine_num_runs(int P) { return (P + RUN_TOKENS - 1) / RUN_TOKENS; }
// number of per-slot partial blocks `part` must hold (max over the two backward paths)
extern "C" int mm_combine_num_part_blocks(int P, const int32_t* Ps) {
    int fast = (mm_combine_num_runs(P) + RUNS_PER_BLOCK - 1) / RUNS_PER_BLOCK;
    fast += ((Ps[1] + ZR_ROWS_PER_WARP - 1) / ZR_ROWS_PER_WARP + (Ps[2] + ZR_ROWS_PER_WARP - 1) / ZR_ROWS_PER_WARP +
             (Ps[3] + ZR_ROWS_PER_WARP - 1) / ZR_ROWS_PER_WARP + ZR_WARPS - 1) / ZR_WARPS;      // interval-prefix dZ path
    const int generic = mm_combine_num_row_blocks(Ps);
    return fast > generic ? fast : generic;
}
// floats of scratch per item that `mom_z` must provide (max of the moment layout and the interval-prefix layout)
extern "C" long long mm_combine_bwd_z_scratch_floats(int P, const int32_t* Ps, int D) {
    long long need = static_cast<long long>(mm_combine_num_runs(P)) * 2 * (D / 2);
    bool ok = Ps[0] == P;
    for (int s = 1; s < 4 && ok; ++s) ok = Ps[s] > 0 && P % Ps[s] == 0;
    if (ok) {
        int ps[4] = {Ps[0], Ps[1], Ps[2], Ps[3]};
        const long long z = z_scratch_layout(P, ps).per_slot;
        if (z > need) need = z;
    }
    return need;
}

// classify the scale ratios for the token-centric backward; returns 1 when it applies
static int classify_scales(CombineArgs& a) {
    int n_moment = 0;
    a.halo = 0;
    for (int s = 0; s < 4; ++s) {
        a.ratio[s] = 0; a.mode[s] = SCALE_UNSUPPORTED;
        if (a.Ps[s] <= 0 || a.P % a.Ps[s] != 0) return 0;
        const int r = a.P / a.Ps[s];
        a.ratio[s] = r;
        if (s == 0) {
            if (r != 1) return 0;
            a.mode[s] = SCALE_IDENT;
        } else if (r <= RUN_TOKENS && RUN_TOKENS % r == 0) {
            a.mode[s] = SCALE_DIRECT;
            if (r / 2 > a.halo) a.halo = r / 2;
        } else if (r % 2 == 0 && (r / 2) % RUN_TOKENS == 0) {
            a.mode[s] = SCALE_MOMENT;
            ++n_moment;
        } else {
            return 0;
        }
    }
    return n_moment <= 1;
}

extern "C" int mm_interp_softmax_combine_bwd(const void* Y, const void* Z, const float* w2, int B, int topk, int P,
                                             const int32_t* Ps, int D, int K, const int32_t* perm, const int32_t* inv_perm,
                                             const int32_t* slot_expert, const int32_t* slot_row, const int32_t* counts,
                                             const int32_t* seg_start, const int32_t* offsets, const float* gate,
                                             const float* beta, const void* dlocal, int dlocal_f32, const float* dglobal,
                                             float* dlogit, float* dgate, void* dUT, void* dZ, float* part,
                                             float* dw2_db1_db2, float* mom_u, float* mom_z, int force_generic,
                                             void* stream) {
    CombineArgs a{};
    int rc = fill_common(a, B, topk, P, Ps, D, "mm_interp_softmax_combine_bwd");
    if (rc) return rc;
    a.perm = perm; a.inv_perm = inv_perm; a.slot_expert = slot_expert; a.slot_row = slot_row; a.gate = gate;
    a.counts = counts; a.seg_start = seg_start; a.K = K;
    a.Y = static_cast<const __nv_bfloat16*>(Y); a.Z = static_cast<const __nv_bfloat16*>(Z);
    a.w2 = w2; a.beta = const_cast<float*>(beta);
    a.dlocal = dlocal; a.dglobal = dglobal; a.dlogit = dlogit; a.dgate = dgate;
    a.dUT = static_cast<__nv_bfloat16*>(dUT); a.dZ = static_cast<__nv_bfloat16*>(dZ);
    a.part = part; a.mom_u = mom_u; a.mom_z = mom_z;
    a.nblk = mm_combine_num_token_blocks(P);
    a.nruns = mm_combine_num_runs(P);
    const bool fast = !force_generic && mom_u && mom_z && classify_scales(a);
    a.nrb = fast ? (a.nruns + RUNS_PER_BLOCK - 1) / RUNS_PER_BLOCK : mm_combine_num_row_blocks(Ps);
    a.zscr = mom_z;
    a.z_rows_path = fast && z_rows_path_ok(a) && !(force_generic & 2);
    if (a.z_rows_path) a.nrb = z_ident_blocks(a) + z_rows_blocks(a);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!fast) {
        dim3 grid(a.nblk, B);
        MM_DISPATCH_D(D, dlocal_f32, combine_bwd_logit_kernel, grid, st, a)
        mm::note_launches(1);
        rc = mm_check_launch("mm_interp_softmax_combine_bwd(logit)");
        if (rc) return rc;
    }
    if (fast) {
        {
            sg_setup_tiles(a, 16);
            a.dlogit_is_halves = 1;
            mm::trace_mark("begin", st);
            int staged = 1;
            MM_STAGED_D(D, dlocal_f32, sg_launch_bwd_dbeta, a, st, staged)
            if (staged < 0) return staged;
            if (staged != 0) {   // tile does not fit in shared memory: run-based direct-load kernel (column-half partial dbeta)
                a.dlogit_is_halves = 1;
                dim3 grid((a.nruns + OUT_RUNS_PER_BLOCK - 1) / OUT_RUNS_PER_BLOCK, B);
                MM_DISPATCH_D(D, dlocal_f32, combine_bwd_dbeta_kernel, grid, st, a)
                mm::note_launches(1);
            }
        }
        {
            dim3 grid((2 * a.nruns + 7) / 8, a.n_items);
            mm::trace_mark("combine_bwd.dbeta", st);
            MM_DISPATCH_D(D, dlocal_f32, combine_bwd_u_kernel, grid, st, a)
            mm::note_launches(1);
            mm::trace_mark("combine_bwd.dUT", st);
        }
        if (a.z_rows_path) {   // interval prefix sums: O(rows) instead of O(tokens x scales)
            switch (D) {
                case 256: rc = launch_bwd_z_rows_path<256>(a, st); break;
                case 512: rc = launch_bwd_z_rows_path<512>(a, st); break;
                case 768: rc = launch_bwd_z_rows_path<768>(a, st); break;
                case 1024: rc = launch_bwd_z_rows_path<1024>(a, st); break;
            }
            if (rc) return rc;
        } else {
            dim3 gridz(a.nrb, a.n_items + K);
            switch (D) {
                case 256: combine_bwd_z_kernel<256><<<gridz, 256, 0, st>>>(a); break;
                case 512: combine_bwd_z_kernel<512><<<gridz, 256, 0, st>>>(a); break;
                case 768: combine_bwd_z_kernel<768><<<gridz, 256, 0, st>>>(a); break;
                case 1024: combine_bwd_z_kernel<1024><<<gridz, 256, 0, st>>>(a); break;
            }
            mm::note_launches(1);
            mm::trace_mark("combine_bwd.dZ", st);
        }
        for (int s = 1; s < 4; ++s) {
            if (a.mode[s] != SCALE_MOMENT) continue;
            dim3 gridf((a.Ps[s] + 7) / 8, a.n_items);
            switch (D) {
                case 256: combine_bwd_finalize_kernel<256><<<gridf, 256, 0, st>>>(a, s); break;
                case 512: combine_bwd_finalize_kernel<512><<<gridf, 256, 0, st>>>(a, s); break;
                case 768: combine_bwd_finalize_kernel<768><<<gridf, 256, 0, st>>>(a, s); break;
                case 1024: combine_bwd_finalize_kernel<1024><<<gridf, 256, 0, st>>>(a, s); break;
            }
            mm::note_launches(1);
        }
        mm::trace_mark("combine_bwd.finalize", st);
        rc = mm_check_launch("mm_interp_softmax_combine_bwd(token-centric)");
        if (rc) return rc;
    } else {
        dim3 grid(a.nrb, a.n_items + K);
        MM_DISPATCH_D(D, dlocal_f32, combine_bwd_rows_kernel, grid, st, a)
        mm::note_launches(1);
        rc = mm_check_launch("mm_interp_softmax_combine_bwd(rows)");
        if (rc) return rc;
    }
    const int C = 2 * (D / 2) + 1;
    expert_reduce_kernel<<<dim3((C + 31) / 32, K), 32 * ER_GROUPS, 0, st>>>(part, offsets, a.nrb, C, dw2_db1_db2);
    mm::note_launches(1);
    mm::trace_mark("combine_bwd.expert_reduce", st);
    return mm_check_launch("mm_interp_softmax_combine_bwd(reduce)");
}

// tensor-core dbeta (combine_mma.cuh): 0 = launched, 1 = does not apply, < 0 = error
static int cm_launch_dbeta(CombineArgs& a, CdArgs& c, int D, long long total_rows, const void* dlocal, float* dbeta_loc,
                           cudaStream_t st) {
    CmArgs g{};
    CombineArgs one = a;
    one.topk = 1;
    if (!c.tile_info || !cm_geometry(one, D, g) || a.P % 64 != 0 || D % 64 != 0) return 1;
    for (int s = 0; s < 4; ++s) { a.ratio[s] = a.P / a.Ps[s]; c.cap[s] = g.cap[s]; c.koff[s] = g.koff[s]; }
    c.ktot = g.ktot;
    c.D = D;
    c.n_kb = D / 64;
    c.stages = 5;
    c.dbeta_loc = dbeta_loc;
    const size_t stage = static_cast<size_t>(TILE_M) * 128 + static_cast<size_t>(c.ktot) * 128;
    while (c.stages > 2 && c.stages * stage + 512 + 1024 > 227 * 1024) --c.stages;
    const size_t smem = c.stages * stage + 512 + 1024;
    if (smem > 227 * 1024) return 1;
    CUtensorMap tmDL, tmY[4];
    int rc = mm::encode_tmap_bf16(&tmDL, dlocal, static_cast<uint64_t>(D), static_cast<uint64_t>(a.B) * a.P, static_cast<uint64_t>(D),
                                  64, 64, "combine_dbeta(dlocal)");
    if (rc) return rc;
    for (int s = 0; s < 4; ++s) {
        rc = mm::encode_tmap_bf16(&tmY[s], a.Y, static_cast<uint64_t>(D), static_cast<uint64_t>(total_rows),
                                  static_cast<uint64_t>(D), 64, static_cast<uint32_t>(c.cap[s]), "combine_dbeta(Y)");
        if (rc) return rc;
    }
    auto kern = cm_dbeta_kernel;
    if (int r2 = opt_in_smem(kern, smem, "combine_dbeta(mma)")) return r2;
    const int grid = c.n_tiles < mm::sm_count() ? c.n_tiles : mm::sm_count();
    kern<<<grid, CD_THREADS, smem, st>>>(tmDL, tmY[0], tmY[1], tmY[2], tmY[3], a, c);
    mm::note_launches(1);
    return MM_OK;
}

static int force_cuda_core_dut = 0;     // test hook (mm_debug_force_cuda_core_dut): run the CUDA-core dUT kernel instead of cm_dut_kernel
extern "C" void mm_debug_force_cuda_core_dut(int on) { force_cuda_core_dut = on; }

// ---- backward on the tensor-core / rank-1 path (combine_rank1.cuh, combine_mma.cuh) ----
//   dglobal != NULL : dF's per-image constant -> row_dot / row_coef / row_img (rank-1, never materialised)
//   dlocal  != NULL : (bf16) dbeta by GEMM (cm_dbeta_kernel) and dUT [rows, D] bf16 (the local part of d fused / d Y)
extern "C" int mm_combine_bwd_global_supported(int P, const int32_t* Ps, int D) {
    if (!(D == 256 || D == 512 || D == 768 || D == 1024)) return 0;
    CombineArgs a{};
    a.P = P;
    for (int s = 0; s < 4; ++s) a.Ps[s] = Ps[s];
    return z_rows_path_ok(a) ? 1 : 0;
}
extern "C" int mm_combine_bwd_tc_supported(int P, const int32_t* Ps, int D) {
    if (!mm_combine_bwd_global_supported(P, Ps, D)) return 0;
    CombineArgs a{};
    a.P = P; a.topk = 1;
    for (int s = 0; s < 4; ++s) a.Ps[s] = Ps[s];
    CmArgs g{};
    return (cm_geometry(a, D, g) && P % 64 == 0 && classify_scales(a)) ? 1 : 0;
}

extern "C" int mm_interp_softmax_combine_bwd_tc(const void* Y, const void* Z, const float* w2, int B, int topk, int P,
                                                const int32_t* Ps, int D, int K, const int32_t* perm,
                                                const int32_t* inv_perm, const int32_t* slot_expert,
                                                const int32_t* slot_row, const int32_t* counts, const int32_t* seg_start,
                                                const int32_t* offsets, const int32_t* tile_info0, int n_tiles0,
                                                int region0_row, long long total_rows, const float* gate, const float* beta,
                                                const void* dlocal, const float* dglobal, float* row_dot, float* row_coef,
                                                int32_t* row_img, float* dbeta_loc, void* dUT, float* mom_u, float* dgate,
                                                void* dZ, float* part, float* dw2_db1_db2, float* zscr, void* stream) {
    CombineArgs a{};
    int rc = fill_common(a, B, topk, P, Ps, D, "mm_interp_softmax_combine_bwd_tc");
    if (rc) return rc;
    if (!(Y && Z && w2 && beta && dZ && part && dw2_db1_db2 && zscr && (dglobal || dlocal))) {
        mm::set_error("mm_interp_softmax_combine_bwd_tc: null operand");
        return MM_ERR_BAD_SHAPE;
    }
    if (dglobal && !(row_dot && row_coef && row_img)) {
        mm::set_error("mm_interp_softmax_combine_bwd_tc: dglobal needs row_dot / row_coef / row_img");
        return MM_ERR_BAD_SHAPE;
    }
    if (dlocal && !(dbeta_loc && dUT && mom_u && tile_info0)) {
        mm::set_error("mm_interp_softmax_combine_bwd_tc: dlocal needs dbeta_loc / dUT / mom_u / tile_info0");
        return MM_ERR_BAD_SHAPE;
    }
    a.perm = perm; a.inv_perm = inv_perm; a.slot_expert = slot_expert; a.slot_row = slot_row; a.gate = gate;
    a.counts = counts; a.seg_start = seg_start; a.K = K;
    a.Y = static_cast<const __nv_bfloat16*>(Y); a.Z = static_cast<const __nv_bfloat16*>(Z);
    a.w2 = w2; a.beta = const_cast<float*>(beta);
    a.dglobal = dglobal; a.dgate = dgate; a.dZ = static_cast<__nv_bfloat16*>(dZ);
    a.part = part; a.zscr = zscr; a.mom_z = zscr;
    a.row_dot = dglobal ? row_dot : nullptr;
    a.nblk = mm_combine_num_token_blocks(P);
    a.nruns = mm_combine_num_runs(P);
    if (!z_rows_path_ok(a)) {
        mm::set_error("mm_interp_softmax_combine_bwd_tc: needs Ps[0] == P and even integer scale ratios (use mm_interp_softmax_combine_bwd)");
        return MM_ERR_UNSUPPORTED;
    }
    for (int s = 0; s < 4; ++s) a.ratio[s] = P / Ps[s];
    a.z_rows_path = 1;
    a.nrb = z_ident_blocks(a) + z_rows_blocks(a);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int total = Ps[0] + Ps[1] + Ps[2] + Ps[3];
    mm::trace_mark("begin", st);
    if (dglobal) {
        dim3 grid((total + R1_WARPS * R1_ROWS_PER_WARP - 1) / (R1_WARPS * R1_ROWS_PER_WARP), a.n_items);
        switch (D) {
            case 256: rank1_rowdot_kernel<256><<<grid, R1_WARPS * 32, 0, st>>>(a, row_dot, row_img); break;
            case 512: rank1_rowdot_kernel<512><<<grid, R1_WARPS * 32, 0, st>>>(a, row_dot, row_img); break;
            case 768: rank1_rowdot_kernel<768><<<grid, R1_WARPS * 32, 0, st>>>(a, row_dot, row_img); break;
            case 1024: rank1_rowdot_kernel<1024><<<grid, R1_WARPS * 32, 0, st>>>(a, row_dot, row_img); break;
        }
        mm::note_launches(1);
        mm::trace_mark("combine_bwd.rowdot", st);
        rank1_coef_kernel<<<dim3((rank1_coef_units(a) + 255) / 256, a.n_items), 256, 0, st>>>(a, row_coef);
        mm::note_launches(1);
        mm::trace_mark("combine_bwd.coef", st);
        rc = mm_check_launch("mm_interp_softmax_combine_bwd_tc(rank-1)");
        if (rc) return rc;
    }
    if (dlocal) {
        CdArgs cd{};
        cd.n_tiles = n_tiles0;
        cd.tile_info = reinterpret_cast<const int2*>(tile_info0);
        cd.region_row0 = region0_row;
        cd.seg_start = seg_start; cd.offsets = offsets; cd.K = K;
        const int r = cm_launch_dbeta(a, cd, D, total_rows, dlocal, dbeta_loc, st);
        if (r < 0) return r;
        if (r != 0) {
            mm::set_error("mm_interp_softmax_combine_bwd_tc: geometry not supported (check mm_combine_bwd_tc_supported)");
            return MM_ERR_UNSUPPORTED;
        }
        a.dbeta_loc = dbeta_loc;
        mm::trace_mark("combine_bwd.dbeta", st);
        // dUT = C^T dlocal (local part only: the dglobal part stays rank-1)
        bool dut_mma = !(force_cuda_core_dut);
        for (int s = 1; s < 4; ++s) dut_mma = dut_mma && a.ratio[s] >= 2 && a.ratio[s] <= 2 * CU_HALO;
        if (dut_mma) {      // tensor cores: tile-owned rows, 32-token halo (combine_mma.cuh)
            CuArgs cu{};
            cu.n_tiles = n_tiles0;
            cu.tile_info = reinterpret_cast<const int2*>(tile_info0);
            cu.region_row0 = region0_row;
            cu.seg_start = seg_start; cu.offsets = offsets; cu.counts = counts; cu.K = K;
            int orow = 0;
            for (int s = 1; s < 4; ++s) { cu.n_own[s] = TILE_M / a.ratio[s]; cu.o_row[s] = orow; orow += cu.n_own[s]; }
            cu.D = D;
            cu.n_pass = D / CU_BN;
            cu.dUT = static_cast<__nv_bfloat16*>(dUT);
            CUtensorMap tmDL, tmOut;
            rc = mm::encode_tmap_bf16(&tmDL, dlocal, static_cast<uint64_t>(D), static_cast<uint64_t>(B) * P, static_cast<uint64_t>(D),
                                      64, 32, "combine_dut(dlocal)");
            if (rc) return rc;
            rc = mm::encode_tmap_bf16_swz(&tmOut, dUT, static_cast<uint64_t>(D), static_cast<uint64_t>(total_rows),
                                          static_cast<uint64_t>(D), 32, 32, 64, "combine_dut(dUT)");
            if (rc) return rc;
            const size_t smem = cu_smem_bytes();
            auto kern = cm_dut_kernel;
            if (int r2 = opt_in_smem(kern, smem, "combine_dut(mma)")) return r2;
            const int grid = n_tiles0 < mm::sm_count() ? n_tiles0 : mm::sm_count();
            kern<<<grid, CU_THREADS, smem, st>>>(tmDL, tmOut, a, cu);
            mm::note_launches(1);
        } else {            // token-centric CUDA-core kernel
            CombineArgs u = a;
            u.dlocal = dlocal; u.dglobal = nullptr; u.dUT = static_cast<__nv_bfloat16*>(dUT); u.mom_u = mom_u;
            if (!classify_scales(u)) {
                mm::set_error("mm_interp_softmax_combine_bwd_tc: scale ratios not supported by the dUT kernel");
                return MM_ERR_UNSUPPORTED;
            }
            u.z_rows_path = 1;      // the finalize kernel then only writes dUT
            dim3 grid((2 * u.nruns + 7) / 8, u.n_items);
            MM_DISPATCH_D(D, 0, combine_bwd_u_kernel, grid, st, u)
            mm::note_launches(1);
            for (int s = 1; s < 4; ++s) {
                if (u.mode[s] != SCALE_MOMENT) continue;
                dim3 gridf((u.Ps[s] + 7) / 8, u.n_items);
                switch (D) {
                    case 256: combine_bwd_finalize_kernel<256><<<gridf, 256, 0, st>>>(u, s); break;
                    case 512: combine_bwd_finalize_kernel<512><<<gridf, 256, 0, st>>>(u, s); break;
                    case 768: combine_bwd_finalize_kernel<768><<<gridf, 256, 0, st>>>(u, s); break;
                    case 1024: combine_bwd_finalize_kernel<1024><<<gridf, 256, 0, st>>>(u, s); break;
                }
                mm::note_launches(1);
            }
        }
        mm::trace_mark("combine_bwd.dUT", st);
        rc = mm_check_launch("mm_interp_softmax_combine_bwd_tc(dUT)");
        if (rc) return rc;
    }
    switch (D) {
        case 256: rc = launch_bwd_z_rows_path<256>(a, st); break;
        case 512: rc = launch_bwd_z_rows_path<512>(a, st); break;
        case 768: rc = launch_bwd_z_rows_path<768>(a, st); break;
        case 1024: rc = launch_bwd_z_rows_path<1024>(a, st); break;
    }
    if (rc) return rc;
    const int C = 2 * (D / 2) + 1;
    expert_reduce_kernel<<<dim3((C + 31) / 32, K), 32 * ER_GROUPS, 0, st>>>(part, offsets, a.nrb, C, dw2_db1_db2);
    mm::note_launches(1);
    mm::trace_mark("combine_bwd.expert_reduce", st);
    return mm_check_launch("mm_interp_softmax_combine_bwd_tc(reduce)");
}

extern "C" int mm_interp_softmax_combine_bwd_global(const void* Y, const void* Z, const float* w2, int B, int topk, int P,
                                                    const int32_t* Ps, int D, int K, const int32_t* perm,
                                                    const int32_t* inv_perm, const int32_t* slot_expert,
                                                    const int32_t* slot_row, const int32_t* counts, const int32_t* seg_start,
                                                    const int32_t* offsets, const float* gate, const float* beta,
                                                    const float* dglobal, float* row_dot, float* row_coef, int32_t* row_img,
                                                    float* dgate, void* dZ, float* part, float* dw2_db1_db2, float* zscr,
                                                    void* stream) {
    if (!dglobal) {
        mm::set_error("mm_interp_softmax_combine_bwd_global: dglobal is NULL");
        return MM_ERR_BAD_SHAPE;
    }
    return mm_interp_softmax_combine_bwd_tc(Y, Z, w2, B, topk, P, Ps, D, K, perm, inv_perm, slot_expert, slot_row, counts,
                                            seg_start, offsets, nullptr, 0, 0, 0, gate, beta, nullptr, dglobal, row_dot,
                                            row_coef, row_img, nullptr, nullptr, nullptr, dgate, dZ, part, dw2_db1_db2, zscr,
                                            stream);
}
