// medmoe_b200 — backward of the combine when only global_feat carries a cotangent (dlocal == NULL).
// Included by combine.cu after CombineArgs and the row helpers.
//
// global_feat = mean_p fused  =>  dF(p) = dglobal[b] / P for every token p of image b: a per-image constant.
// Autograd of swin.py:68-80,110 then collapses to scalar fields over the native rows:
//   dbeta_s(p) = <dF, interp(Y_s)(p)> = interp(G_s)(p)            with  G[row] = <dglobal[b], Y[row]> / P
//   dU_s[i, :] = (sum_p w_i(p) g beta_s(p)) dF = c[row] * dglobal[b, :]   with  c[row] = sum_p w_i(p) g beta_s(p) / P
// so the [rows, D] gradient w.r.t. the projected features is rank-1 per image and is never written: the dY GEMM
// (mm_grouped_gemm_rows_rank1) rebuilds it in its epilogue from c[row] and row_img[row].
//   rank1_rowdot_kernel : one streaming pass over Y -> G [rows] fp32, row_img [rows]   (HBM-bound: reads Y once)
//   rank1_coef_kernel   : c[row] from beta (a few KB per image)
// The dlogit / dZ kernels of combine_bwd_z.cuh consume G through token_dlogit().
#pragma once

namespace mm {

constexpr int R1_ROWS_PER_WARP = 8;
constexpr int R1_WARPS = 8;

// (scale, native row) of the u-th row of an item when the four scales are concatenated
MM_DEVINL void r1_split(const CombineArgs& a, int u, int& s, int& i) {
    s = 0;
    while (s < 3 && u >= a.Ps[s]) { u -= a.Ps[s]; ++s; }
    i = u;
}

// grid = (ceil(sum Ps / 64), n_items); warp = 8 consecutive rows of one item
template <int D>
__global__ void __launch_bounds__(R1_WARPS * 32)
rank1_rowdot_kernel(const CombineArgs a, float* __restrict__ row_dot, int* __restrict__ row_img) {
    constexpr int N = D / 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.y;
    const int item = a.perm[slot];
    const int b = item / a.topk;
    const int total = a.Ps[0] + a.Ps[1] + a.Ps[2] + a.Ps[3];
    const int u0 = (blockIdx.x * R1_WARPS + warp) * R1_ROWS_PER_WARP;
    if (u0 >= total) return;
    float dg[N * 8];
    load_row_x8<N, float>(a.dglobal + static_cast<size_t>(b) * D, lane, dg);
    const float inv_p = 1.0f / static_cast<float>(a.P);
    long long rows[R1_ROWS_PER_WARP];
#pragma unroll
    for (int t = 0; t < R1_ROWS_PER_WARP; ++t) {
        int s, i;
        r1_split(a, min(u0 + t, total - 1), s, i);
        rows[t] = static_cast<long long>(a.slot_row[s * a.n_items + slot]) + i;
    }
    float mine = 0.f;
#pragma unroll
    for (int h = 0; h < R1_ROWS_PER_WARP; h += 4) {
        uint4 v[4][N];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int c = 0; c < N; ++c) v[t][c] = ldg_nc_v4(a.Y + rows[h + t] * D + 8 * (lane + 32 * c));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                acc = fmaf(dg[8 * c + 0], bf16lo(v[t][c].x), acc); acc = fmaf(dg[8 * c + 1], bf16hi(v[t][c].x), acc);
                acc = fmaf(dg[8 * c + 2], bf16lo(v[t][c].y), acc); acc = fmaf(dg[8 * c + 3], bf16hi(v[t][c].y), acc);
                acc = fmaf(dg[8 * c + 4], bf16lo(v[t][c].z), acc); acc = fmaf(dg[8 * c + 5], bf16hi(v[t][c].z), acc);
                acc = fmaf(dg[8 * c + 6], bf16lo(v[t][c].w), acc); acc = fmaf(dg[8 * c + 7], bf16hi(v[t][c].w), acc);
            }
            acc = warp_sum(acc) * inv_p;
            if (lane == h + t) mine = acc;
        }
    }
    if (lane < R1_ROWS_PER_WARP && u0 + lane < total) {
        int s, i;
        r1_split(a, u0 + lane, s, i);
        const long long r = static_cast<long long>(a.slot_row[s * a.n_items + slot]) + i;
        row_dot[r] = mine;
        row_img[r] = b;
    }
}

// c[row] = g / P * sum over the row's token window of w_i(p) beta_s(p).
// grid = (ceil(units / 256), n_items).  The two finest scales take one thread per native row (windows of 3 and ~10 tokens at
// 224^2); the rows of the two coarsest scales have windows of ~34 and ~130 tokens, which one thread walks in 130 dependent
// steps while the rest of the grid has long finished — they take a whole warp each (lanes stride the window, shuffle reduction).
__global__ void __launch_bounds__(256)
rank1_coef_kernel(const CombineArgs a, float* __restrict__ row_coef) {
    const int slot = blockIdx.y;
    const int fine = a.Ps[0] + a.Ps[1], coarse = a.Ps[2] + a.Ps[3];
    const int fine_pad = (fine + 31) & ~31;                     // warps never mix the two kinds of units
    const int unit = blockIdx.x * 256 + threadIdx.x;            // unit < fine: one thread per row; from fine_pad on 32 threads per coarse row
    const int lane = threadIdx.x & 31;
    const bool wide = unit >= fine_pad;
    int u = unit, p_step = 1, p_off = 0;
    if (wide) {
        u = fine + (unit - fine_pad) / 32;
        p_step = 32; p_off = lane;
    } else if (unit >= fine) {
        return;
    }
    if (u >= fine + coarse) return;
    int s, i;
    r1_split(a, u, s, i);
    const int item = a.perm[slot];
    const float g = a.gate ? a.gate[item] : 1.0f;
    const int Ps = a.Ps[s];
    const float scale = a.scale[s];
    // token window of native row i (same bounds as the generic gather kernel): tokens whose i0 is i-1 or i, +-1 margin
    const float inv_scale = static_cast<float>(a.P) / static_cast<float>(Ps);
    int p_lo = static_cast<int>(floorf((static_cast<float>(i) - 0.5f) * inv_scale - 0.5f)) - 1;
    int p_hi = static_cast<int>(ceilf((static_cast<float>(i) + 1.5f) * inv_scale - 0.5f)) + 1;
    if (i == 0) p_lo = 0;
    if (i == Ps - 1) p_hi = a.P;
    p_lo = max(p_lo, 0);
    p_hi = min(p_hi, a.P);
    const float* bt = a.beta + static_cast<size_t>(slot) * a.P * 4 + s;
    float acc = 0.f;
    for (int p = p_lo + p_off; p < p_hi; p += p_step) {
        const LerpSrc L = lerp_src(p, scale, Ps);
        float w = 0.f;
        if (L.i0 == i) w += 1.0f - L.lam;
        if (L.i1 == i) w += L.lam;
        if (w != 0.f) acc = fmaf(w, bt[4LL * p], acc);
    }
    if (wide) {
        acc = warp_sum(acc);
        if (lane != 0) return;
    }
    row_coef[static_cast<long long>(a.slot_row[s * a.n_items + slot]) + i] = acc * g / static_cast<float>(a.P);
}
// number of thread units of rank1_coef_kernel per item
static inline int rank1_coef_units(const CombineArgs& a) { return ((a.Ps[0] + a.Ps[1] + 31) & ~31) + 32 * (a.Ps[2] + a.Ps[3]); }

}  // namespace mm
