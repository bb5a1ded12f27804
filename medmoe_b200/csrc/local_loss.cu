// medmoe_b200 — the non-GEMM stages of the word-patch attention loss (GLORIALocalContrastiveLoss, src/losses.py:954-1026,
// attention_fn :698-736).  The six large products of its forward / backward run on the tcgen05 grouped GEMM kernels
// (gemm.cuh); these kernels are the passes in between.  Column n = caption * Wp + word of every [*, N] matrix below, words
// at or beyond the caption's length are masked (E = 0, no gradient).
//
//   forward :  S = ctx words^T (GEMM) -> E = exp(temp1 * softmax_w(S))            ll_softmax_exp_fwd
//              wcU[b] = E_b^T ctx_b (GEMM; the softmax over patches is E / colsum, and the cosine below is scale free,
//              so the column sums are never needed)
//              cos[b, n] = <word_n, wcU[b, n]> / max(|word_n| |wcU[b, n]|, eps),  sim[b, i] = log sum_w exp(temp2 cos)
//                                                                                 ll_cos_lse_fwd
//   backward:  dsim -> dwcU (bf16, GEMM operand) and the direct gradient of the words   ll_cos_lse_bwd
//              dE = ctx dwcU_b^T (GEMM) -> dS through exp / temp1 / softmax_w           ll_softmax_exp_bwd
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include "../../include/medmoe_b200.h"
#include "api_internal.h"

namespace mm {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// G lanes per (row, caption), four consecutive words per lane (one 16-byte load of S, one 8-byte store of E), 32 / G captions
// per warp; the reductions are xor-shuffles inside the G-lane group.  Work items = (row, group of 32 / G captions).
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int LL_ILP = 4;     // work items in flight per warp: the passes are bound by load latency, not by issue or bytes

template <int G>
__global__ void __launch_bounds__(256)
ll_softmax_exp_fwd_kernel(const float* __restrict__ S, long long ld_s, __nv_bfloat16* __restrict__ E, long long ld_e,
                          unsigned rows, int n_caps, int Wp, const int* __restrict__ cap_len, float temp1) {
    constexpr int CPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, w0 = (lane % G) * 4;
    const unsigned gpr = (n_caps + CPW - 1) / CPW;
    const unsigned total = rows * gpr;
    const unsigned stride = gridDim.x * (blockDim.x >> 5);
    for (unsigned t0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t0 < total; t0 += LL_ILP * stride) {
        float4 v[LL_ILP];
        long long off[LL_ILP];
        int len[LL_ILP];
        bool on[LL_ILP];
#pragma unroll
        for (int u = 0; u < LL_ILP; ++u) {
            const unsigned t = t0 + u * stride;
            const unsigned row = t / gpr;
            const int cap = static_cast<int>(t - row * gpr) * CPW + sub;
            on[u] = t < total && cap < n_caps && w0 < Wp;
            len[u] = (t < total && cap < n_caps) ? min(__ldg(cap_len + cap), Wp) : 0;
            off[u] = static_cast<long long>(cap) * Wp + w0;
            v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (on[u]) v[u] = *reinterpret_cast<const float4*>(S + static_cast<long long>(row) * ld_s + off[u]);
            off[u] += static_cast<long long>(row) * ld_e;
        }
#pragma unroll
        for (int u = 0; u < LL_ILP; ++u) {
            if (t0 + u * stride >= total) break;                 // warp-uniform
            const int ln = len[u];
            const bool m0 = w0 < ln, m1 = w0 + 1 < ln, m2 = w0 + 2 < ln, m3 = w0 + 3 < ln;
            float mx = fmaxf(fmaxf(m0 ? v[u].x : -INFINITY, m1 ? v[u].y : -INFINITY),
                             fmaxf(m2 ? v[u].z : -INFINITY, m3 ? v[u].w : -INFINITY));
            mx = group_max<G>(mx);
            const float e0 = m0 ? __expf(v[u].x - mx) : 0.f, e1 = m1 ? __expf(v[u].y - mx) : 0.f;
            const float e2 = m2 ? __expf(v[u].z - mx) : 0.f, e3 = m3 ? __expf(v[u].w - mx) : 0.f;
            const float sum = group_sum<G>((e0 + e1) + (e2 + e3));
            const float sc = ln > 0 ? temp1 / sum : 0.f;
            if (on[u]) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(m0 ? __expf(e0 * sc) : 0.f, m1 ? __expf(e1 * sc) : 0.f);
                const __nv_bfloat162 hi = __floats2bfloat162_rn(m2 ? __expf(e2 * sc) : 0.f, m3 ? __expf(e3 * sc) : 0.f);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(E + off[u]) = pk;
            }
        }
    }
}

// dE (in place -> dS): a = log(E) / temp1 is the first softmax's output, g = dE * temp1 * E, dS = a (g - sum_w a g)
template <int G>
__global__ void __launch_bounds__(256)
ll_softmax_exp_bwd_kernel(const __nv_bfloat16* __restrict__ E, long long ld_e, __nv_bfloat16* __restrict__ dE, long long ld_d,
                          unsigned rows, int n_caps, int Wp, const int* __restrict__ cap_len, float temp1) {
    constexpr int CPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, w0 = (lane % G) * 4;
    const unsigned gpr = (n_caps + CPW - 1) / CPW;
    const unsigned total = rows * gpr;
    const unsigned stride = gridDim.x * (blockDim.x >> 5);
    const float inv_t = 1.0f / temp1;
    for (unsigned t0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t0 < total; t0 += LL_ILP * stride) {
        uint2 ev[LL_ILP], dv[LL_ILP];
        __nv_bfloat16* dp[LL_ILP];
        int len[LL_ILP];
        bool on[LL_ILP];
#pragma unroll
        for (int u = 0; u < LL_ILP; ++u) {
            const unsigned t = t0 + u * stride;
            const unsigned row = t / gpr;
            const int cap = static_cast<int>(t - row * gpr) * CPW + sub;
            on[u] = t < total && cap < n_caps && w0 < Wp;
            len[u] = (t < total && cap < n_caps) ? min(__ldg(cap_len + cap), Wp) : 0;
            const long long col = static_cast<long long>(cap) * Wp + w0;
            dp[u] = dE + static_cast<long long>(row) * ld_d + col;
            ev[u] = make_uint2(0u, 0u); dv[u] = make_uint2(0u, 0u);
            if (on[u]) {
                ev[u] = *reinterpret_cast<const uint2*>(E + static_cast<long long>(row) * ld_e + col);
                dv[u] = *reinterpret_cast<const uint2*>(dp[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < LL_ILP; ++u) {
            if (t0 + u * stride >= total) break;                 // warp-uniform
            const float ef[4] = {__uint_as_float(ev[u].x << 16), __uint_as_float(ev[u].x & 0xffff0000u),
                                 __uint_as_float(ev[u].y << 16), __uint_as_float(ev[u].y & 0xffff0000u)};
            const float df[4] = {__uint_as_float(dv[u].x << 16), __uint_as_float(dv[u].x & 0xffff0000u),
                                 __uint_as_float(dv[u].y << 16), __uint_as_float(dv[u].y & 0xffff0000u)};
            float a[4], g[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool m = on[u] && w0 + k < len[u];
                a[k] = m ? __logf(ef[k]) * inv_t : 0.f;
                g[k] = m ? df[k] * temp1 * ef[k] : 0.f;
            }
            const float dot = group_sum<G>((a[0] * g[0] + a[1] * g[1]) + (a[2] * g[2] + a[3] * g[3]));
            if (on[u]) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(a[0] * (g[0] - dot), a[1] * (g[1] - dot));
                const __nv_bfloat162 hi = __floats2bfloat162_rn(a[2] * (g[2] - dot), a[3] * (g[3] - dot));
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dp[u]) = pk;
            }
        }
    }
}

// Forward pass with eight words per lane (two 16-byte loads of S, one 16-byte store of E): captions that are not a power of two
// long waste fewer lanes (80 word slots: 10 of 16 lanes instead of 20 of 32) and the instruction count per byte halves.
template <int G>
__global__ void __launch_bounds__(256)
ll_softmax_exp_fwd8_kernel(const float* __restrict__ S, long long ld_s, __nv_bfloat16* __restrict__ E, long long ld_e,
                           unsigned rows, int n_caps, int Wp, const int* __restrict__ cap_len, float temp1) {
    constexpr int CPW = 32 / G;
    constexpr int ILP = 2;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, w0 = (lane % G) * 8;
    const unsigned gpr = (n_caps + CPW - 1) / CPW;
    const unsigned total = rows * gpr;
    const unsigned stride = gridDim.x * (blockDim.x >> 5);
    for (unsigned t0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t0 < total; t0 += ILP * stride) {
        float v[ILP][8];
        long long off[ILP];
        int len[ILP];
        bool on[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const unsigned t = t0 + u * stride;
            const unsigned row = t / gpr;
            const int cap = static_cast<int>(t - row * gpr) * CPW + sub;
            on[u] = t < total && cap < n_caps && w0 < Wp;
            len[u] = (t < total && cap < n_caps) ? min(__ldg(cap_len + cap), Wp) : 0;
            off[u] = static_cast<long long>(cap) * Wp + w0;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[u][k] = -INFINITY;
            if (on[u]) {
                const float4* sp = reinterpret_cast<const float4*>(S + static_cast<long long>(row) * ld_s + off[u]);
                const float4 lo = sp[0], hi = sp[1];
                v[u][0] = lo.x; v[u][1] = lo.y; v[u][2] = lo.z; v[u][3] = lo.w;
                v[u][4] = hi.x; v[u][5] = hi.y; v[u][6] = hi.z; v[u][7] = hi.w;
            }
            off[u] += static_cast<long long>(row) * ld_e;
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            if (t0 + u * stride >= total) break;                 // warp-uniform
            const int ln = len[u];
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 8; ++k) mx = fmaxf(mx, (w0 + k < ln) ? v[u][k] : -INFINITY);
            mx = group_max<G>(mx);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[u][k] = (w0 + k < ln) ? __expf(v[u][k] - mx) : 0.f;
                sum += v[u][k];
            }
            sum = group_sum<G>(sum);
            const float sc = ln > 0 ? temp1 / sum : 0.f;
            if (on[u]) {
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __nv_bfloat162 p = __floats2bfloat162_rn((w0 + 2 * k < ln) ? __expf(v[u][2 * k] * sc) : 0.f,
                                                                   (w0 + 2 * k + 1 < ln) ? __expf(v[u][2 * k + 1] * sc) : 0.f);
                    o[k] = *reinterpret_cast<const uint32_t*>(&p);
                }
                *reinterpret_cast<uint4*>(E + off[u]) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// Same pass with eight words per lane (one 16-byte load of E and of dE): half the instructions per byte of the 4-word version.
template <int G>
__global__ void __launch_bounds__(256)
ll_softmax_exp_bwd8_kernel(const __nv_bfloat16* __restrict__ E, long long ld_e, __nv_bfloat16* __restrict__ dE, long long ld_d,
                           unsigned rows, int n_caps, int Wp, const int* __restrict__ cap_len, float temp1) {
    constexpr int CPW = 32 / G;
    constexpr int ILP = 2;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, w0 = (lane % G) * 8;
    const unsigned gpr = (n_caps + CPW - 1) / CPW;
    const unsigned total = rows * gpr;
    const unsigned stride = gridDim.x * (blockDim.x >> 5);
    const float inv_t = 1.0f / temp1;
    for (unsigned t0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t0 < total; t0 += ILP * stride) {
        uint4 ev[ILP], dv[ILP];
        __nv_bfloat16* dp[ILP];
        int len[ILP];
        bool on[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const unsigned t = t0 + u * stride;
            const unsigned row = t / gpr;
            const int cap = static_cast<int>(t - row * gpr) * CPW + sub;
            on[u] = t < total && cap < n_caps && w0 < Wp;
            len[u] = (t < total && cap < n_caps) ? min(__ldg(cap_len + cap), Wp) : 0;
            const long long col = static_cast<long long>(cap) * Wp + w0;
            dp[u] = dE + static_cast<long long>(row) * ld_d + col;
            ev[u] = make_uint4(0u, 0u, 0u, 0u); dv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (on[u]) {
                ev[u] = *reinterpret_cast<const uint4*>(E + static_cast<long long>(row) * ld_e + col);
                dv[u] = *reinterpret_cast<const uint4*>(dp[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            if (t0 + u * stride >= total) break;                 // warp-uniform
            const uint32_t ew[4] = {ev[u].x, ev[u].y, ev[u].z, ev[u].w};
            const uint32_t dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
            float a[8], g[8];
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float ef = (k & 1) ? __uint_as_float(ew[k >> 1] & 0xffff0000u) : __uint_as_float(ew[k >> 1] << 16);
                const float df = (k & 1) ? __uint_as_float(dw[k >> 1] & 0xffff0000u) : __uint_as_float(dw[k >> 1] << 16);
                const bool m = on[u] && w0 + k < len[u];
                a[k] = m ? __logf(ef) * inv_t : 0.f;
                g[k] = m ? df * temp1 * ef : 0.f;
                dot = fmaf(a[k], g[k], dot);
            }
            dot = group_sum<G>(dot);
            if (on[u]) {
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __nv_bfloat162 p = __floats2bfloat162_rn(a[2 * k] * (g[2 * k] - dot), a[2 * k + 1] * (g[2 * k + 1] - dot));
                    o[k] = *reinterpret_cast<const uint32_t*>(&p);
                }
                *reinterpret_cast<uint4*>(dp[u]) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// one warp per (image b, caption i): cosine of every word with its attended context, log-sum-exp over the words
__global__ void __launch_bounds__(256)
ll_cos_lse_fwd_kernel(const float* __restrict__ wcU, const float* __restrict__ words, int B, int n_caps, int Wp, int D,
                      const int* __restrict__ cap_len, float temp2, int agg_mean, float eps, float* __restrict__ cosv,
                      float* __restrict__ sim, long long ld_sim) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * n_caps) return;
    const int b = warp / n_caps, cap = warp - b * n_caps;
    const int len = min(cap_len[cap], Wp);
    const long long N = static_cast<long long>(n_caps) * Wp;
    float m = 0.f;
    for (int w = 0; w < Wp; ++w) {
        const long long n = static_cast<long long>(cap) * Wp + w;
        float c = 0.f;
        if (w < len) {
            const float* x = words + n * D;
            const float* y = wcU + (static_cast<long long>(b) * N + n) * D;
            float xy = 0.f, xx = 0.f, yy = 0.f;
            for (int d = lane; d < D; d += 32) {
                const float xv = x[d], yv = y[d];
                xy = fmaf(xv, yv, xy); xx = fmaf(xv, xv, xx); yy = fmaf(yv, yv, yy);
            }
            xy = warp_sum_f(xy); xx = warp_sum_f(xx); yy = warp_sum_f(yy);
            c = xy / fmaxf(sqrtf(xx) * sqrtf(yy), eps);                      // cosine_similarity, losses.py:690-696
            m += __expf(temp2 * c);
        }
        if (lane == 0) cosv[static_cast<long long>(b) * N + n] = c;
    }
    if (lane == 0) {
        float v = -INFINITY;
        if (len > 0) v = logf(agg_mean ? m / static_cast<float>(len) : m);
        sim[static_cast<long long>(b) * ld_sim + cap] = v;
    }
}

// one warp per column n = (caption, word), looping over the images: d cos -> d wcU[b, n, :] (bf16) and the direct part of d word_n
// A block of 16 warps takes 16 consecutive columns, so that the transposed copy dwcUT[b, d, n] (the K-major weight of the
// d ctx GEMM) can be written in full 32-byte sectors through a shared-memory tile.
constexpr int LL_BWD_WARPS = 16;
template <int DPL>   // D / 32 elements per lane kept in registers
__global__ void __launch_bounds__(LL_BWD_WARPS * 32)
ll_cos_lse_bwd_kernel(const float* __restrict__ dsim, long long ld_dsim, const float* __restrict__ sim, long long ld_sim,
                      const float* __restrict__ cosv, const float* __restrict__ wcU, const float* __restrict__ words, int B,
                      int n_caps, int Wp, const int* __restrict__ cap_len, float temp2, int agg_mean, float eps,
                      __nv_bfloat16* __restrict__ dwcU, __nv_bfloat16* __restrict__ dwcUT, long long ld_t,
                      float* __restrict__ dwords) {
    constexpr int D = DPL * 32;
    __shared__ __align__(16) __nv_bfloat16 tile[LL_BWD_WARPS][D];      // [n - n0][d]: a warp's stores are contiguous (no bank conflicts)
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long N = static_cast<long long>(n_caps) * Wp;
    const long long n0 = static_cast<long long>(blockIdx.x) * LL_BWD_WARPS;
    const long long n = n0 + wib;                                      // N is a multiple of 16: always in range
    const int cap = static_cast<int>(n / Wp), w = static_cast<int>(n - static_cast<long long>(cap) * Wp);
    const int len = min(cap_len[cap], Wp);
    const bool live = w < len;
    float x[DPL], dx[DPL];
    float xx = 0.f;
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
        x[k] = live ? words[n * D + lane + 32 * k] : 0.f;
        dx[k] = 0.f;
        xx = fmaf(x[k], x[k], xx);
    }
    xx = warp_sum_f(xx);
    const float nx = sqrtf(xx);
    const float log_len = agg_mean && len > 0 ? logf(static_cast<float>(len)) : 0.f;
    for (int b = 0; b < B; ++b) {
        __nv_bfloat16* out = dwcU + (static_cast<long long>(b) * N + n) * D;
        float o[DPL];
#pragma unroll
        for (int k = 0; k < DPL; ++k) o[k] = 0.f;
        if (live) {
            const float* yp = wcU + (static_cast<long long>(b) * N + n) * D;
            float y[DPL];
            float xy = 0.f, yy = 0.f;
#pragma unroll
            for (int k = 0; k < DPL; ++k) {
                y[k] = yp[lane + 32 * k];
                xy = fmaf(x[k], y[k], xy); yy = fmaf(y[k], y[k], yy);
            }
            xy = warp_sum_f(xy); yy = warp_sum_f(yy);
            const float ny = sqrtf(yy);
            const float den = nx * ny;
            const float c = cosv[static_cast<long long>(b) * N + n];
            // sim = log(sum_w exp(temp2 cos)) [- log len]:  d sim / d cos = temp2 exp(temp2 cos - log sum)
            const float logm = sim[static_cast<long long>(b) * ld_sim + cap] + log_len;
            const float dc = dsim[static_cast<long long>(b) * ld_dsim + cap] * temp2 * __expf(temp2 * c - logm);
            float gy_x = 0.f, gy_y = 0.f, gx_y = 0.f, gx_x = 0.f;     // d cos / d y = gy_x x + gy_y y ; d cos / d x = gx_y y + gx_x x
            if (den > eps) {                                            // below the clamp the denominator is constant
                const float r = 1.0f / den;
                gy_x = r; gy_y = -c / fmaxf(yy, 1e-30f);
                gx_y = r; gx_x = -c / fmaxf(xx, 1e-30f);
            } else {
                gy_x = 1.0f / eps; gx_y = 1.0f / eps;
            }
#pragma unroll
            for (int k = 0; k < DPL; ++k) {
                o[k] = dc * (gy_x * x[k] + gy_y * y[k]);
                dx[k] = fmaf(dc, gx_y * y[k] + gx_x * x[k], dx[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
            const __nv_bfloat16 v = __float2bfloat16(o[k]);
            out[lane + 32 * k] = v;
            tile[wib][lane + 32 * k] = v;
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += LL_BWD_WARPS * 32) {     // column d of the tile -> 32 contiguous bytes of row (b, d)
            uint32_t w[LL_BWD_WARPS / 2];
#pragma unroll
            for (int j = 0; j < LL_BWD_WARPS / 2; ++j)
                w[j] = static_cast<uint32_t>(__bfloat16_as_ushort(tile[2 * j][d])) |
                       (static_cast<uint32_t>(__bfloat16_as_ushort(tile[2 * j + 1][d])) << 16);
            uint4* dst = reinterpret_cast<uint4*>(dwcUT + (static_cast<long long>(b) * D + d) * ld_t + n0);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < DPL; ++k) dwords[n * D + lane + 32 * k] = dx[k];
}

}  // namespace
}  // namespace mm

using namespace mm;

extern "C" int mm_local_softmax_exp_fwd(const float* S, long long ld_s, void* E, long long ld_e, long long rows, int n_caps,
                                        int Wp, const int32_t* cap_len, float temp1, void* stream) {
    MM_REQUIRE(S && E && cap_len && rows >= 0 && n_caps > 0 && Wp > 0, MM_ERR_BAD_SHAPE, "mm_local_softmax_exp_fwd: bad arguments");
    MM_REQUIRE(ld_s >= static_cast<long long>(n_caps) * Wp && ld_e >= static_cast<long long>(n_caps) * Wp, MM_ERR_BAD_SHAPE,
               "mm_local_softmax_exp_fwd: leading dimensions must cover n_caps * Wp columns");
    const long long total = rows * n_caps;
    if (total == 0) return MM_OK;
    MM_REQUIRE(Wp <= 128 && Wp % 8 == 0 && ld_s % 4 == 0 && ld_e % 4 == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(E) & 7) == 0,
               MM_ERR_UNSUPPORTED, "mm_local_softmax_exp_fwd: Wp must be a multiple of 8 up to 128, rows 16-byte aligned");
    MM_REQUIRE(rows * (n_caps + 1) < (1LL << 32), MM_ERR_UNSUPPORTED, "mm_local_softmax_exp_fwd: too many (row, caption) pairs");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* Eb = static_cast<__nv_bfloat16*>(E);
    const unsigned grid = 148 * 16, r = static_cast<unsigned>(rows);
    const int lanes = (Wp + 3) / 4;
    const bool vec8 = ld_s % 8 == 0 && ld_e % 8 == 0 && (reinterpret_cast<uintptr_t>(S) & 31) == 0 && (reinterpret_cast<uintptr_t>(E) & 15) == 0;
    if (vec8) {      // Wp is a multiple of 8: every caption starts on a 32-byte (S) / 16-byte (E) boundary
        const int l8 = Wp / 8;
        if (l8 <= 1) ll_softmax_exp_fwd8_kernel<1><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
        else if (l8 <= 2) ll_softmax_exp_fwd8_kernel<2><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
        else if (l8 <= 4) ll_softmax_exp_fwd8_kernel<4><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
        else if (l8 <= 8) ll_softmax_exp_fwd8_kernel<8><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
        else ll_softmax_exp_fwd8_kernel<16><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
    } else
    if (lanes <= 2) ll_softmax_exp_fwd_kernel<2><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
    else if (lanes <= 4) ll_softmax_exp_fwd_kernel<4><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
    else if (lanes <= 8) ll_softmax_exp_fwd_kernel<8><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
    else if (lanes <= 16) ll_softmax_exp_fwd_kernel<16><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
    else ll_softmax_exp_fwd_kernel<32><<<grid, 256, 0, st>>>(S, ld_s, Eb, ld_e, r, n_caps, Wp, cap_len, temp1);
    note_launches(1);
    return check_launch("mm_local_softmax_exp_fwd");
}

extern "C" int mm_local_softmax_exp_bwd(const void* E, long long ld_e, void* dE, long long ld_d, long long rows, int n_caps,
                                        int Wp, const int32_t* cap_len, float temp1, void* stream) {
    MM_REQUIRE(E && dE && cap_len && rows >= 0 && n_caps > 0 && Wp > 0, MM_ERR_BAD_SHAPE, "mm_local_softmax_exp_bwd: bad arguments");
    const long long total = rows * n_caps;
    if (total == 0) return MM_OK;
    MM_REQUIRE(Wp <= 128 && Wp % 8 == 0 && ld_e % 4 == 0 && ld_d % 4 == 0 && (reinterpret_cast<uintptr_t>(E) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(dE) & 7) == 0,
               MM_ERR_UNSUPPORTED, "mm_local_softmax_exp_bwd: Wp must be a multiple of 8 up to 128, rows 8-byte aligned");
    MM_REQUIRE(rows * (n_caps + 1) < (1LL << 32), MM_ERR_UNSUPPORTED, "mm_local_softmax_exp_bwd: too many (row, caption) pairs");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* Eb = static_cast<const __nv_bfloat16*>(E);
    __nv_bfloat16* Db = static_cast<__nv_bfloat16*>(dE);
    const unsigned grid = 148 * 16, r = static_cast<unsigned>(rows);
    const int lanes = (Wp + 3) / 4;
    const bool vec8 = ld_e % 8 == 0 && ld_d % 8 == 0 && ((reinterpret_cast<uintptr_t>(E) | reinterpret_cast<uintptr_t>(dE)) & 15) == 0;
    if (vec8) {      // Wp is a multiple of 8: every caption starts on a 16-byte boundary
        const int l8 = Wp / 8;
        if (l8 <= 1) ll_softmax_exp_bwd8_kernel<1><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
        else if (l8 <= 2) ll_softmax_exp_bwd8_kernel<2><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
        else if (l8 <= 4) ll_softmax_exp_bwd8_kernel<4><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
        else if (l8 <= 8) ll_softmax_exp_bwd8_kernel<8><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
        else ll_softmax_exp_bwd8_kernel<16><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
    } else
    if (lanes <= 2) ll_softmax_exp_bwd_kernel<2><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
    else if (lanes <= 4) ll_softmax_exp_bwd_kernel<4><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
    else if (lanes <= 8) ll_softmax_exp_bwd_kernel<8><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
    else if (lanes <= 16) ll_softmax_exp_bwd_kernel<16><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
    else ll_softmax_exp_bwd_kernel<32><<<grid, 256, 0, st>>>(Eb, ld_e, Db, ld_d, r, n_caps, Wp, cap_len, temp1);
    note_launches(1);
    return check_launch("mm_local_softmax_exp_bwd");
}

extern "C" int mm_local_cos_lse_fwd(const float* wcU, const float* words, int B, int n_caps, int Wp, int D,
                                    const int32_t* cap_len, float temp2, int agg_mean, float* cosv, float* sim,
                                    long long ld_sim, void* stream) {
    MM_REQUIRE(wcU && words && cap_len && cosv && sim && B > 0 && n_caps > 0 && Wp > 0 && D > 0, MM_ERR_BAD_SHAPE,
               "mm_local_cos_lse_fwd: bad arguments");
    const long long warps = static_cast<long long>(B) * n_caps;
    ll_cos_lse_fwd_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        wcU, words, B, n_caps, Wp, D, cap_len, temp2, agg_mean, 1e-8f, cosv, sim, ld_sim);
    note_launches(1);
    return check_launch("mm_local_cos_lse_fwd");
}

extern "C" int mm_local_cos_lse_bwd(const float* dsim, long long ld_dsim, const float* sim, long long ld_sim, const float* cosv,
                                    const float* wcU, const float* words, int B, int n_caps, int Wp, int D,
                                    const int32_t* cap_len, float temp2, int agg_mean, void* dwcU, void* dwcUT,
                                    long long ld_t, float* dwords, void* stream) {
    MM_REQUIRE(dsim && sim && cosv && wcU && words && cap_len && dwcU && dwcUT && dwords && B > 0 && n_caps > 0 && Wp > 0,
               MM_ERR_BAD_SHAPE, "mm_local_cos_lse_bwd: bad arguments");
    const long long warps = static_cast<long long>(n_caps) * Wp;
    MM_REQUIRE(warps % LL_BWD_WARPS == 0 && ld_t % 8 == 0 && (reinterpret_cast<uintptr_t>(dwcUT) & 15) == 0, MM_ERR_BAD_SHAPE,
               "mm_local_cos_lse_bwd: n_caps * Wp must be a multiple of 16, dwcUT rows 16-byte aligned");
    const unsigned grid = static_cast<unsigned>(warps / LL_BWD_WARPS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MM_LL_CASE(dpl)                                                                                                  \
    case dpl * 32:                                                                                                       \
        ll_cos_lse_bwd_kernel<dpl><<<grid, LL_BWD_WARPS * 32, 0, st>>>(                                                  \
            dsim, ld_dsim, sim, ld_sim, cosv, wcU, words, B, n_caps, Wp, cap_len, temp2, agg_mean, 1e-8f,                \
            static_cast<__nv_bfloat16*>(dwcU), static_cast<__nv_bfloat16*>(dwcUT), ld_t, dwords);                        \
        break;
    switch (D) {
        MM_LL_CASE(24)
        MM_LL_CASE(16)
        MM_LL_CASE(8)
        MM_LL_CASE(4)
        default:
            set_error("mm_local_cos_lse_bwd: embedding width %d not built (768, 512, 256, 128)", D);
            return MM_ERR_UNSUPPORTED;
    }
#undef MM_LL_CASE
    note_launches(1);
    return check_launch("mm_local_cos_lse_bwd");
}
