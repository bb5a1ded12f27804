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

// one thread per (row, caption): a caption's Wp scores are contiguous, neighbouring threads take neighbouring captions
__global__ void __launch_bounds__(256)
ll_softmax_exp_fwd_kernel(const float* __restrict__ S, long long ld_s, __nv_bfloat16* __restrict__ E, long long ld_e,
                          long long rows, int n_caps, int Wp, const int* __restrict__ cap_len, float temp1) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= rows * n_caps) return;
    const long long row = t / n_caps;
    const int cap = static_cast<int>(t - row * n_caps);
    const int len = min(cap_len[cap], Wp);
    const float* s = S + row * ld_s + static_cast<long long>(cap) * Wp;
    __nv_bfloat16* e = E + row * ld_e + static_cast<long long>(cap) * Wp;
    float mx = -INFINITY;
    for (int w = 0; w < len; ++w) mx = fmaxf(mx, s[w]);
    float sum = 0.f;
    for (int w = 0; w < len; ++w) sum += __expf(s[w] - mx);
    const float inv = len > 0 ? 1.0f / sum : 0.f;
    for (int w = 0; w < Wp; ++w) {
        float v = 0.f;
        if (w < len) v = __expf(temp1 * __expf(s[w] - mx) * inv);
        e[w] = __float2bfloat16(v);
    }
}

// dE (in place -> dS): a = log(E) / temp1 is the first softmax's output, g = dE * temp1 * E, dS = a (g - sum_w a g)
__global__ void __launch_bounds__(256)
ll_softmax_exp_bwd_kernel(const __nv_bfloat16* __restrict__ E, long long ld_e, __nv_bfloat16* __restrict__ dE, long long ld_d,
                          long long rows, int n_caps, int Wp, const int* __restrict__ cap_len, float temp1) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= rows * n_caps) return;
    const long long row = t / n_caps;
    const int cap = static_cast<int>(t - row * n_caps);
    const int len = min(cap_len[cap], Wp);
    const __nv_bfloat16* e = E + row * ld_e + static_cast<long long>(cap) * Wp;
    __nv_bfloat16* d = dE + row * ld_d + static_cast<long long>(cap) * Wp;
    const float inv_t = 1.0f / temp1;
    float dot = 0.f;
    for (int w = 0; w < len; ++w) {
        const float ev = __bfloat162float(e[w]);
        const float a = __logf(ev) * inv_t;
        dot += a * (__bfloat162float(d[w]) * temp1 * ev);
    }
    for (int w = 0; w < Wp; ++w) {
        float v = 0.f;
        if (w < len) {
            const float ev = __bfloat162float(e[w]);
            const float a = __logf(ev) * inv_t;
            v = a * (__bfloat162float(d[w]) * temp1 * ev - dot);
        }
        d[w] = __float2bfloat16(v);
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
template <int DPL>   // D / 32 elements per lane kept in registers
__global__ void __launch_bounds__(256)
ll_cos_lse_bwd_kernel(const float* __restrict__ dsim, long long ld_dsim, const float* __restrict__ sim, long long ld_sim,
                      const float* __restrict__ cosv, const float* __restrict__ wcU, const float* __restrict__ words, int B,
                      int n_caps, int Wp, const int* __restrict__ cap_len, float temp2, int agg_mean, float eps,
                      __nv_bfloat16* __restrict__ dwcU, float* __restrict__ dwords) {
    constexpr int D = DPL * 32;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long long N = static_cast<long long>(n_caps) * Wp;
    if (warp >= N) return;
    const long long n = warp;
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
        if (!live) {
#pragma unroll
            for (int k = 0; k < DPL; ++k) out[lane + 32 * k] = __float2bfloat16(0.f);
            continue;
        }
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
            out[lane + 32 * k] = __float2bfloat16(dc * (gy_x * x[k] + gy_y * y[k]));
            dx[k] = fmaf(dc, gx_y * y[k] + gx_x * x[k], dx[k]);
        }
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
    ll_softmax_exp_fwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        S, ld_s, static_cast<__nv_bfloat16*>(E), ld_e, rows, n_caps, Wp, cap_len, temp1);
    note_launches(1);
    return check_launch("mm_local_softmax_exp_fwd");
}

extern "C" int mm_local_softmax_exp_bwd(const void* E, long long ld_e, void* dE, long long ld_d, long long rows, int n_caps,
                                        int Wp, const int32_t* cap_len, float temp1, void* stream) {
    MM_REQUIRE(E && dE && cap_len && rows >= 0 && n_caps > 0 && Wp > 0, MM_ERR_BAD_SHAPE, "mm_local_softmax_exp_bwd: bad arguments");
    const long long total = rows * n_caps;
    if (total == 0) return MM_OK;
    ll_softmax_exp_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(E), ld_e, static_cast<__nv_bfloat16*>(dE), ld_d, rows, n_caps, Wp, cap_len, temp1);
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
                                    const int32_t* cap_len, float temp2, int agg_mean, void* dwcU, float* dwords,
                                    void* stream) {
    MM_REQUIRE(dsim && sim && cosv && wcU && words && cap_len && dwcU && dwords && B > 0 && n_caps > 0 && Wp > 0, MM_ERR_BAD_SHAPE,
               "mm_local_cos_lse_bwd: bad arguments");
    const long long warps = static_cast<long long>(n_caps) * Wp;
    const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MM_LL_CASE(dpl)                                                                                                  \
    case dpl * 32:                                                                                                       \
        ll_cos_lse_bwd_kernel<dpl><<<grid, 256, 0, st>>>(dsim, ld_dsim, sim, ld_sim, cosv, wcU, words, B, n_caps, Wp, cap_len,  \
                                                         temp2, agg_mean, 1e-8f, static_cast<__nv_bfloat16*>(dwcU), dwords); \
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
