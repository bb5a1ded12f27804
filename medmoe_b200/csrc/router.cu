// medmoe_b200 — router gate (north-star kernel 1).
//
// Forward restates reference swin.py:98-100:
//     probs = softmax(Linear(128,K)(ReLU(Linear(768,128)(swin_feat)))) ; top = argmax(probs)
// entirely in fp32 (argmax parity: ties break to the lowest index like torch.argmax).
// top-k (k > 1) is the documented extension: k largest probs, gate weights =
// those probs renormalised to sum 1; k = 1 gives weight 1.0, i.e. the reference gather.
//
// Backward is the gradient of the returned probabilities (the only path into the router
// in the reference: F.cross_entropy(probs, label), medmoe_module.py:235-237).
#include "mm_common.cuh"
#include "api_internal.h"

namespace mm {

constexpr int ROUTER_HID = 128;   // fixed by the reference (swin.py:88-90)
constexpr int ROUTER_MAX_K = 64;

// One CTA (512 threads = 16 warps) per image: the 128 hidden units are latency-bound dot products over W1 rows, so the
// more warps share them the shorter each warp's serial chain of L2 round trips.
constexpr int ROUTER_FWD_THREADS = 512;
__global__ void __launch_bounds__(ROUTER_FWD_THREADS)
router_fwd_kernel(const float* __restrict__ x, int D, const float* __restrict__ W1, const float* __restrict__ b1,
                  const float* __restrict__ W2, const float* __restrict__ b2, int K, int topk,
                  float* __restrict__ hidden, float* __restrict__ probs, int* __restrict__ topk_idx,
                  float* __restrict__ topk_w, int* __restrict__ near_tie, float tie_tol) {
    extern __shared__ float sm[];
    float* sx = sm;                 // [D]
    float* sh = sm + D;             // [128]
    float* sl = sh + ROUTER_HID;    // [K]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < D; i += blockDim.x) sx[i] = x[static_cast<size_t>(b) * D + i];
    __syncthreads();
    // hidden: each warp owns 32 outputs; lanes stride the 768-long dot (coalesced W1 reads).
    // Four outputs at a time: four independent load -> FMA chains per lane keep the L2 latency of W1 covered.
    constexpr int PER_WARP = ROUTER_HID / (ROUTER_FWD_THREADS / 32);
    for (int o = warp * PER_WARP; o < warp * PER_WARP + PER_WARP; o += 4) {
        const float* w = W1 + static_cast<size_t>(o) * D;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int i = lane; i < D; i += 32) {
            const float xv = sx[i];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmaf(w[static_cast<size_t>(u) * D + i], xv, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float v = warp_sum(acc[u]);
            if (lane == 0) {
                const float pre = v + b1[o + u];
                const float h = pre < 0.f ? 0.f : pre;        // ReLU that lets NaN through like torch's (fmaxf would drop it)
                sh[o + u] = h;
                hidden[static_cast<size_t>(b) * ROUTER_HID + o + u] = h;
            }
        }
    }
    __syncthreads();
    for (int e = warp; e < K; e += ROUTER_FWD_THREADS / 32) {
        const float* w = W2 + static_cast<size_t>(e) * ROUTER_HID;
        float acc = 0.f;
        for (int i = lane; i < ROUTER_HID; i += 32) acc = fmaf(w[i], sh[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) sl[e] = acc + b2[e];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = sl[0];
        for (int e = 1; e < K; ++e) mx = (sl[e] > mx || sl[e] != sl[e]) ? sl[e] : mx;      // NaN-propagating maximum
        float sum = 0.f;
        for (int e = 0; e < K; ++e) { const float t = expf(sl[e] - mx); sl[e] = t; sum += t; }
        const float inv = 1.0f / sum;
        for (int e = 0; e < K; ++e) {
            sl[e] *= inv;
            probs[static_cast<size_t>(b) * K + e] = sl[e];
        }
        // top-k by repeated first-max selection (k is 1 or 2 in every configuration).  NaN counts as the maximum and the
        // first one wins, exactly like torch.argmax / torch.topk: a NaN / Inf in swin_feat makes every probability NaN,
        // the image goes to expert 0 and the NaN propagates to the outputs and the loss as in the reference (swin.py:99-100)
        // instead of leaving the selection empty.
        unsigned long long taken = 0ull;
        float wsum = 0.f;
        for (int j = 0; j < topk; ++j) {
            int best = -1; float bv = 0.f;
            for (int e = 0; e < K; ++e) {
                if ((taken >> e) & 1ull) continue;
                const float v = sl[e];
                if (best < 0 || v > bv || (v != v && bv == bv)) { bv = v; best = e; }
            }
            taken |= 1ull << best;
            topk_idx[static_cast<size_t>(b) * topk + j] = best;
            topk_w[static_cast<size_t>(b) * topk + j] = bv;
            wsum += bv;
        }
        if (near_tie) {
            // near-tie report (north star: "near-tie logits documented"): the gap between the last selected probability and
            // the best one left out; below tie_tol a bf16 / reordered evaluation of the router could pick another expert.
            float last = topk_w[static_cast<size_t>(b) * topk + topk - 1], runner = -1.f;
            for (int e = 0; e < K; ++e)
                if (!((taken >> e) & 1ull) && sl[e] > runner) runner = sl[e];
            near_tie[b] = (K > topk && !(last - runner >= tie_tol)) ? 1 : 0;
        }
        for (int j = 0; j < topk; ++j)
            topk_w[static_cast<size_t>(b) * topk + j] = (topk == 1) ? 1.0f : topk_w[static_cast<size_t>(b) * topk + j] / wsum;
    }
}

// Per image: dlogit = p * (dp - <p, dp>); dh = (dlogit W2) * [h > 0]   (dx = dh W1 and dW1 = dh^T x are sgemms).
__global__ void __launch_bounds__(128)
router_bwd_sample_kernel(const float* __restrict__ dprobs, const float* __restrict__ probs,
                         const float* __restrict__ hidden, const float* __restrict__ W1,
                         const float* __restrict__ W2, int D, int K, float* __restrict__ dlogit,
                         float* __restrict__ dhidden) {
    __shared__ float sdl[ROUTER_MAX_K];
    __shared__ float sdh[ROUTER_HID];
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        float dot = 0.f;
        for (int e = 0; e < K; ++e) dot += probs[b * K + e] * dprobs[b * K + e];
        for (int e = 0; e < K; ++e) {
            const float d = probs[b * K + e] * (dprobs[b * K + e] - dot);
            sdl[e] = d;
            dlogit[b * K + e] = d;
        }
    }
    __syncthreads();
    {
        const int o = threadIdx.x;   // 128 hidden units
        float acc = 0.f;
        for (int e = 0; e < K; ++e) acc = fmaf(sdl[e], W2[e * ROUTER_HID + o], acc);
        const float dh = hidden[static_cast<size_t>(b) * ROUTER_HID + o] > 0.f ? acc : 0.f;
        sdh[o] = dh;
        dhidden[static_cast<size_t>(b) * ROUTER_HID + o] = dh;
    }
}

// dW2[e, o] = sum_b dlogit[b, e] hidden[b, o], db2[e] = sum_b dlogit[b, e]   (blocks 0 .. K-1, thread = hidden unit o)
// db1[o]    = sum_b dhidden[b, o]                                            (block K)
// coalesced over o, batch reduction in a fixed order (deterministic); replaces two latency-bound launches of batch_outer_kernel.
constexpr int ROUTER_SMALL_SLICES = 8;      // batch slices per block: 8 x 128 threads, eight loads in flight per thread
__global__ void __launch_bounds__(ROUTER_HID * ROUTER_SMALL_SLICES)
router_bwd_small_kernel(const float* __restrict__ dlogit, const float* __restrict__ hidden, const float* __restrict__ dhidden,
                        int B, int K, float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ db1) {
    __shared__ float s_acc[ROUTER_SMALL_SLICES][ROUTER_HID];
    __shared__ float s_accb[ROUTER_SMALL_SLICES];
    const int o = threadIdx.x, sl = threadIdx.y, e = blockIdx.x;
    // slice sl takes the rows b = sl, sl + SLICES, ...; partial sums are combined in slice order (deterministic)
    float acc = 0.f, accb = 0.f;
    if (e < K) {
        int b = sl;
        for (; b + 7 * ROUTER_SMALL_SLICES < B; b += 8 * ROUTER_SMALL_SLICES) {
            float l[8], h[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                l[u] = __ldg(dlogit + static_cast<size_t>(b + u * ROUTER_SMALL_SLICES) * K + e);
                h[u] = __ldg(hidden + static_cast<size_t>(b + u * ROUTER_SMALL_SLICES) * ROUTER_HID + o);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc = fmaf(l[u], h[u], acc); accb += l[u]; }
        }
        for (; b < B; b += ROUTER_SMALL_SLICES) {
            const float l = __ldg(dlogit + static_cast<size_t>(b) * K + e);
            acc = fmaf(l, __ldg(hidden + static_cast<size_t>(b) * ROUTER_HID + o), acc);
            accb += l;
        }
    } else {
        int b = sl;
        for (; b + 7 * ROUTER_SMALL_SLICES < B; b += 8 * ROUTER_SMALL_SLICES) {
            float h[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) h[u] = __ldg(dhidden + static_cast<size_t>(b + u * ROUTER_SMALL_SLICES) * ROUTER_HID + o);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += h[u];
        }
        for (; b < B; b += ROUTER_SMALL_SLICES) acc += __ldg(dhidden + static_cast<size_t>(b) * ROUTER_HID + o);
    }
    s_acc[sl][o] = acc;
    if (o == 0) s_accb[sl] = accb;
    __syncthreads();
    if (sl == 0) {
        float t = 0.f, tb = 0.f;
#pragma unroll
        for (int k = 0; k < ROUTER_SMALL_SLICES; ++k) { t += s_acc[k][o]; tb += s_accb[k]; }
        if (e < K) {
            dW2[e * ROUTER_HID + o] = t;
            if (o == 0) db2[e] = tb;
        } else {
            db1[o] = t;
        }
    }
}

}  // namespace mm

using namespace mm;

extern "C" int mm_router_topk(const float* x, int B, int D, const float* W1, const float* b1, const float* W2,
                              const float* b2, int K, int topk, float* hidden, float* probs, int32_t* topk_idx,
                              float* topk_w, int32_t* near_tie, float tie_tol, void* stream) {
    MM_REQUIRE(B >= 0 && D > 0 && K > 0 && K <= ROUTER_MAX_K && topk >= 1 && topk <= K, MM_ERR_BAD_SHAPE,
               "mm_router_topk: need 0 < K <= 64, 1 <= topk <= K");
    if (B == 0) return MM_OK;
    const size_t smem = (static_cast<size_t>(D) + ROUTER_HID + K) * sizeof(float);
    router_fwd_kernel<<<B, ROUTER_FWD_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(x, D, W1, b1, W2, b2, K, topk, hidden,
                                                                          probs, topk_idx, topk_w, near_tie, tie_tol);
    mm::note_launches(1);
    return mm_check_launch("mm_router_topk");
}

extern "C" int mm_router_bwd(const float* dprobs, const float* probs, const float* hidden, const float* x,
                             const float* W1, const float* W2, int B, int D, int K, float* dlogit, float* dhidden,
                             float* dx, float* dW1, float* db1, float* dW2, float* db2, void* stream) {
    MM_REQUIRE(B > 0 && D > 0 && K > 0 && K <= ROUTER_MAX_K, MM_ERR_BAD_SHAPE, "mm_router_bwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    router_bwd_sample_kernel<<<B, 128, 0, st>>>(dprobs, probs, hidden, W1, W2, D, K, dlogit, dhidden);
    mm::note_launches(1);
    if (dx) {   // dx[B, D] = dh[B, 128] W1[128, D]
        SgemmArgs g{};
        g.A = dhidden; g.sam = ROUTER_HID; g.sak = 1; g.B = W1; g.sbk = D; g.sbn = 1; g.C = dx; g.ldc = D;
        g.M = B; g.N = D; g.K = ROUTER_HID; g.alpha = 1.f;
        if (int rc = run_sgemm(g, st, "mm_router_bwd(dx)")) return rc;
    }
    {           // dW1[128, D] = dh^T x   (batch reduction in a fixed order)
        SgemmArgs g{};
        g.A = dhidden; g.sam = 1; g.sak = ROUTER_HID; g.B = x; g.sbk = D; g.sbn = 1; g.C = dW1; g.ldc = D;
        g.M = ROUTER_HID; g.N = D; g.K = B; g.alpha = 1.f;
        if (int rc = run_sgemm(g, st, "mm_router_bwd(dW1)")) return rc;
    }
    // dW2[K, 128] = dlogit^T h, db2 = sum_b dlogit, db1 = sum_b dh: one launch
    router_bwd_small_kernel<<<K + 1, dim3(ROUTER_HID, ROUTER_SMALL_SLICES), 0, st>>>(dlogit, hidden, dhidden, B, K, dW2, db2, db1);
    mm::note_launches(1);
    return mm_check_launch("mm_router_bwd");
}
