// medmoe_b200 — global contrastive (InfoNCE) loss kernels and the zero-shot classifier
// (north-star kernel 5).
//
// Two reference semantics share these kernels:
//   GLORIA  (losses.py:766-794)  S = temp3 * (I T^T) / max(|I||T|^T, 1e-8);  loss = CE(S) + CE(S^T)
//   FLAVA   (losses.py:527-592)  L_a = exp(logit_scale) * a b_all^T, L_b likewise; labels = B*rank + i;
//                                loss = (CE(L_a) + CE(L_b)) / 2, optional boolean row mask
// Everything here is fp32: the matrices are at most [256, 2048] x 768 per rank (~5 GFLOP
// fwd+bwd), i.e. launch-latency bound, and fp32 keeps the loss within 1e-6 of the oracle
// and the zero-shot argmax exact.
#include "mm_common.cuh"
#include "api_internal.h"

namespace mm {

// ------------------------------------------------------------------------------------
// generic strided fp32 GEMM:  C[m, n] = alpha * den(m, n) * sum_k A(m, k) B(k, n) + rs[m] * X[m, n]
//   A(m, k) = A[m * sam + k * sak],  B(k, n) = B[k * sbk + n * sbn]
//   den(m, n) = 1 / max(na[m] * nb[n], eps) when na != nullptr, else 1
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------
// T x T outputs per thread, 16 x 16 threads: tile = 16 T (64 x 64 for T = 4; 32 x 32 for T = 2, used when the
// 64-wide grid would leave most of the 148 SMs idle, e.g. the 256 x 256 single-rank logits).
template <int T>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmArgs a) {
    constexpr int TILE = 16 * T;
    __shared__ __align__(16) float As[16][TILE + 4];
    __shared__ __align__(16) float Bs[16][TILE + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * TILE, n0 = blockIdx.x * TILE;
    const int ty = t >> 4, tx = t & 15;
    float acc[T][T];
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = 0.f;

    // split-K: blockIdx.z takes one contiguous k range; partial results are added to a pre-zeroed C (run_sgemm)
    const int kc = ((a.K + 15) / 16 + gridDim.z - 1) / gridDim.z * 16;
    const int k_begin = blockIdx.z * kc, k_end = min(a.K, k_begin + kc);
    for (int k0 = k_begin; k0 < k_end; k0 += 16) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
            int m, k;
            if (a.sak == 1) { k = t & 15; m = (t >> 4) + 16 * j; } else { m = t % TILE; k = t / TILE + (256 / TILE) * j; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < a.M && gk < k_end) ? a.A[gm * a.sam + gk * a.sak] : 0.f;
            int n, kb;
            if (a.sbk == 1) { kb = t & 15; n = (t >> 4) + 16 * j; } else { n = t % TILE; kb = t / TILE + (256 / TILE) * j; }
            const int gn = n0 + n, gkb = k0 + kb;
            Bs[kb][n] = (gn < a.N && gkb < k_end) ? a.B[gkb * a.sbk + gn * a.sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float ar[T], br[T];
#pragma unroll
            for (int i = 0; i < T; ++i) { ar[i] = As[k][ty * T + i]; br[i] = Bs[k][tx * T + i]; }
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float alpha = a.alpha * (a.alpha_dev ? *a.alpha_dev : 1.0f);
#pragma unroll
    for (int i = 0; i < T; ++i) {
        const int m = m0 + ty * T + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < T; ++j) {
            const int n = n0 + tx * T + j;
            if (n >= a.N) continue;
            float v = acc[i][j];
            if (a.na) v = v / fmaxf(a.na[m] * a.nb[n], a.eps);
            v *= alpha;
            if (a.rs && blockIdx.z == 0) v = fmaf(a.rs[m], a.X[m * a.ldx + n], v);
            if (gridDim.z == 1) a.C[m * a.ldc + n] = v;
            else atomicAdd(a.C + m * a.ldc + n, v);
        }
    }
}

// n[r] = ||X[r, :]||_2 ; one warp per row
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ X, int R, int D, float* __restrict__ n) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    float acc = 0.f;
    for (int i = lane; i < D; i += 32) { const float v = X[static_cast<size_t>(r) * D + i]; acc = fmaf(v, v, acc); }
    acc = warp_sum(acc);
    if (lane == 0) n[r] = sqrtf(acc);
}

// y = x / max(||x||, eps) (F.normalize); also returns the norm. One warp per row.
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ X, int R, int D, float eps, float* __restrict__ Y, float* __restrict__ n) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    float acc = 0.f;
    for (int i = lane; i < D; i += 32) { const float v = X[static_cast<size_t>(r) * D + i]; acc = fmaf(v, v, acc); }
    acc = sqrtf(warp_sum(acc));
    const float inv = 1.0f / fmaxf(acc, eps);
    for (int i = lane; i < D; i += 32) Y[static_cast<size_t>(r) * D + i] = X[static_cast<size_t>(r) * D + i] * inv;
    if (lane == 0) n[r] = acc;
}
// dx = (dy - y <dy, y>) / max(n, eps)   (for n > eps; for n <= eps the clamp has zero slope: dx = dy / eps)
__global__ void __launch_bounds__(256) l2_normalize_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, const float* __restrict__ n, int R, int D, float eps, float* __restrict__ dX) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    float dot = 0.f;
    for (int i = lane; i < D; i += 32) dot = fmaf(dY[static_cast<size_t>(r) * D + i], Y[static_cast<size_t>(r) * D + i], dot);
    dot = warp_sum(dot);
    const float nr = n[r];
    const float inv = 1.0f / fmaxf(nr, eps);
    if (nr <= eps) dot = 0.f;
    for (int i = lane; i < D; i += 32)
        dX[static_cast<size_t>(r) * D + i] = (dY[static_cast<size_t>(r) * D + i] - Y[static_cast<size_t>(r) * D + i] * dot) * inv;
}

// row log-sum-exp + label pick: lse[r], picked[r] = L[r, label0 + r]. One warp per row.
__global__ void __launch_bounds__(256) lse_rows_kernel(const float* __restrict__ L, int R, int C, long long ld, int label0, float* __restrict__ lse, float* __restrict__ picked) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* row = L + r * ld;
    float mx = -INFINITY;
    for (int i = lane; i < C; i += 32) mx = fmaxf(mx, row[i]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < C; i += 32) s += expf(row[i] - mx);
    s = warp_sum(s);
    if (lane == 0) {
        lse[r] = mx + logf(s);
        if (picked) picked[r] = row[label0 + r];
    }
}
// column log-sum-exp: one thread per column (coalesced across threads)
__global__ void __launch_bounds__(256) lse_cols_kernel(const float* __restrict__ L, int R, int C, long long ld, float* __restrict__ lse) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mx = -INFINITY;
#pragma unroll 8
    for (int r = 0; r < R; ++r) mx = fmaxf(mx, L[r * ld + c]);
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < R; ++r) s += expf(L[r * ld + c] - mx);
    lse[c] = mx + logf(s);
}

// out[0] = sum_r w[r] * (lse[r] - picked[r])   (w == nullptr: 1/R).  Single block, fixed order.
__global__ void __launch_bounds__(256) ce_reduce_kernel(const float* __restrict__ lse, const float* __restrict__ picked, const float* __restrict__ w, int R, float* __restrict__ out) {
    __shared__ float sm[256];
    float acc = 0.f;
    for (int r = threadIdx.x; r < R; r += 256) acc += (w ? w[r] : 1.0f / static_cast<float>(R)) * (lse[r] - picked[r]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// GLORIA backward, elementwise part.  In: S [B, B], lse_r, lse_c, na, nb, upstream g (device scalar).
// Out: G[i, j] = temp * dS / c  (coefficient on the raw dot products), and
//      rn[i, j] = -dS * S / c * [na_i nb_j > eps]   (gradient w.r.t. the clamped norm product)
// reduced on the fly into rs[i] = sum_j rn * nb_j / na_i  (one warp per row).
__global__ void __launch_bounds__(256)
gloria_bwd_rows_kernel(const float* __restrict__ S, int B, const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                       const float* __restrict__ na, const float* __restrict__ nb, float temp, float eps,
                       const float* __restrict__ gout, float* __restrict__ G, float* __restrict__ rs) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= B) return;
    const float g = gout[0] / static_cast<float>(B);
    const float lr = lse_r[i], nai = na[i];
    float acc = 0.f;
    for (int j = lane; j < B; j += 32) {
        const float s = S[static_cast<size_t>(i) * B + j];
        const float onehot = (i == j) ? 1.f : 0.f;
        const float dS = g * ((expf(s - lr) - onehot) + (expf(s - lse_c[j]) - onehot));
        const float prod = nai * nb[j];
        const float c = fmaxf(prod, eps);
        G[static_cast<size_t>(i) * B + j] = temp * dS / c;
        if (prod > eps) acc += -dS * s / c * nb[j];
    }
    acc = warp_sum(acc);
    if (lane == 0) rs[i] = nai > 0.f ? acc / nai : 0.f;
}
// cs[j] = sum_i rn[i, j] * na_i / nb_j, recomputed from S (thread per column)
__global__ void __launch_bounds__(256)
gloria_bwd_cols_kernel(const float* __restrict__ S, int B, const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                       const float* __restrict__ na, const float* __restrict__ nb, float eps,
                       const float* __restrict__ gout, float* __restrict__ cs) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    const float g = gout[0] / static_cast<float>(B);
    const float lc = lse_c[j], nbj = nb[j];
    float acc = 0.f;
    for (int i = 0; i < B; ++i) {
        const float s = S[static_cast<size_t>(i) * B + j];
        const float onehot = (i == j) ? 1.f : 0.f;
        const float dS = g * ((expf(s - lse_r[i]) - onehot) + (expf(s - lc) - onehot));
        const float prod = na[i] * nbj;
        if (prod > eps) acc += -dS * s / fmaxf(prod, eps) * na[i];
    }
    cs[j] = nbj > 0.f ? acc / nbj : 0.f;
}

// FLAVA backward, elementwise: dL[r, c] = coef[r] * (exp(L - lse[r]) - [c == label0 + r]) in place into dL,
// and dscale partial[r] = sum_c dL * L.  coef[r] = g * w[r] (w == nullptr: 1/R).  One warp per row.
__global__ void __launch_bounds__(256)
softmax_ce_bwd_kernel(const float* __restrict__ L, int R, int C, long long ld, int label0, const float* __restrict__ lse,
                      const float* __restrict__ w, const float* __restrict__ gout, float gmul, float* __restrict__ dL,
                      float* __restrict__ dscale_part) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float coef = gout[0] * gmul * (w ? w[r] : 1.0f / static_cast<float>(R));
    const float l = lse[r];
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = L[r * ld + c];
        const float d = coef * (expf(v - l) - ((c == label0 + r) ? 1.f : 0.f));
        dL[r * ld + c] = d;
        acc = fmaf(d, v, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) dscale_part[r] = acc;
}

// out[0] (+)= sum_r x[r], single block, fixed order
__global__ void __launch_bounds__(256) sum_reduce_kernel(const float* __restrict__ x, int R, float* __restrict__ out, int accumulate) {
    __shared__ float sm[256];
    float acc = 0.f;
    for (int r = threadIdx.x; r < R; r += 256) acc += x[r];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + sm[0] : sm[0];
}

// zero-shot: pred[m] = argmax_c cos(I_m, T_c); first max wins (torch.argmax). One warp per image.
__global__ void __launch_bounds__(256)
zeroshot_kernel(const float* __restrict__ I, const float* __restrict__ T, int M, int C, int D, float eps,
                long long* __restrict__ pred, float* __restrict__ sim) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* x = I + static_cast<size_t>(m) * D;
    float nx = 0.f;
    for (int i = lane; i < D; i += 32) nx = fmaf(x[i], x[i], nx);
    nx = sqrtf(warp_sum(nx));
    float best = -INFINITY;
    int best_c = 0;
    for (int c = 0; c < C; ++c) {
        const float* tc = T + static_cast<size_t>(c) * D;
        float dot = 0.f, nt = 0.f;
        for (int i = lane; i < D; i += 32) { const float tv = tc[i]; dot = fmaf(x[i], tv, dot); nt = fmaf(tv, tv, nt); }
        dot = warp_sum(dot);
        nt = sqrtf(warp_sum(nt));
        const float cs = dot / fmaxf(nx * nt, eps);
        if (sim && lane == 0) sim[static_cast<size_t>(m) * C + c] = cs;
        if (cs > best) { best = cs; best_c = c; }
    }
    if (lane == 0) pred[m] = best_c;
}

#ifndef MM_SGEMM_WAVES
#define MM_SGEMM_WAVES 2
#endif
int run_sgemm(const SgemmArgs& a, cudaStream_t st, const char* what) {
    if (a.M <= 0 || a.N <= 0) return MM_OK;
    const int blocks64 = ((a.N + 63) / 64) * ((a.M + 63) / 64);
    const bool big = blocks64 >= 96;      // small problem: quarter-size tiles put four times as many SMs to work
    const int blocks = big ? blocks64 : ((a.N + 31) / 32) * ((a.M + 31) / 32);
    // these GEMMs have few output tiles and a long K (logits of a 256-row batch): split K until two waves of CTAs exist
    int splits = (MM_SGEMM_WAVES * mm::sm_count() + blocks - 1) / blocks;
    splits = splits < 1 ? 1 : (splits > a.K / 128 ? (a.K / 128 > 0 ? a.K / 128 : 1) : splits);
    if (splits > 8) splits = 8;
    if (splits > 1 && (a.X == a.C)) splits = 1;          // in-place residual: the partial sums would race with the reads of X
    if (splits > 1) {
        cudaError_t e = cudaMemset2DAsync(a.C, static_cast<size_t>(a.ldc) * sizeof(float), 0, static_cast<size_t>(a.N) * sizeof(float),
                                          static_cast<size_t>(a.M), st);
        if (e != cudaSuccess) { mm::set_error("%s: cudaMemset2DAsync failed (%s)", what, cudaGetErrorString(e)); return MM_ERR_CUDA; }
    }
    if (big) {
        sgemm_kernel<4><<<dim3((a.N + 63) / 64, (a.M + 63) / 64, splits), 256, 0, st>>>(a);
    } else {
        sgemm_kernel<2><<<dim3((a.N + 31) / 32, (a.M + 31) / 32, splits), 256, 0, st>>>(a);
    }
    mm::note_launches(1);
    return check_launch(what);
}

}  // namespace mm

using namespace mm;

static inline unsigned warp_rows_grid(int R) { return static_cast<unsigned>((R + 7) / 8); }

// ---- GLORIA ---------------------------------------------------------------------------
// workspace (fp32 elements): S [B*B] | G [B*B] | na [B] | nb [B] | lse_r [B] | lse_c [B] | diag [B] | rs [B] | cs [B] | tmp [2]
extern "C" long long mm_gloria_workspace_floats(int B) { return 2LL * B * B + 7LL * B + 8; }

extern "C" int mm_gloria_global_fwd(const float* img, const float* txt, int B, int D, float temp, float eps, float* ws,
                                    float* loss, void* stream) {
    MM_REQUIRE(B > 0 && D > 0, MM_ERR_BAD_SHAPE, "mm_gloria_global_fwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* S = ws; float* na = ws + 2LL * B * B; float* nb = na + B; float* lse_r = nb + B; float* lse_c = lse_r + B;
    float* diag = lse_c + B; float* tmp = diag + 3LL * B;
    row_norm_kernel<<<warp_rows_grid(B), 256, 0, st>>>(img, B, D, na);
    mm::note_launches(1);
    row_norm_kernel<<<warp_rows_grid(B), 256, 0, st>>>(txt, B, D, nb);
    mm::note_launches(1);
    SgemmArgs g{};
    g.A = img; g.sam = D; g.sak = 1; g.B = txt; g.sbk = 1; g.sbn = D; g.C = S; g.ldc = B; g.M = B; g.N = B; g.K = D;
    g.alpha = temp; g.na = na; g.nb = nb; g.eps = eps;
    int rc = run_sgemm(g, st, "mm_gloria_global_fwd(sgemm)");
    if (rc) return rc;
    lse_rows_kernel<<<warp_rows_grid(B), 256, 0, st>>>(S, B, B, B, 0, lse_r, diag);
    mm::note_launches(1);
    lse_cols_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, B, B, B, lse_c);
    mm::note_launches(1);
    ce_reduce_kernel<<<1, 256, 0, st>>>(lse_r, diag, nullptr, B, tmp);
    mm::note_launches(1);
    ce_reduce_kernel<<<1, 256, 0, st>>>(lse_c, diag, nullptr, B, tmp + 1);
    mm::note_launches(1);
    sum_reduce_kernel<<<1, 256, 0, st>>>(tmp, 2, loss, 0);
    mm::note_launches(1);
    return mm_check_launch("mm_gloria_global_fwd");
}

extern "C" int mm_gloria_global_bwd(const float* img, const float* txt, int B, int D, float temp, float eps, float* ws,
                                    const float* gout, float* dimg, float* dtxt, void* stream) {
    MM_REQUIRE(B > 0 && D > 0, MM_ERR_BAD_SHAPE, "mm_gloria_global_bwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* S = ws; float* G = ws + 1LL * B * B; float* na = ws + 2LL * B * B; float* nb = na + B; float* lse_r = nb + B;
    float* lse_c = lse_r + B; float* rs = lse_c + 2LL * B; float* cs = rs + B;
    gloria_bwd_rows_kernel<<<warp_rows_grid(B), 256, 0, st>>>(S, B, lse_r, lse_c, na, nb, temp, eps, gout, G, rs);
    mm::note_launches(1);
    gloria_bwd_cols_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, B, lse_r, lse_c, na, nb, eps, gout, cs);
    mm::note_launches(1);
    int rc = MM_OK;
    if (dimg) {   // dI = G T + rs * I
        SgemmArgs g{};
        g.A = G; g.sam = B; g.sak = 1; g.B = txt; g.sbk = D; g.sbn = 1; g.C = dimg; g.ldc = D; g.M = B; g.N = D; g.K = B;
        g.alpha = 1.f; g.rs = rs; g.X = img; g.ldx = D;
        rc = run_sgemm(g, st, "mm_gloria_global_bwd(dimg)");
        if (rc) return rc;
    }
    if (dtxt) {   // dT = G^T I + cs * T
        SgemmArgs g{};
        g.A = G; g.sam = 1; g.sak = B; g.B = img; g.sbk = D; g.sbn = 1; g.C = dtxt; g.ldc = D; g.M = B; g.N = D; g.K = B;
        g.alpha = 1.f; g.rs = cs; g.X = txt; g.ldx = D;
        rc = run_sgemm(g, st, "mm_gloria_global_bwd(dtxt)");
    }
    return rc;
}

// ---- FLAVA / CLIP-style ---------------------------------------------------------------
// One direction: logits[R, N] = exp(logit_scale) * a[R, D] b_all[N, D]^T ; lse, picked, loss = sum_r w_r (lse - picked)
extern "C" int mm_infonce_fwd(const float* a, const float* b_all, int R, int N, int D, const float* logit_scale_exp,
                              int label0, const float* row_w, float* logits, float* lse, float* picked, float* loss,
                              void* stream) {
    MM_REQUIRE(R > 0 && N > 0 && D > 0 && label0 >= 0 && label0 + R <= N, MM_ERR_BAD_SHAPE, "mm_infonce_fwd: bad shape / label offset");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SgemmArgs g{};
    g.A = a; g.sam = D; g.sak = 1; g.B = b_all; g.sbk = 1; g.sbn = D; g.C = logits; g.ldc = N; g.M = R; g.N = N; g.K = D;
    g.alpha = 1.f; g.alpha_dev = logit_scale_exp;
    int rc = run_sgemm(g, st, "mm_infonce_fwd(sgemm)");
    if (rc) return rc;
    lse_rows_kernel<<<warp_rows_grid(R), 256, 0, st>>>(logits, R, N, N, label0, lse, picked);
    mm::note_launches(1);
    ce_reduce_kernel<<<1, 256, 0, st>>>(lse, picked, row_w, R, loss);
    mm::note_launches(1);
    return mm_check_launch("mm_infonce_fwd");
}

// Backward of one direction. dlogits is scratch [R, N]; da [R, D] = temp * dL b_all ; db_all [N, D] = temp * dL^T a ;
// dscale[0] (+)= sum dL * L  (gradient w.r.t. logit_scale, since d temp / d logit_scale = temp).
extern "C" int mm_infonce_bwd(const float* a, const float* b_all, int R, int N, int D, const float* logit_scale_exp,
                              int label0, const float* row_w, const float* logits, const float* lse, const float* gout,
                              float gmul, float* dlogits, float* row_tmp, float* da, float* db_all, float* dscale,
                              int accumulate_dscale, void* stream) {
    MM_REQUIRE(R > 0 && N > 0 && D > 0, MM_ERR_BAD_SHAPE, "mm_infonce_bwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    softmax_ce_bwd_kernel<<<warp_rows_grid(R), 256, 0, st>>>(logits, R, N, N, label0, lse, row_w, gout, gmul, dlogits, row_tmp);
    mm::note_launches(1);
    if (dscale) {
        sum_reduce_kernel<<<1, 256, 0, st>>>(row_tmp, R, dscale, accumulate_dscale);
        mm::note_launches(1);
    }
    int rc = MM_OK;
    if (da) {
        SgemmArgs g{};
        g.A = dlogits; g.sam = N; g.sak = 1; g.B = b_all; g.sbk = D; g.sbn = 1; g.C = da; g.ldc = D; g.M = R; g.N = D; g.K = N;
        g.alpha = 1.f; g.alpha_dev = logit_scale_exp;
        rc = run_sgemm(g, st, "mm_infonce_bwd(da)");
        if (rc) return rc;
    }
    if (db_all) {
        SgemmArgs g{};
        g.A = dlogits; g.sam = 1; g.sak = N; g.B = a; g.sbk = D; g.sbn = 1; g.C = db_all; g.ldc = D; g.M = N; g.N = D; g.K = R;
        g.alpha = 1.f; g.alpha_dev = logit_scale_exp;
        rc = run_sgemm(g, st, "mm_infonce_bwd(db_all)");
    }
    return rc;
}

extern "C" int mm_l2_normalize_fwd(const float* x, int R, int D, float eps, float* y, float* norms, void* stream) {
    MM_REQUIRE(R > 0 && D > 0, MM_ERR_BAD_SHAPE, "mm_l2_normalize_fwd: bad shape");
    l2_normalize_kernel<<<warp_rows_grid(R), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, R, D, eps, y, norms);
    mm::note_launches(1);
    return mm_check_launch("mm_l2_normalize_fwd");
}
extern "C" int mm_l2_normalize_bwd(const float* dy, const float* y, const float* norms, int R, int D, float eps, float* dx, void* stream) {
    MM_REQUIRE(R > 0 && D > 0, MM_ERR_BAD_SHAPE, "mm_l2_normalize_bwd: bad shape");
    l2_normalize_bwd_kernel<<<warp_rows_grid(R), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, norms, R, D, eps, dx);
    mm::note_launches(1);
    return mm_check_launch("mm_l2_normalize_bwd");
}

extern "C" int mm_zeroshot_argmax(const float* img, const float* txt, int M, int C, int D, float eps, long long* pred,
                                  float* sim, void* stream) {
    MM_REQUIRE(M > 0 && C > 0 && D > 0, MM_ERR_BAD_SHAPE, "mm_zeroshot_argmax: bad shape");
    zeroshot_kernel<<<warp_rows_grid(M), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, txt, M, C, D, eps, pred, sim);
    mm::note_launches(1);
    return mm_check_launch("mm_zeroshot_argmax");
}
