// medmoe_b200 — internals shared by the C-ABI translation units (error state, launch checks,
// TMA descriptor encoding).  The public contract is include/medmoe_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace mm {
// thread-local last-error text (backward runs on the autograd engine's worker thread)
void set_error(const char* fmt, ...);
int check_launch(const char* what);
// 2-D bf16 tensor map, 128-byte swizzle.  inner = contiguous extent (elements), rows = outer extent,
// row_stride = elements between rows, box = {box_inner (<= 64), box_rows (<= 256)}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                     uint32_t box_inner, uint32_t box_rows, const char* what);
// same with a selectable swizzle (0 / 32 / 64 / 128 bytes)
int encode_tmap_bf16_swz(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                         uint32_t box_inner, uint32_t box_rows, int swizzle_bytes, const char* what);
int sm_count();
// slot of the current device in a per-kernel `static bool[64]`: cudaFuncSetAttribute opt-ins are per device
static inline bool* per_device_flag(bool* flags) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return &flags[dev];
}
// number of kernels launched through the C-ABI since load (bench.py's gpu_launches evidence)
void note_launches(int n);
// optional per-kernel timing inside composite entry points: one CUDA event per mark; the time between two
// consecutive marks on a stream is attributed to the later mark's name ("begin" restarts the chain)
void trace_mark(const char* name, cudaStream_t st);
// generic strided fp32 GEMM used by the loss and router kernels (loss.cu):
//   C[m, n] = alpha * den(m, n) * sum_k A(m, k) B(k, n) + rs[m] * X[m, n]
//   A(m, k) = A[m * sam + k * sak],  B(k, n) = B[k * sbk + n * sbn];  den = 1 / max(na[m] nb[n], eps) when na != nullptr
struct SgemmArgs {
    const float* A; long long sam, sak;
    const float* B; long long sbk, sbn;
    float* C; long long ldc;
    int M, N, K;
    float alpha;
    const float* alpha_dev;  // optional device scalar multiplied into alpha (exp(logit_scale))
    const float* na; const float* nb; float eps;
    const float* rs; const float* X; long long ldx;
};
int run_sgemm(const SgemmArgs& a, cudaStream_t st, const char* what);
}  // namespace mm

#define MM_REQUIRE(cond, code, msg)        \
    do {                                   \
        if (!(cond)) {                     \
            mm::set_error("%s", msg);      \
            return code;                   \
        }                                  \
    } while (0)

static inline int mm_check_launch(const char* what) { return mm::check_launch(what); }
