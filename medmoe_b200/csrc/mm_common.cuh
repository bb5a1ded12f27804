// medmoe_b200 — shared device/host helpers for the sm_100a kernels.
//
// Thin inline-PTX wrappers for the Blackwell primitives the hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// shared-memory / instruction descriptors consumed by tcgen05.mma.  Nothing in this
// file is a port of reference code (the reference, shivangchopra11/MedMoE, is pure
// Python); it is the plumbing the CUDA restatement of swin.py / losses.py sits on.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef MM_DEVINL
#define MM_DEVINL __device__ __forceinline__
#endif

namespace mm {

// ------------------------------------------------------------------------------------
// Status codes (mirrored in include/medmoe_b200.h)
// ------------------------------------------------------------------------------------
enum : int {
    MM_OK = 0,
    MM_ERR_BAD_SHAPE = -1,
    MM_ERR_MISALIGNED = -2,
    MM_ERR_UNSUPPORTED = -3,
    MM_ERR_CUDA = -4,
    MM_ERR_NO_DEVICE = -5,
    MM_ERR_WORKSPACE = -6,
};

constexpr int TILE_M = 128;   // rows per GEMM tile == TMEM lanes
constexpr int SEG_ALIGN = 256;  // an expert's segment starts on a multiple of 256 rows in every region: the tiles (2j, 2j + 1) of a
                              // region never belong to two experts, so a CTA pair (tcgen05 cta_group::2) can share one weight tile

// ------------------------------------------------------------------------------------
// Small math helpers
// ------------------------------------------------------------------------------------
MM_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
MM_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Column sums of a 32x32 register tile: lane r holds row r (f[0..31]); on return lane c
// holds sum over rows of column c in f[0].  31 shuffles.
MM_DEVINL float warp_colsum32(float (&f)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            float send = up ? f[i] : f[i + o];
            float keep = up ? f[i + o] : f[i];
            f[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return f[0];
}

MM_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
MM_DEVINL float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
MM_DEVINL float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// 1-D linear interpolation source of output index j (align_corners=False), the closed
// form of F.interpolate(mode='linear') used at reference swin.py:42.
// scale = P_src / P_dst as float, exactly like ATen's area_pixel_compute_scale.
struct LerpSrc { int i0, i1; float lam; };
MM_DEVINL LerpSrc lerp_src(int j, float scale, int p_src) {
    float src = fmaxf(scale * (static_cast<float>(j) + 0.5f) - 0.5f, 0.0f);
    int i0 = min(static_cast<int>(src), p_src - 1);
    float lam = fminf(fmaxf(src - static_cast<float>(i0), 0.0f), 1.0f);
    int i1 = i0 + (i0 < p_src - 1 ? 1 : 0);
    return LerpSrc{i0, i1, lam};
}

// ------------------------------------------------------------------------------------
// Shared-memory address / mbarrier
// ------------------------------------------------------------------------------------
MM_DEVINL uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

MM_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MM_DEVINL void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MM_DEVINL void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
MM_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
MM_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MM_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (mbarrier.test_wait): for roles that poll several barriers in turn
MM_DEVINL bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must become a trap (reported as a CUDA error by the
// C-ABI), never a hung GPU box.
MM_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) { __trap(); }
    }
}

// ------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------
MM_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates are (c0 = innermost/contiguous, c1 = row).
MM_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 1-D bulk copy global -> shared (TMA engine, no tensor map): 16-byte aligned, size multiple of 16.
MM_DEVINL void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// named barrier among a subset of the CTA's warps
MM_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 2-D tile store smem -> global (bulk async group); out-of-bounds rows/cols are clipped by the tensor map.
MM_DEVINL void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
MM_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk store groups still READ their shared-memory source
template <int N>
MM_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
MM_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------
MM_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
}
MM_DEVINL void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MM_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
MM_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MM_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs, fp32 accumulate.
MM_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a converged warp (elect.sync): the MMA-issue and TMA-producer roles run as "whole warp enters, one lane
// elected" so that the compiler keeps descriptor arithmetic on the uniform datapath without per-instruction election loops.
MM_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// Shared-memory descriptor advanced by `bytes` (a multiple of 16 that stays inside the 14-bit start-address field):
// one 32-bit add on the low word instead of rebuilding the descriptor for every k step.
MM_DEVINL uint64_t smem_desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread retire.
MM_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: lane = row (32 lanes of this warp's quarter), 32 consecutive columns.
MM_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
MM_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for tcgen05.mma (sm_100 "version 1" format):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
// K-major  SW128: rows of 128 B, 8-row swizzle atoms of 1024 B; SBO = stride between atoms
//                 along M/N (1024 B for a dense tile); LBO unused (1).
// MN-major SW128: an atom is 64 MN-elements (128 B) x 8 k-rows = 1024 B; SBO = stride
//                 between 8-row k groups, LBO = stride between 64-element MN chunks.
MM_DEVINL uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    d |= 2ull << 61;   // SWIZZLE_128B
    return d;
}

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32.
//   [4,6) c_format=1 (F32) | [7,10) a_format=1 (BF16) | [10,13) b_format=1 (BF16)
//   [15] a_major (0 = K, 1 = MN) | [16] b_major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------
// Vector global memory access
// ------------------------------------------------------------------------------------
MM_DEVINL uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
MM_DEVINL void stg_v4(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
MM_DEVINL void red_add_v4_f32(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

}  // namespace mm
