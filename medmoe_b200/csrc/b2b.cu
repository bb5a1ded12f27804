// medmoe_b200 — C-ABI of the back-to-back expert GEMM kernels (b2b.cuh).  Contract: include/medmoe_b200.h.
#include <cstdlib>
#include "api_internal.h"
#include "b2b.cuh"

using namespace mm;

namespace {

// MEDMOE_B2B_DEBUG (tuning / A-B runs): bit 0 = never use CTA pairs
int b2b_debug_flags() {
    static const int v = [] { const char* e = getenv("MEDMOE_B2B_DEBUG"); return e ? atoi(e) : 0; }();
    return v;
}

template <int NKB1>
int launch_b2b_fwd(const CUtensorMap& tA1, const CUtensorMap& tB1, const CUtensorMap& tB2, const CUtensorMap& tY,
                   const CUtensorMap& tZ, const B2BFwdArgs& args, cudaStream_t st) {
    using S = B2BSmem<NKB1>;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget exceeded");
    auto kern = b2b_fwd_kernel<NKB1>;
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("expert_b2b_fwd: cannot opt in to %d B of shared memory (%s)", S::TOTAL, cudaGetErrorString(e));
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    const int grid = args.tile_count < sm_count() ? args.tile_count : sm_count();
    if (grid <= 0) return MM_OK;
    kern<<<grid, B2B_THREADS, S::TOTAL, st>>>(tA1, tB1, tB2, tY, tZ, args);
    note_launches(1);
    return check_launch("expert_b2b_fwd");
}

template <int NKB1>
int launch_b2b_pair_fwd(const CUtensorMap& tA1, const CUtensorMap& tB1, const CUtensorMap& tB2, const CUtensorMap& tY,
                        const CUtensorMap& tZ, const B2BFwdArgs& args, cudaStream_t st) {
    using S = B2BPairSmem<NKB1>;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget exceeded");
    auto kern = b2b_pair_fwd_kernel<NKB1>;
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("expert_b2b_fwd(pair): cannot opt in to %d B of shared memory (%s)", S::TOTAL, cudaGetErrorString(e));
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    const int n_pairs = (args.tile_count + 1) / 2;
    const int clusters = n_pairs < sm_count() / 2 ? n_pairs : sm_count() / 2;
    if (clusters <= 0) return MM_OK;
    kern<<<2 * clusters, B2B_THREADS, S::TOTAL, st>>>(tA1, tB1, tB2, tY, tZ, args);      // __cluster_dims__(2, 1, 1)
    note_launches(1);
    return check_launch("expert_b2b_fwd(pair)");
}

}  // namespace

// 1 when mm_expert_b2b_fwd covers this shape (otherwise run the two GEMMs separately through mm_grouped_gemm_rows)
// (K1 in (128, 192] only on CTA pairs — flags & 1 of mm_expert_b2b_fwd —: the halved weight stages leave room for a 48 KB f tile)
extern "C" int mm_expert_b2b_fwd_supported(int K1, int D, int H) {
    if (!(D == B2B_D && H == B2B_H && K1 > 0 && K1 % 16 == 0)) return 0;
    if (K1 <= 128) return 1;
    return (K1 <= 192 && !(b2b_debug_flags() & 1)) ? 2 : 0;
}

// Y = ReLU(f Wp_e^T + bp_e) (bf16, written once) and Z = Y W1_e^T + b1_e (bf16) over the 128-row tiles
// [tile_begin, tile_begin + tile_count) of one scale region; f / Y / Z point at the region's first row.
static int b2b_fwd_impl(const void* f, long long f_rows, int K1, long long ldf, const void* Wp, int E, int D,
                        long long ldwp, const float* bias1, const void* W1, int H, long long ldw1,
                        const float* bias2, const int32_t* tile_info, int tile_begin, int tile_count, void* Y,
                        long long ld_y, void* Z, long long ld_z, int flags, const int32_t* f_g64, void* stream) {
    MM_REQUIRE(f && Wp && W1 && bias1 && bias2 && tile_info && Y && Z, MM_ERR_BAD_SHAPE, "mm_expert_b2b_fwd: null operand");
    const int sup = mm_expert_b2b_fwd_supported(K1, D, H);
    MM_REQUIRE(sup == 1 || (sup == 2 && (flags & 1) && tile_count >= 2), MM_ERR_UNSUPPORTED,
               "mm_expert_b2b_fwd: needs D = 768, H = 384 and K1 a multiple of 16 up to 128 (up to 192 on CTA pairs, flags & 1)");
    if (tile_count <= 0) return MM_OK;
    const uint64_t io_rows = static_cast<uint64_t>(tile_count) * TILE_M;
    // flags & 1: the caller guarantees that the tiles (2j, 2j + 1) of this launch never belong to two experts (SEG_ALIGN row
    // layout) -> CTA pairs (cta_group::2) share every weight tile: each CTA stages half of its rows
    const bool pairs = (flags & 1) && !(b2b_debug_flags() & 1) && tile_count >= 2;
    MM_REQUIRE(!f_g64 || pairs, MM_ERR_UNSUPPORTED, "mm_expert_b2b_fwd_gather: the group-map addressing of f needs the CTA-pair kernel (flags & 1)");
    CUtensorMap tA1, tB1, tB2, tY, tZ;
    int rc = encode_tmap_bf16(&tA1, f, static_cast<uint64_t>(K1), static_cast<uint64_t>(f_rows), static_cast<uint64_t>(ldf), 64,
                              f_g64 ? 64 : TILE_M, "mm_expert_b2b_fwd(f)");
    if (rc) return rc;
    rc = encode_tmap_bf16(&tB1, Wp, static_cast<uint64_t>(K1), static_cast<uint64_t>(E) * D, static_cast<uint64_t>(ldwp), 64,
                          pairs ? B2B_NC / 2 : B2B_NC, "mm_expert_b2b_fwd(Wp)");
    if (rc) return rc;
    rc = encode_tmap_bf16(&tB2, W1, static_cast<uint64_t>(D), static_cast<uint64_t>(E) * H, static_cast<uint64_t>(ldw1), 64,
                          pairs ? B2B_H / 4 : B2B_H / 2, "mm_expert_b2b_fwd(W1)");
    if (rc) return rc;
    rc = encode_tmap_bf16(&tY, Y, static_cast<uint64_t>(D), io_rows, static_cast<uint64_t>(ld_y), 64, TILE_M,
                          "mm_expert_b2b_fwd(Y)");
    if (rc) return rc;
    rc = encode_tmap_bf16(&tZ, Z, static_cast<uint64_t>(H), io_rows, static_cast<uint64_t>(ld_z), 64, TILE_M,
                          "mm_expert_b2b_fwd(Z)");
    if (rc) return rc;
    B2BFwdArgs g;
    g.tile_info = reinterpret_cast<const int2*>(tile_info);
    g.tile_begin = tile_begin;
    g.tile_count = tile_count;
    g.K1 = K1;
    g.bias1 = bias1;
    g.bias2 = bias2;
    g.y = static_cast<__nv_bfloat16*>(Y);
    g.ld_y = ld_y;
    g.z = static_cast<__nv_bfloat16*>(Z);
    g.ld_z = ld_z;
    g.f_g64 = f_g64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pairs) return K1 <= 64    ? launch_b2b_pair_fwd<1>(tA1, tB1, tB2, tY, tZ, g, st)
                      : K1 <= 128 ? launch_b2b_pair_fwd<2>(tA1, tB1, tB2, tY, tZ, g, st)
                                  : launch_b2b_pair_fwd<3>(tA1, tB1, tB2, tY, tZ, g, st);
    return K1 <= 64 ? launch_b2b_fwd<1>(tA1, tB1, tB2, tY, tZ, g, st) : launch_b2b_fwd<2>(tA1, tB1, tB2, tY, tZ, g, st);
}

extern "C" int mm_expert_b2b_fwd(const void* f, long long f_rows, int K1, long long ldf, const void* Wp, int E, int D,
                                 long long ldwp, const float* bias1, const void* W1, int H, long long ldw1,
                                 const float* bias2, const int32_t* tile_info, int tile_begin, int tile_count, void* Y,
                                 long long ld_y, void* Z, long long ld_z, int flags, void* stream) {
    return b2b_fwd_impl(f, f_rows, K1, ldf, Wp, E, D, ldwp, bias1, W1, H, ldw1, bias2, tile_info, tile_begin, tile_count, Y, ld_y, Z,
                        ld_z, flags, nullptr, stream);
}

// same with f in IMAGE order ([f_rows, K1] = the reference's [B, P, K1] stage feature itself): the 64-row groups of the launch's
// expert-sorted row space are addressed through f_g64 (mm_dispatch_group_map), so no sorted copy of f is made.  Needs flags & 1.
extern "C" int mm_expert_b2b_fwd_gather(const void* f, long long f_rows, int K1, long long ldf, const void* Wp, int E, int D,
                                        long long ldwp, const float* bias1, const void* W1, int H, long long ldw1,
                                        const float* bias2, const int32_t* tile_info, int tile_begin, int tile_count, void* Y,
                                        long long ld_y, void* Z, long long ld_z, int flags, const int32_t* f_g64, void* stream) {
    MM_REQUIRE(f_g64, MM_ERR_BAD_SHAPE, "mm_expert_b2b_fwd_gather: f_g64 is NULL");
    return b2b_fwd_impl(f, f_rows, K1, ldf, Wp, E, D, ldwp, bias1, W1, H, ldw1, bias2, tile_info, tile_begin, tile_count, Y, ld_y, Z,
                        ld_z, flags, f_g64, stream);
}
