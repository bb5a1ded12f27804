// medmoe_b200 — C-ABI core: error state, device queries, TMA descriptor encoding and the
// grouped-GEMM entry points.  See include/medmoe_b200.h for the contract.
#include "api_internal.h"
#include "gemm.cuh"
#include "gemm_pair.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#ifndef MM_EPI_WARPS_PLAIN
#define MM_EPI_WARPS_PLAIN 16
#endif
#ifndef MM_EPI_WARPS_RANK1
#define MM_EPI_WARPS_RANK1 12      // rank-1 aux epilogue: only the gate tile is prefetched, which leaves room for twelve warps
#endif

namespace mm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct TraceMark { std::string name; cudaEvent_t ev; };
static std::atomic<bool> g_trace{false};
static std::mutex g_trace_mu;
static std::vector<TraceMark> g_marks;

void trace_mark(const char* name, cudaStream_t st) {
    if (!g_trace.load(std::memory_order_relaxed)) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_marks.push_back({name, ev});
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, static_cast<int>(e), cudaGetErrorString(e));
        return MM_ERR_CUDA;
    }
    return MM_OK;
}

int sm_count() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime API so that the library still loads
// (and exports its symbols) on a box without a driver; compute calls then fail loudly.
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int encode_tmap(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                       uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz, const char* what) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver)", what);
        return MM_ERR_NO_DEVICE;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride * 2) % 16 != 0) {
        set_error("%s: TMA operand must be 16-byte aligned with a 16-byte multiple row pitch", what);
        return MM_ERR_MISALIGNED;
    }
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride * 2};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu pitch=%llu box=%ux%u)", what,
                  static_cast<int>(r), (unsigned long long)inner, (unsigned long long)rows,
                  (unsigned long long)row_stride, box_inner, box_rows);
        return MM_ERR_CUDA;
    }
    return MM_OK;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                     uint32_t box_inner, uint32_t box_rows, const char* what) {
    return encode_tmap(out, base, inner, rows, row_stride, box_inner, box_rows, CU_TENSOR_MAP_SWIZZLE_128B, what);
}

int encode_tmap_bf16_swz(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                         uint32_t box_inner, uint32_t box_rows, int swizzle_bytes, const char* what) {
    const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
    return encode_tmap(out, base, inner, rows, row_stride, box_inner, box_rows, swz, what);
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
struct RowsMaps { CUtensorMap a, b, out, aux, gate; };

template <int BN, bool OUT_F32, int AUX>
static int launch_rows(const RowsMaps& m, const RowsGemmArgs& args, cudaStream_t st) {
    // plain bf16 epilogues are bound by the latency of their TMEM -> registers -> staging -> TMA-store chain: sixteen
    // epilogue warps (four per scheduler, one staging slot each) hide it; aux/gate epilogues keep eight (their prefetch
    // slots take the shared memory) as does the fp32-output path.
    constexpr int EW = (AUX == 0 && !OUT_F32 && BN >= 128) ? MM_EPI_WARPS_PLAIN : (AUX == 2 ? MM_EPI_WARPS_RANK1 : 8);
    // smem: pipeline stages + output staging (32 KB, or 16 KB + 64 KB aux/gate staging) must fit 227 KB
#ifdef MM_ROWS_STAGES_DELTA     // tuning experiments: MEDMOE_NVCC_EXTRA="-DMM_ROWS_STAGES_DELTA=-1" (measured: -3..5 % with one stage less)
    constexpr int STAGES = (AUX ? ((BN > 128) ? 3 : 4) : ((BN > 192) ? 4 : (BN > 128 ? 4 : 5))) + (AUX ? 0 : MM_ROWS_STAGES_DELTA);
#else
    constexpr int STAGES = AUX ? ((BN > 128) ? 3 : 4) : ((BN > 192) ? 4 : (BN > 128 ? 4 : 5));
#endif
    using S = GemmSmem<BN, STAGES, AUX, EW>;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_rows_kernel<BN, STAGES, OUT_F32, AUX, EW>;
    static bool configured_dev[64];   // per device (the attribute is per device); benign race: the set is idempotent
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("gemm_rows: cannot opt in to %d B of shared memory (%s)", S::TOTAL, cudaGetErrorString(e));
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    const int work = args.tile_count * args.n_tiles;
    const int grid = work < sm_count() ? work : sm_count();
    if (grid <= 0) return MM_OK;
    kern<<<grid, rows_threads(EW), S::TOTAL, st>>>(m.a, m.b, m.out, m.aux, m.gate, args);
    note_launches(1);
    return check_launch("gemm_rows");
}

template <int BN, bool COLSUM>
static int launch_wgrad(const CUtensorMap& tA, const CUtensorMap& tB, const WgradArgs& args, cudaStream_t st) {
    constexpr int STAGES = COLSUM ? ((BN > 128) ? 4 : (BN > 64 ? 5 : 6)) : ((BN > 192) ? 4 : (BN > 128 ? 5 : 6));
    using S = GemmSmem<BN, STAGES>;
    constexpr int SMEM = S::WGRAD_TOTAL + (COLSUM ? STAGES * 8192 : 0);
    static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_wgrad_kernel<BN, STAGES, COLSUM>;
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            set_error("gemm_wgrad: cannot opt in to %d B of shared memory (%s)", SMEM, cudaGetErrorString(e));
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    const int work = args.chunk_count * args.n_i * args.n_j;
    const int grid = work < sm_count() ? work : sm_count();
    if (grid <= 0) return MM_OK;
    kern<<<grid, 256, SMEM, st>>>(tA, tB, args);
    note_launches(1);
    return check_launch("gemm_wgrad");
}

// CTA-pair variant (gemm_pair.cuh): plain bf16 epilogue only
template <int BN>
static int launch_rows_pair(const RowsMaps& m, const RowsGemmArgs& args, cudaStream_t st) {
    constexpr int EW = 16;
    constexpr int STAGES = 6;
    using S = PairSmem<BN, STAGES, EW>;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_rows_pair_kernel<BN, STAGES, EW>;
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("gemm_rows_pair: cannot opt in to %d B of shared memory (%s)", S::TOTAL, cudaGetErrorString(e));
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    const int work = ((args.tile_count + 1) / 2) * args.n_tiles;
    int grid = 2 * work < sm_count() ? 2 * work : (sm_count() & ~1);
    if (grid <= 0) return MM_OK;
    kern<<<grid, rows_threads(EW), S::TOTAL, st>>>(m.a, m.b, m.out, args);
    note_launches(1);
    return check_launch("gemm_rows_pair");
}

// rank-1 aux + gate epilogue on CTA pairs (the dY GEMM)
template <int BN>
static int launch_rows_pair_r1(const RowsMaps& m, const RowsGemmArgs& args, cudaStream_t st) {
#ifndef MM_PAIR_R1_EW            // tuning experiments: MEDMOE_NVCC_EXTRA="-DMM_PAIR_R1_EW=8 -DMM_PAIR_R1_STAGES=5"
#define MM_PAIR_R1_EW MM_EPI_WARPS_RANK1
#endif
#ifndef MM_PAIR_R1_STAGES
#define MM_PAIR_R1_STAGES 4
#endif
    constexpr int EW = MM_PAIR_R1_EW;
    constexpr int STAGES = MM_PAIR_R1_STAGES;
    using S = PairR1Smem<BN, STAGES, EW>;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_rows_pair_r1_kernel<BN, STAGES, EW>;
    static bool configured_dev[64];
    bool& configured = *per_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("gemm_rows_pair_r1: cannot opt in to %d B of shared memory (%s)", S::TOTAL, cudaGetErrorString(e));
            return MM_ERR_CUDA;
        }
        configured = true;
    }
    const int work = ((args.tile_count + 1) / 2) * args.n_tiles;
    int grid = 2 * work < sm_count() ? 2 * work : (sm_count() & ~1);
    if (grid <= 0) return MM_OK;
    kern<<<grid, rows_threads(EW), S::TOTAL, st>>>(m.a, m.b, m.out, m.gate, args);
    note_launches(1);
    return check_launch("gemm_rows_pair_r1");
}

// CTA pairs are used wherever the caller vouches for the row layout (EPI_PAIR_OK); MEDMOE_GEMM_PAIR=0 / mm_debug_gemm_pair(0)
// switch them off for A/B measurements (bit 0: plain epilogue GEMMs, bit 1: the rank-1 dY GEMM; default 3 = both)
static int g_pair_mode = -1;
extern "C" void mm_debug_gemm_pair(int on) { g_pair_mode = on; }
static int pair_mode() {
    if (g_pair_mode < 0) {
        const char* v = getenv("MEDMOE_GEMM_PAIR");
        g_pair_mode = v ? atoi(v) : 3;
    }
    return g_pair_mode;
}

static int pick_bn_rows(int N) {
    const int cands[] = {256, 192, 128, 96, 64, 32};
    for (int c : cands)
        if (N % c == 0) return c;
    return 0;
}

}  // namespace mm

using namespace mm;

extern "C" const char* mm_last_error(void) { return g_err; }

extern "C" int mm_abi_version(void) { return 1; }

extern "C" long long mm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int mm_device_sm_count(void) { return sm_count(); }

extern "C" void mm_trace_enable(int on) {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    for (auto& m : g_marks) cudaEventDestroy(m.ev);
    g_marks.clear();
    g_trace.store(on != 0);
}

// Synchronises the device and writes "name calls total_ms\n" lines for every traced kernel into buf.
extern "C" int mm_trace_collect(char* buf, int len) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_trace_mu);
    std::map<std::string, std::pair<int, double>> acc;
    for (size_t i = 1; i < g_marks.size(); ++i) {
        if (g_marks[i].name == "begin") continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev) != cudaSuccess) continue;
        auto& a = acc[g_marks[i].name];
        a.first += 1;
        a.second += ms;
    }
    int off = 0;
    for (auto& kv : acc) {
        int n = snprintf(buf + off, off < len ? len - off : 0, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (n < 0 || off + n >= len) break;
        off += n;
    }
    for (auto& m : g_marks) cudaEventDestroy(m.ev);
    g_marks.clear();
    return off;
}

// C[rows, N] = epi(A[rows, K] * W[e][N, K]^T) over 128-row tiles; see include/medmoe_b200.h.
struct Rank1Aux { const float* row_coef; const int32_t* row_vec; const float* vecs; long long ld_vecs; };

static int gemm_rows_impl(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                          long long ldw, const int32_t* tile_info, int tile_begin, int tile_count, int M,
                          const float* bias, const void* aux, long long ld_aux, const Rank1Aux* r1, const void* gate,
                          long long ld_gate, void* out, long long ld_out, int out_f32, float* colsum,
                          float out_scale, int flags, void* stream, const int32_t* cap_len = nullptr, float cap_temp = 0.f,
                          const int32_t* out_g64 = nullptr, long long out_rows = 0) {
    MM_REQUIRE(A && W && out, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_rows: null operand");
    MM_REQUIRE(!(flags & EPI_CAP_SOFTMAX) || (cap_len && !aux && !r1 && !gate && !out_f32 && !colsum), MM_ERR_UNSUPPORTED,
               "mm_grouped_gemm_rows: the caption-softmax epilogue is a plain bf16 epilogue and needs cap_len");
    MM_REQUIRE(K > 0 && K % 8 == 0 && N > 0 && E > 0, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_rows: K must be a positive multiple of 8");
    MM_REQUIRE(((aux != nullptr) || (r1 != nullptr)) == (gate != nullptr), MM_ERR_UNSUPPORTED,
               "mm_grouped_gemm_rows: aux and gate must be given together");
    MM_REQUIRE(!(gate && out_f32), MM_ERR_UNSUPPORTED, "mm_grouped_gemm_rows: aux/gate need a bf16 output");
    if (r1) {
        MM_REQUIRE(r1->row_coef && r1->row_vec && r1->vecs && r1->ld_vecs >= N && r1->ld_vecs % 4 == 0 &&
                       (reinterpret_cast<uintptr_t>(r1->vecs) & 15) == 0,
                   MM_ERR_BAD_SHAPE, "mm_grouped_gemm_rows_rank1: row_coef / row_vec / vecs (16-byte aligned rows) required");
    }
    const int BN = pick_bn_rows(N);
    MM_REQUIRE(BN != 0, MM_ERR_UNSUPPORTED, "mm_grouped_gemm_rows: N must be a multiple of 32");
    if (!tile_info) {
        MM_REQUIRE(M >= 0, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_rows: M < 0");
        tile_count = (M + TILE_M - 1) / TILE_M;
        tile_begin = 0;
    }
    if (tile_count <= 0) return MM_OK;
    const uint64_t io_rows = tile_info ? static_cast<uint64_t>(tile_count) * TILE_M : static_cast<uint64_t>(M);
    MM_REQUIRE((ld_out * (out_f32 ? 4 : 2)) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, MM_ERR_MISALIGNED,
               "mm_grouped_gemm_rows: out must be 16-byte aligned");
    RowsMaps m;
    int rc = encode_tmap_bf16(&m.a, A, static_cast<uint64_t>(K), static_cast<uint64_t>(a_rows), static_cast<uint64_t>(lda), 64,
                              TILE_M, "mm_grouped_gemm_rows(A)");
    if (rc) return rc;
    rc = encode_tmap_bf16(&m.b, W, static_cast<uint64_t>(K), static_cast<uint64_t>(E) * N, static_cast<uint64_t>(ldw), 64, BN,
                          "mm_grouped_gemm_rows(W)");
    if (rc) return rc;
    m.out = m.a; m.aux = m.a; m.gate = m.a;   // placeholders when unused
    MM_REQUIRE(!out_g64 || (!out_f32 && !aux && !r1 && !gate && !colsum && tile_info && out_rows > 0), MM_ERR_UNSUPPORTED,
               "mm_grouped_gemm_rows_scatter: plain bf16 epilogue with tile_info only");
    if (!out_f32) {
        rc = encode_tmap(&m.out, out, static_cast<uint64_t>(N), out_g64 ? static_cast<uint64_t>(out_rows) : io_rows,
                         static_cast<uint64_t>(ld_out), 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "mm_grouped_gemm_rows(out)");
        if (rc) return rc;
    }
    if (aux) {
        rc = encode_tmap(&m.aux, aux, static_cast<uint64_t>(N), io_rows, static_cast<uint64_t>(ld_aux), 32, 32,
                         CU_TENSOR_MAP_SWIZZLE_64B, "mm_grouped_gemm_rows(aux)");
        if (rc) return rc;
    }
    if (gate) {
        rc = encode_tmap(&m.gate, gate, static_cast<uint64_t>(N), io_rows, static_cast<uint64_t>(ld_gate), 32, 32,
                         CU_TENSOR_MAP_SWIZZLE_64B, "mm_grouped_gemm_rows(gate)");
        if (rc) return rc;
    }
    RowsGemmArgs g;
    g.tile_info = reinterpret_cast<const int2*>(tile_info);
    g.tile_begin = tile_begin;
    g.tile_count = tile_count;
    g.M = M;
    g.N = N;
    g.K = K;
    g.n_tiles = N / BN;
    g.bias = bias;
    g.out = out;
    g.ld_out = ld_out;
    g.colsum = colsum;
    g.out_scale = out_scale;
    g.flags = flags;
    g.row_coef = r1 ? r1->row_coef : nullptr;
    g.row_vec = r1 ? r1->row_vec : nullptr;
    g.vecs = r1 ? r1->vecs : nullptr;
    g.ld_vecs = r1 ? r1->ld_vecs : 0;
    g.cap_len = cap_len;
    g.cap_temp = cap_temp;
    g.out_g64 = out_g64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool pair_ok = !out_g64 && (flags & EPI_PAIR_OK) && (tile_info ? (tile_begin % 2 == 0 && tile_count >= 2) : M > TILE_M);
    if (pair_ok && (pair_mode() & 2) && tile_info && r1 && !aux && gate && !colsum && out_scale == 1.0f && !bias && BN == 256) {
        rc = encode_tmap_bf16(&m.b, W, static_cast<uint64_t>(K), static_cast<uint64_t>(E) * N, static_cast<uint64_t>(ldw), 64,
                              BN / 2, "mm_grouped_gemm_rows_rank1(W, pair)");
        if (rc) return rc;
        return launch_rows_pair_r1<256>(m, g, st);
    }
    if (pair_ok && (pair_mode() & 1) && !aux && !r1 && !gate && !out_f32 && !colsum && !cap_len && out_scale == 1.0f &&
        (BN == 192 || BN == 256)) {
        // each CTA of a pair stages half of the W tile
        rc = encode_tmap_bf16(&m.b, W, static_cast<uint64_t>(K), static_cast<uint64_t>(E) * N, static_cast<uint64_t>(ldw), 64,
                              BN / 2, "mm_grouped_gemm_rows(W, pair)");
        if (rc) return rc;
        return BN == 192 ? launch_rows_pair<192>(m, g, st) : launch_rows_pair<256>(m, g, st);
    }
#define MM_ROWS_CASE(bn)                                                                          \
    case bn:                                                                                      \
        if (out_f32) return launch_rows<bn, true, 0>(m, g, st);                                   \
        if (r1) return aux ? launch_rows<bn, false, 3>(m, g, st) : launch_rows<bn, false, 2>(m, g, st);      \
        return aux ? launch_rows<bn, false, 1>(m, g, st) : launch_rows<bn, false, 0>(m, g, st);
    switch (BN) {
        MM_ROWS_CASE(256)
        MM_ROWS_CASE(192)
        MM_ROWS_CASE(128)
        MM_ROWS_CASE(96)
        MM_ROWS_CASE(64)
        MM_ROWS_CASE(32)
    }
#undef MM_ROWS_CASE
    set_error("mm_grouped_gemm_rows: unreachable tile width %d", BN);
    return MM_ERR_UNSUPPORTED;
}

extern "C" int mm_grouped_gemm_rows(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                                    long long ldw, const int32_t* tile_info, int tile_begin, int tile_count, int M,
                                    const float* bias, const void* aux, long long ld_aux, const void* gate,
                                    long long ld_gate, void* out, long long ld_out, int out_f32, float* colsum,
                                    float out_scale, int flags, void* stream) {
    return gemm_rows_impl(A, a_rows, K, lda, W, E, N, ldw, tile_info, tile_begin, tile_count, M, bias, aux, ld_aux, nullptr,
                          gate, ld_gate, out, ld_out, out_f32, colsum, out_scale, flags, stream);
}

// same plain bf16 GEMM whose output is in IMAGE order: 64-row group g of the launch's row space is stored at rows
// out_g64[g] .. + 64 of out [out_rows, N] (mm_dispatch_group_map), padding groups are dropped — the un-permute of the result
// (reference: the gradient of swin.py:105-108's gather) happens in the store.
extern "C" int mm_grouped_gemm_rows_scatter(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                                            long long ldw, const int32_t* tile_info, int tile_begin, int tile_count,
                                            const float* bias, void* out, long long out_rows, long long ld_out,
                                            const int32_t* out_g64, int flags, void* stream) {
    MM_REQUIRE(out_g64 && tile_info, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_rows_scatter: out_g64 and tile_info required");
    return gemm_rows_impl(A, a_rows, K, lda, W, E, N, ldw, tile_info, tile_begin, tile_count, 0, bias, nullptr, 0, nullptr, nullptr, 0,
                          out, ld_out, 0, nullptr, 1.0f, flags & ~EPI_PAIR_OK, stream, nullptr, 0.f, out_g64, out_rows);
}

// out = (A W_e^T + row_coef[row] * vecs[row_vec[row], :]) * [gate > 0]: the dY GEMM when only global_feat has a
// cotangent, so that the combine's gradient w.r.t. Y is rank-1 per image and never materialised.
extern "C" int mm_grouped_gemm_rows_rank1(const void* A, long long a_rows, int K, long long lda, const void* W, int E, int N,
                                          long long ldw, const int32_t* tile_info, int tile_begin, int tile_count,
                                          const float* row_coef, const int32_t* row_vec, const float* vecs,
                                          long long ld_vecs, const void* aux, long long ld_aux, const void* gate,
                                          long long ld_gate, void* out, long long ld_out, float* colsum, int flags,
                                          void* stream) {
    MM_REQUIRE(tile_info, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_rows_rank1: tile_info required");
    const Rank1Aux r1{row_coef, row_vec, vecs, ld_vecs};
    return gemm_rows_impl(A, a_rows, K, lda, W, E, N, ldw, tile_info, tile_begin, tile_count, 0, nullptr, aux, ld_aux, &r1,
                          gate, ld_gate, out, ld_out, 0, colsum, 1.0f, flags & EPI_PAIR_OK, stream);
}

// E = exp(temp1 * softmax over each caption's words of A W^T): the score GEMM of the word-patch attention loss with the first
// softmax in its epilogue, for captions padded to exactly 32 word slots (one 32-column epilogue chunk per caption).
extern "C" int mm_local_scores_softmax_exp(const void* A, long long rows, int K, long long lda, const void* W, int n_caps,
                                           long long ldw, const int32_t* cap_len, float temp1, void* E, long long ld_e,
                                           void* stream) {
    MM_REQUIRE(cap_len && n_caps > 0 && (n_caps * 32) % 128 == 0, MM_ERR_BAD_SHAPE,
               "mm_local_scores_softmax_exp: cap_len required, 32 * n_caps must be a multiple of 128");
    return gemm_rows_impl(A, rows, K, lda, W, 1, n_caps * 32, ldw, nullptr, 0, 0, static_cast<int>(rows), nullptr, nullptr, 0,
                          nullptr, nullptr, 0, E, ld_e, 0, nullptr, 1.0f, EPI_CAP_SOFTMAX, stream, cap_len, temp1);
}

// dW[e][N1, N2] += sum_rows A[row, N1]^T B[row, N2] over the chunks of expert e (+ optional column sums of A).
static int gemm_wgrad_impl(const void* A, long long a_rows, int N1, long long lda, const void* B, long long b_rows, int N2,
                           long long ldb, const int32_t* chunks, int chunk_begin, int chunk_count, int tile_base,
                           float* out, float* colsum, void* stream, const int32_t* b_g64 = nullptr) {
    MM_REQUIRE(A && B && out && chunks, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_wgrad: null operand");
    MM_REQUIRE(N1 > 0 && N1 % 8 == 0 && N2 > 0 && N2 % 8 == 0, MM_ERR_BAD_SHAPE,
               "mm_grouped_gemm_wgrad: N1, N2 must be positive multiples of 8");
    if (chunk_count <= 0) return MM_OK;
    int BN;
    if (N2 % 256 == 0 && !colsum) BN = 256;
    else if (N2 % 192 == 0) BN = 192;
    else if (N2 <= 64) BN = 64;
    else if (N2 <= 128) BN = 128;
    else if (N2 <= 192 || colsum) BN = 192;
    else BN = 256;
    CUtensorMap tA, tB;
    int rc = encode_tmap_bf16(&tA, A, static_cast<uint64_t>(N1), static_cast<uint64_t>(a_rows), static_cast<uint64_t>(lda), 64, 64,
                              "mm_grouped_gemm_wgrad(A)");
    if (rc) return rc;
    rc = encode_tmap_bf16(&tB, B, static_cast<uint64_t>(N2), static_cast<uint64_t>(b_rows), static_cast<uint64_t>(ldb), 64, 64,
                          "mm_grouped_gemm_wgrad(B)");
    if (rc) return rc;
    WgradArgs g;
    g.chunks = reinterpret_cast<const int4*>(chunks);
    g.chunk_begin = chunk_begin;
    g.chunk_count = chunk_count;
    g.tile_base = tile_base;
    g.N1 = N1;
    g.N2 = N2;
    g.n_i = (N1 + TILE_M - 1) / TILE_M;
    g.n_j = (N2 + BN - 1) / BN;
    g.out = out;
    g.colsum = colsum;
    g.b_g64 = b_g64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (colsum) {
        switch (BN) {
            case 192: return launch_wgrad<192, true>(tA, tB, g, st);
            case 128: return launch_wgrad<128, true>(tA, tB, g, st);
            case 64: return launch_wgrad<64, true>(tA, tB, g, st);
        }
    } else {
        switch (BN) {
            case 256: return launch_wgrad<256, false>(tA, tB, g, st);
            case 192: return launch_wgrad<192, false>(tA, tB, g, st);
            case 128: return launch_wgrad<128, false>(tA, tB, g, st);
            case 64: return launch_wgrad<64, false>(tA, tB, g, st);
        }
    }
    set_error("mm_grouped_gemm_wgrad: unreachable tile width %d", BN);
    return MM_ERR_UNSUPPORTED;
}

extern "C" int mm_grouped_gemm_wgrad(const void* A, long long a_rows, int N1, long long lda, const void* B,
                                     long long b_rows, int N2, long long ldb, const int32_t* chunks, int chunk_begin,
                                     int chunk_count, int tile_base, float* out, void* stream) {
    return gemm_wgrad_impl(A, a_rows, N1, lda, B, b_rows, N2, ldb, chunks, chunk_begin, chunk_count, tile_base, out, nullptr,
                           stream);
}

// same, and colsum[e][i] += sum over the rows of expert e of A[row, i]: the bias gradient that goes with a weight
// gradient, produced by the same MMAs through a constant ones block appended to B.
extern "C" int mm_grouped_gemm_wgrad_colsum(const void* A, long long a_rows, int N1, long long lda, const void* B,
                                            long long b_rows, int N2, long long ldb, const int32_t* chunks, int chunk_begin,
                                            int chunk_count, int tile_base, float* out, float* colsum, void* stream) {
    MM_REQUIRE(colsum, MM_ERR_BAD_SHAPE, "mm_grouped_gemm_wgrad_colsum: colsum is NULL");
    return gemm_wgrad_impl(A, a_rows, N1, lda, B, b_rows, N2, ldb, chunks, chunk_begin, chunk_count, tile_base, out, colsum,
                           stream);
}

// mm_grouped_gemm_wgrad_colsum with B in IMAGE order: 64-row group g of the launch's row space = rows b_g64[g] .. + 64 of B
// (mm_dispatch_group_map; the rows of A that belong to padding groups are zero).
extern "C" int mm_grouped_gemm_wgrad_colsum_gather(const void* A, long long a_rows, int N1, long long lda, const void* B,
                                                   long long b_rows, int N2, long long ldb, const int32_t* chunks,
                                                   int chunk_begin, int chunk_count, int tile_base, float* out, float* colsum,
                                                   const int32_t* b_g64, void* stream) {
    MM_REQUIRE(colsum && b_g64 && tile_base == 0, MM_ERR_BAD_SHAPE,
               "mm_grouped_gemm_wgrad_colsum_gather: colsum / b_g64 required, tile_base must be 0 (groups are launch-relative)");
    return gemm_wgrad_impl(A, a_rows, N1, lda, B, b_rows, N2, ldb, chunks, chunk_begin, chunk_count, tile_base, out, colsum,
                           stream, b_g64);
}

