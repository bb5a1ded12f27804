// medmoe_b200 — token dispatch (north-star kernel 2; SURVEY §2.2 "dispatch = the absence of M1").
//
// The reference runs every expert on every image and gathers afterwards
// (swin.py:105-108).  Here images ("items" = (image, top-k choice) pairs) are counting-
// sorted by expert so every expert owns contiguous, 128-row-aligned row segments in each
// of the S scale regions; 128-row GEMM tiles therefore never straddle experts.
//
//   mm_dispatch_build : routing ints -> counts/offsets, stable perm / inv_perm, per-slot
//                       row starts, per-tile {expert, valid_rows} table, wgrad chunk table.
//                       Stable order uses warp-level match/popc prefix ranks.
//   mm_dispatch_rows  : gather each image's contiguous [P_s, D_s] block into its sorted
//                       slot (cast fp32 -> bf16 on the way, zero the segment padding);
//                       16-byte vectorised, fully coalesced on both sides.
//   mm_undispatch_rows: the transpose — sorted bf16 gradients back to image order in the
//                       caller's dtype, summing the k slots of an image.
#include "mm_common.cuh"
#include "api_internal.h"

namespace mm {

constexpr int MAX_EXPERTS = 64;
constexpr int MAX_SCALES = 8;

struct DispatchPlanArgs {
    int n_items, K, S;
    int P[MAX_SCALES];            // rows per item at scale s
    int region_base[MAX_SCALES];  // first global row of region s (multiple of 128)
    int region_tiles[MAX_SCALES]; // capacity of region s in 128-row tiles
    int chunk_base[MAX_SCALES];   // first wgrad chunk entry of region s
    int chunk_cap[MAX_SCALES];    // chunk entries reserved for region s
    int chunk_tiles[MAX_SCALES];  // tiles per wgrad chunk in region s
};

__global__ void __launch_bounds__(1024)
dispatch_build_kernel(const int* __restrict__ item_expert, DispatchPlanArgs a, int* __restrict__ counts,
                      int* __restrict__ offsets, int* __restrict__ perm, int* __restrict__ inv_perm,
                      int* __restrict__ slot_expert, int* __restrict__ seg_start, int* __restrict__ slot_row,
                      int2* __restrict__ tile_info, int4* __restrict__ chunks) {
    __shared__ int s_cnt[MAX_EXPERTS];
    __shared__ int s_off[MAX_EXPERTS + 1];
    __shared__ int s_run[MAX_EXPERTS];
    __shared__ int s_seg[MAX_SCALES][MAX_EXPERTS];
    __shared__ int s_wcnt[32][MAX_EXPERTS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K;

    if (tid < MAX_EXPERTS) { s_cnt[tid] = 0; s_run[tid] = 0; }
    __syncthreads();
    // expert ids outside [0, K) (a caller's bug, or garbage from a faulted producer) are routed to expert 0 instead of
    // indexing shared memory out of bounds: every item then still owns a slot, so no consumer reads unwritten tables
    auto expert_of = [&](int i) { const int e = item_expert[i]; return (static_cast<unsigned>(e) < static_cast<unsigned>(K)) ? e : 0; };
    for (int i = tid; i < a.n_items; i += blockDim.x) atomicAdd(&s_cnt[expert_of(i)], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int e = 0; e < K; ++e) { s_off[e] = acc; acc += s_cnt[e]; }
        s_off[K] = acc;
    }
    if (tid < a.S) {
        const int s = tid;
        int r = a.region_base[s];
        for (int e = 0; e < K; ++e) {
            s_seg[s][e] = r;
            r += (s_cnt[e] * a.P[s] + SEG_ALIGN - 1) / SEG_ALIGN * SEG_ALIGN;
        }
    }
    __syncthreads();
    if (tid < K) counts[tid] = s_cnt[tid];
    if (tid <= K) offsets[tid] = s_off[tid];
    for (int i = tid; i < a.S * K; i += blockDim.x) seg_start[i] = s_seg[i / K][i % K];

    // ---- stable rank of every item within its expert ----
    for (int base = 0; base < a.n_items; base += blockDim.x) {
        for (int i = tid; i < 32 * MAX_EXPERTS; i += blockDim.x) (&s_wcnt[0][0])[i] = 0;
        __syncthreads();
        const int item = base + tid;
        const int e = item < a.n_items ? expert_of(item) : -1 - lane;   // distinct dummy keys
        const unsigned peers = __match_any_sync(0xffffffffu, e);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (e >= 0 && rank_in_warp == 0) s_wcnt[warp][e] = __popc(peers);
        __syncthreads();
        if (e >= 0) {
            int before = s_run[e];
            for (int w = 0; w < warp; ++w) before += s_wcnt[w][e];
            const int rank = before + rank_in_warp;
            const int slot = s_off[e] + rank;
            perm[slot] = item;
            inv_perm[item] = slot;
            slot_expert[slot] = e;
            for (int s = 0; s < a.S; ++s) slot_row[s * a.n_items + slot] = s_seg[s][e] + rank * a.P[s];
        }
        __syncthreads();
        if (tid < K) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += s_wcnt[w][tid];
            s_run[tid] += tot;
        }
        __syncthreads();
    }

    // ---- per-tile {expert, valid rows} ----
    int tile_base = 0;
    for (int s = 0; s < a.S; ++s) {
        for (int t = tid; t < a.region_tiles[s]; t += blockDim.x) {
            const int row = a.region_base[s] + t * TILE_M;
            int2 info = make_int2(-1, 0);
            for (int e = 0; e < K; ++e) {
                const int rows_e = s_cnt[e] * a.P[s];
                if (rows_e > 0 && row >= s_seg[s][e] && row < s_seg[s][e] + rows_e)
                    info = make_int2(e, min(TILE_M, s_seg[s][e] + rows_e - row));
            }
            tile_info[tile_base + t] = info;
        }
        tile_base += a.region_tiles[s];
    }

    // ---- wgrad chunk table: consecutive tiles of one expert, at most chunk_tiles[s] each ----
    if (tid < a.S) {
        const int s = tid;
        int n = 0;
        const int G = a.chunk_tiles[s];
        for (int e = 0; e < K; ++e) {
            const int nt = (s_cnt[e] * a.P[s] + TILE_M - 1) / TILE_M;
            const int first = s_seg[s][e] / TILE_M;
            for (int c = 0; c < nt; c += G) {
                if (n < a.chunk_cap[s]) chunks[a.chunk_base[s] + n] = make_int4(e, first + c, min(G, nt - c), s);
                ++n;
            }
        }
        for (; n < a.chunk_cap[s]; ++n) chunks[a.chunk_base[s] + n] = make_int4(0, 0, 0, s);
    }
}

// 8 elements per thread-iteration: one 16-byte bf16 store, one or two 16-byte loads.
template <typename SrcT>
MM_DEVINL uint4 load8_as_bf16(const SrcT* p);
template <>
MM_DEVINL uint4 load8_as_bf16<__nv_bfloat16>(const __nv_bfloat16* p) { return ldg_nc_v4(p); }
template <>
MM_DEVINL uint4 load8_as_bf16<float>(const float* p) {
    const uint4 lo = ldg_nc_v4(p), hi = ldg_nc_v4(p + 4);
    uint4 r;
    r.x = pack_bf16x2(__uint_as_float(lo.x), __uint_as_float(lo.y));
    r.y = pack_bf16x2(__uint_as_float(lo.z), __uint_as_float(lo.w));
    r.z = pack_bf16x2(__uint_as_float(hi.x), __uint_as_float(hi.y));
    r.w = pack_bf16x2(__uint_as_float(hi.z), __uint_as_float(hi.w));
    return r;
}

struct DispatchRowsArgs {
    int n_items, topk, K, S;
    int P[MAX_SCALES], D[MAX_SCALES];
    int region_base[MAX_SCALES];
    const void* src[MAX_SCALES];        // [B, P_s, D_s] in image order
    __nv_bfloat16* dst[MAX_SCALES];     // [region rows, D_s] sorted (row 0 == region_base[s])
};

constexpr int DISPATCH_ELEMS_PER_BLOCK = 256 * 8 * 4;

// grid = (blocks per item, n_items + K, S).  y < n_items: copy slot y; y >= n_items: zero
// the padding rows behind expert (y - n_items)'s segment.
template <typename SrcT>
__global__ void __launch_bounds__(256)
dispatch_rows_kernel(DispatchRowsArgs a, const int* __restrict__ perm, const int* __restrict__ slot_row,
                     const int* __restrict__ counts, const int* __restrict__ seg_start) {
    const int s = blockIdx.z;
    const long long per_item = static_cast<long long>(a.P[s]) * a.D[s];
    if (blockIdx.y < a.n_items) {
        const int slot = blockIdx.y;
        const long long e0 = static_cast<long long>(blockIdx.x) * DISPATCH_ELEMS_PER_BLOCK;
        if (e0 >= per_item) return;
        const int img = perm[slot] / a.topk;
        const SrcT* src = static_cast<const SrcT*>(a.src[s]) + img * per_item;
        __nv_bfloat16* dst = a.dst[s] + static_cast<long long>(slot_row[s * a.n_items + slot] - a.region_base[s]) * a.D[s];
        const long long e1 = min(per_item, e0 + DISPATCH_ELEMS_PER_BLOCK);
        for (long long i = e0 + threadIdx.x * 8; i < e1; i += 256 * 8) stg_v4(dst + i, load8_as_bf16<SrcT>(src + i));
    } else {
        const int e = blockIdx.y - a.n_items;
        const long long rows = static_cast<long long>(counts[e]) * a.P[s];
        const long long pad_rows = (rows + TILE_M - 1) / TILE_M * TILE_M - rows;
        const long long n = pad_rows * a.D[s];
        __nv_bfloat16* dst = a.dst[s] + (static_cast<long long>(seg_start[s * a.K + e] - a.region_base[s]) + rows) * a.D[s];
        for (long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8; i < n;
             i += static_cast<long long>(gridDim.x) * 256 * 8)
            stg_v4(dst + i, make_uint4(0, 0, 0, 0));
    }
}

struct UndispatchRowsArgs {
    int n_images, topk, S;
    int P[MAX_SCALES], D[MAX_SCALES];
    int region_base[MAX_SCALES];
    const __nv_bfloat16* src[MAX_SCALES];   // sorted gradients
    void* dst[MAX_SCALES];                  // [B, P_s, D_s] image order
};

template <typename DstT>
__global__ void __launch_bounds__(256)
undispatch_rows_kernel(UndispatchRowsArgs a, const int* __restrict__ inv_perm, const int* __restrict__ slot_row) {
    const int s = blockIdx.z, img = blockIdx.y;
    const long long per_item = static_cast<long long>(a.P[s]) * a.D[s];
    const long long e0 = static_cast<long long>(blockIdx.x) * DISPATCH_ELEMS_PER_BLOCK;
    if (e0 >= per_item) return;
    const long long e1 = min(per_item, e0 + DISPATCH_ELEMS_PER_BLOCK);
    const int n_items = a.n_images * a.topk;
    if (sizeof(DstT) == 2 && a.topk == 1) {
        // one choice per image, bf16 in and out: a plain permuted copy (no unpack / accumulate / repack, source row looked up once)
        const int slot = inv_perm[img];
        const __nv_bfloat16* src = a.src[s] + static_cast<long long>(slot_row[s * n_items + slot] - a.region_base[s]) * a.D[s];
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(a.dst[s]) + img * per_item;
        for (long long i = e0 + threadIdx.x * 8; i < e1; i += 256 * 8) stg_v4(d + i, ldg_nc_v4(src + i));
        return;
    }
    for (long long i = e0 + threadIdx.x * 8; i < e1; i += 256 * 8) {
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < a.topk; ++j) {
            const int slot = inv_perm[img * a.topk + j];
            const __nv_bfloat16* src = a.src[s] + static_cast<long long>(slot_row[s * n_items + slot] - a.region_base[s]) * a.D[s];
            const uint4 u = ldg_nc_v4(src + i);
            f[0] += bf16lo(u.x); f[1] += bf16hi(u.x); f[2] += bf16lo(u.y); f[3] += bf16hi(u.y);
            f[4] += bf16lo(u.z); f[5] += bf16hi(u.z); f[6] += bf16lo(u.w); f[7] += bf16hi(u.w);
        }
        if constexpr (sizeof(DstT) == 4) {
            float* d = static_cast<float*>(a.dst[s]) + img * per_item + i;
            *reinterpret_cast<float4*>(d) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(f[4], f[5], f[6], f[7]);
        } else {
            __nv_bfloat16* d = static_cast<__nv_bfloat16*>(a.dst[s]) + img * per_item + i;
            stg_v4(d, make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                 pack_bf16x2(f[6], f[7])));
        }
    }
}

// fp32 -> bf16 cast of a flat array (weight shadows); n multiple of 8 handled vectorised, tail scalar.
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
    for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (i + 8 <= n) {
            stg_v4(dst + i, load8_as_bf16<float>(src + i));
        } else {
            for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}

// dst[c, r] = bf16(src[r, c]) for a batch of [R, C] fp32 matrices (transposed weight shadows for dgrad).
__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C) {
    __shared__ float tile[32][33];
    const size_t mat = static_cast<size_t>(blockIdx.z) * R * C;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < R && c < C) ? src[mat + static_cast<size_t>(r) * C + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (r < R && c < C) dst[mat + static_cast<size_t>(c) * R + r] = __float2bfloat16_rn(tile[tx][j]);
    }
}

// g64[group] = first row, in image order, of the 64 consecutive rows that make up 64-row group `group` of the finest region's
// expert-sorted row space (or -1 for padding / slack).  P0 % 64 == 0 and segments start on 256-row boundaries, so a group never
// straddles two items.  With this map the finest scale needs no sorted copy: TMA boxes of 64 rows are addressed through it.
// grid = (ceil(P0 / 64 / 256) | ceil(n_groups / 256), n_items + 1); block y == n_items clears, the others fill their item's groups.
__global__ void __launch_bounds__(256)
group_map_fill_kernel(const int* __restrict__ perm, const int* __restrict__ slot_row0, int n_items, int topk, int P0, int region_base0,
                      int* __restrict__ g64) {
    const int slot = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= P0 / 64) return;
    const int img = perm[slot] / topk;
    g64[(slot_row0[slot] - region_base0) / 64 + j] = img * P0 + 64 * j;
}
__global__ void __launch_bounds__(256) group_map_clear_kernel(int* __restrict__ g64, int n_groups) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g < n_groups) g64[g] = -1;
}

}  // namespace mm

using namespace mm;

// see include/medmoe_b200.h
extern "C" int mm_dispatch_group_map(const int32_t* perm, const int32_t* slot_row0, int n_items, int topk, int P0,
                                     int region_base0, int n_groups, int32_t* g64, void* stream) {
    MM_REQUIRE(perm && slot_row0 && g64 && n_items >= 0 && topk >= 1 && P0 > 0 && P0 % 64 == 0 && region_base0 % 64 == 0 &&
                   n_groups >= 0,
               MM_ERR_BAD_SHAPE, "mm_dispatch_group_map: P0 and region_base0 must be multiples of 64");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_groups == 0) return MM_OK;
    group_map_clear_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(g64, n_groups);
    if (n_items > 0) group_map_fill_kernel<<<dim3((P0 / 64 + 255) / 256, n_items), 256, 0, st>>>(perm, slot_row0, n_items, topk, P0,
                                                                                                region_base0, g64);
    mm::note_launches(2);
    return mm_check_launch("mm_dispatch_group_map");
}

extern "C" int mm_dispatch_build(const int32_t* item_expert, int n_items, int K, int S, const int32_t* P,
                                 const int32_t* region_base, const int32_t* region_tiles, const int32_t* chunk_base,
                                 const int32_t* chunk_cap, const int32_t* chunk_tiles, int32_t* counts,
                                 int32_t* offsets, int32_t* perm, int32_t* inv_perm, int32_t* slot_expert,
                                 int32_t* seg_start, int32_t* slot_row, int32_t* tile_info, int32_t* chunks,
                                 void* stream) {
    MM_REQUIRE(n_items >= 0 && K > 0 && K <= MAX_EXPERTS && S > 0 && S <= MAX_SCALES, MM_ERR_BAD_SHAPE,
               "mm_dispatch_build: need 0 < K <= 64 and 0 < S <= 8");
    DispatchPlanArgs a;
    a.n_items = n_items; a.K = K; a.S = S;
    for (int s = 0; s < S; ++s) {
        MM_REQUIRE(region_base[s] % TILE_M == 0 && chunk_tiles[s] > 0, MM_ERR_BAD_SHAPE,
                   "mm_dispatch_build: region_base must be a multiple of 128 and chunk_tiles positive");
        a.P[s] = P[s]; a.region_base[s] = region_base[s]; a.region_tiles[s] = region_tiles[s];
        a.chunk_base[s] = chunk_base[s]; a.chunk_cap[s] = chunk_cap[s]; a.chunk_tiles[s] = chunk_tiles[s];
    }
    dispatch_build_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        item_expert, a, counts, offsets, perm, inv_perm, slot_expert, seg_start, slot_row,
        reinterpret_cast<int2*>(tile_info), reinterpret_cast<int4*>(chunks));
    mm::note_launches(1);
    return mm_check_launch("mm_dispatch_build");
}

extern "C" int mm_dispatch_rows(const void* const* src, int src_is_f32, void* const* dst, int n_items, int topk, int K,
                                int S, const int32_t* P, const int32_t* D, const int32_t* region_base,
                                const int32_t* perm, const int32_t* slot_row, const int32_t* counts,
                                const int32_t* seg_start, void* stream) {
    MM_REQUIRE(S > 0 && S <= MAX_SCALES && topk >= 1 && n_items >= 0, MM_ERR_BAD_SHAPE, "mm_dispatch_rows: bad shape");
    DispatchRowsArgs a;
    a.n_items = n_items; a.topk = topk; a.K = K; a.S = S;
    long long max_per_item = 0;
    for (int s = 0; s < S; ++s) {
        MM_REQUIRE(D[s] % 8 == 0, MM_ERR_BAD_SHAPE, "mm_dispatch_rows: feature widths must be multiples of 8");
        MM_REQUIRE((reinterpret_cast<uintptr_t>(src[s]) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst[s]) & 15) == 0,
                   MM_ERR_MISALIGNED, "mm_dispatch_rows: buffers must be 16-byte aligned");
        a.P[s] = P[s]; a.D[s] = D[s]; a.region_base[s] = region_base[s];
        a.src[s] = src[s]; a.dst[s] = static_cast<__nv_bfloat16*>(dst[s]);
        const long long per = static_cast<long long>(P[s]) * D[s];
        if (per > max_per_item) max_per_item = per;
    }
    const unsigned gx = static_cast<unsigned>((max_per_item + DISPATCH_ELEMS_PER_BLOCK - 1) / DISPATCH_ELEMS_PER_BLOCK);
    dim3 grid(gx > 0 ? gx : 1, n_items + K, S);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (src_is_f32)
        dispatch_rows_kernel<float><<<grid, 256, 0, st>>>(a, perm, slot_row, counts, seg_start);
    else
        dispatch_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, perm, slot_row, counts, seg_start);
    mm::note_launches(1);
    return mm_check_launch("mm_dispatch_rows");
}

extern "C" int mm_undispatch_rows(const void* const* src, void* const* dst, int dst_is_f32, int n_images, int topk, int S,
                                  const int32_t* P, const int32_t* D, const int32_t* region_base,
                                  const int32_t* inv_perm, const int32_t* slot_row, void* stream) {
    MM_REQUIRE(S > 0 && S <= MAX_SCALES && topk >= 1 && n_images >= 0, MM_ERR_BAD_SHAPE, "mm_undispatch_rows: bad shape");
    if (n_images == 0) return MM_OK;
    UndispatchRowsArgs a;
    a.n_images = n_images; a.topk = topk; a.S = S;
    long long max_per_item = 0;
    for (int s = 0; s < S; ++s) {
        MM_REQUIRE(D[s] % 8 == 0, MM_ERR_BAD_SHAPE, "mm_undispatch_rows: feature widths must be multiples of 8");
        a.P[s] = P[s]; a.D[s] = D[s]; a.region_base[s] = region_base[s];
        a.src[s] = static_cast<const __nv_bfloat16*>(src[s]); a.dst[s] = dst[s];
        const long long per = static_cast<long long>(P[s]) * D[s];
        if (per > max_per_item) max_per_item = per;
    }
    const unsigned gx = static_cast<unsigned>((max_per_item + DISPATCH_ELEMS_PER_BLOCK - 1) / DISPATCH_ELEMS_PER_BLOCK);
    dim3 grid(gx, n_images, S);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dst_is_f32)
        undispatch_rows_kernel<float><<<grid, 256, 0, st>>>(a, inv_perm, slot_row);
    else
        undispatch_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, inv_perm, slot_row);
    mm::note_launches(1);
    return mm_check_launch("mm_undispatch_rows");
}

extern "C" int mm_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
    if (n <= 0) return MM_OK;
    MM_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, MM_ERR_MISALIGNED,
               "mm_cast_f32_bf16: buffers must be 16-byte aligned");
    long long blocks = (n / 8 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cast_f32_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(dst), n);
    mm::note_launches(1);
    return mm_check_launch("mm_cast_f32_bf16");
}

extern "C" int mm_transpose_cast_f32_bf16(const float* src, void* dst, int batch, int R, int C, void* stream) {
    MM_REQUIRE(batch > 0 && R > 0 && C > 0, MM_ERR_BAD_SHAPE, "mm_transpose_cast_f32_bf16: bad shape");
    dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
    transpose_cast_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), R, C);
    mm::note_launches(1);
    return mm_check_launch("mm_transpose_cast_f32_bf16");
}

// ---------------------------------------------------------------------------------------
// mm_pack_expert_params: every expert's fp32 master parameters (separate nn.Parameter storages, reference names
// experts.{e}.proj_convs.{s}.0.{weight,bias}, experts.{e}.attn_proj.{0,2}.{weight,bias}; swin.py:18-30) -> the stacked
// operands the kernels read, in ONE launch: bf16 weights [E, rows, cols], their bf16 transposes [E, cols, rows] (the
// dgrad GEMMs), fp32 biases / vectors [E, n].  Replaces ~25 torch stack / cast / transpose launches per step.
// The job table travels by value in the kernel parameters (no host -> device copy, CUDA-graph capturable).
// ---------------------------------------------------------------------------------------
namespace mm {
struct PackJob {
    const float* src;      // [rows, cols] fp32
    void* dst;             // bf16 [rows, cols] (kind 0) or fp32 [rows * cols] (kind 1)
    __nv_bfloat16* dstT;   // bf16 [cols, rows] or nullptr (kind 0 only)
    int rows, cols;
    int kind, pad;
};
constexpr int PACK_MAX_JOBS = 12 * 64;
struct PackArgs { int n_jobs; PackJob job[PACK_MAX_JOBS]; };

__global__ void __launch_bounds__(256) pack_params_kernel(const __grid_constant__ PackArgs a) {
    __shared__ float tile[32][33];
    const PackJob j = a.job[blockIdx.y];
    if (j.kind == 1) {
        const int n = j.rows * j.cols;
        for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) static_cast<float*>(j.dst)[i] = j.src[i];
        return;
    }
    const int tr = (j.rows + 31) / 32, tc = (j.cols + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(j.dst);
    for (int t = blockIdx.x; t < tr * tc; t += gridDim.x) {
        const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty + 8 * i, c = c0 + tx;
            float v = 0.f;
            if (r < j.rows && c < j.cols) {
                v = j.src[static_cast<size_t>(r) * j.cols + c];
                dst[static_cast<size_t>(r) * j.cols + c] = __float2bfloat16_rn(v);
            }
            tile[ty + 8 * i][tx] = v;
        }
        if (j.dstT) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c0 + ty + 8 * i, r = r0 + tx;
                if (r < j.rows && c < j.cols) j.dstT[static_cast<size_t>(c) * j.rows + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
            }
            __syncthreads();
        }
    }
}
}  // namespace mm

// src / dst / dstT: HOST arrays of n_jobs DEVICE pointers; rows / cols / kind: host int arrays.
extern "C" int mm_pack_expert_params(const void* const* src, void* const* dst, void* const* dstT, const int32_t* rows,
                                     const int32_t* cols, const int32_t* kind, int n_jobs, void* stream) {
    MM_REQUIRE(n_jobs >= 0 && n_jobs <= PACK_MAX_JOBS, MM_ERR_BAD_SHAPE, "mm_pack_expert_params: at most 768 tensors (64 experts)");
    if (n_jobs == 0) return MM_OK;
    static thread_local PackArgs a;      // 24 KB: kept off the stack
    a.n_jobs = n_jobs;
    for (int i = 0; i < n_jobs; ++i) {
        MM_REQUIRE(src[i] && dst[i] && rows[i] > 0 && cols[i] > 0, MM_ERR_BAD_SHAPE, "mm_pack_expert_params: null tensor");
        a.job[i] = PackJob{static_cast<const float*>(src[i]), dst[i], static_cast<__nv_bfloat16*>(dstT ? dstT[i] : nullptr), rows[i],
                           cols[i], kind[i], 0};
    }
    pack_params_kernel<<<dim3(48, n_jobs), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    mm::note_launches(1);
    return mm_check_launch("mm_pack_expert_params");
}
