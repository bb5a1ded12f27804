// medmoe_b200 — embedding exchange of the global contrastive loss over NVLink peer memory (one node, <= 8 GPUs).
//
// Replaces, for the all-gather InfoNCE, the two NCCL collectives of reference src/utils/distributed.py:28-58 as used by
// src/losses.py:503-524:  all_gather(cat([a, b], 1)) forward, reduce_scatter(SUM) of its gradient backward.  The messages are
// tiny ([B_loc, 2 * 768] fp32 = 1.5 MB per rank), so the NCCL calls are pure launch + protocol latency (~3 x 80 us exposed
// per step at 8 GPUs).  Here every rank owns ONE workspace allocated with cudaMalloc and shared through CUDA IPC:
//
//   control   flags_ag[8] @ 0, flags_rs[8] @ 64 (written by the peers), step_ag @ 128, step_rs @ 132, CTA counter @ 136 (local)
//   gath      [2 parities][world][bytes_per_rank]   every rank's block, written BY the owners of the blocks (peer stores)
//   dcol      [2 parities][world][bytes_per_rank]   this rank's gradient w.r.t. all gathered rows, read BY the peers (peer loads)
//
//   gather:  p2p_put_kernel     my block -> slot `rank` of every peer's gath (16-byte stores over NVLink); the last CTA fences,
//                               publishes step s in every peer's flags_ag[rank] (st.release.sys) and waits until all of its own
//                               flags_ag reached s (ld.acquire.sys) — the kernel ends when the gathered tensor is complete
//            p2p_copy_kernel    gath[s & 1] -> the caller's tensor
//   scatter: p2p_stage_kernel   gradient -> dcol[s & 1]; last CTA: publish flags_rs, wait for all
//            p2p_pull_kernel    out[i] = sum_q peer_q.dcol[s & 1][my block][i]   (fixed order q = 0..W-1: deterministic)
//
// The step counters live on the device and are advanced by the kernels themselves, so a captured CUDA graph replays
// correctly.  Parity double-buffering makes the exchange safe without any further hand-shake: a rank can be at most one
// exchange ahead of the slowest one (it cannot pass its own wait before every peer has published the same step).
// Every spin loop gives up after ~10 s and traps: a protocol bug or a dead peer becomes an error, never a hung GPU.
#include <cstdio>
#include <cstring>
#include "mm_common.cuh"
#include "api_internal.h"

namespace mm {

constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_CTRL_BYTES = 4096;
constexpr int P2P_OFF_FLAGS_AG = 0, P2P_OFF_FLAGS_RS = 64, P2P_OFF_STEP_AG = 128, P2P_OFF_STEP_RS = 132, P2P_OFF_COUNTER = 136;

struct P2PPeers { char* buf[P2P_MAX_WORLD]; };

MM_DEVINL void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
MM_DEVINL unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// executed by ONE thread after every CTA of the launch has finished its stores (and fenced them to system scope)
MM_DEVINL void p2p_publish_and_wait(const P2PPeers& peers, int rank, int world, int off_flags, int off_step, unsigned s) {
    char* me = peers.buf[rank];
    *reinterpret_cast<volatile unsigned*>(me + off_step) = s;
    __threadfence_system();
    for (int p = 0; p < world; ++p) st_release_sys(reinterpret_cast<unsigned*>(peers.buf[p] + off_flags) + rank, s);
    for (int q = 0; q < world; ++q) {
        const unsigned* f = reinterpret_cast<const unsigned*>(me + off_flags) + q;
        long long spins = 0;
        while (static_cast<int>(ld_acquire_sys(f) - s) < 0) {
            __nanosleep(100);
            if (++spins > (1LL << 26)) {
                printf("medmoe_b200 p2p exchange: rank %d gave up waiting for rank %d (step %u)\n", rank, q, s);
                __trap();
            }
        }
    }
}

// last-CTA-done: returns true in exactly one thread of the grid, after all CTAs passed this point
MM_DEVINL bool p2p_last_cta(unsigned* counter) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x != 0) return false;
    const unsigned done = atomicAdd(counter, 1u);
    if (done != gridDim.x - 1) return false;
    *counter = 0;
    __threadfence_system();
    return true;
}

// src [n16] (16-byte units) -> block `rank` of gath[s & 1] in every peer
__global__ void __launch_bounds__(256)
p2p_put_kernel(const uint4* __restrict__ src, long long n16, const P2PPeers peers, int rank, int world, long long bytes_per_rank) {
    char* me = peers.buf[rank];
    const unsigned s = *reinterpret_cast<volatile unsigned*>(me + P2P_OFF_STEP_AG) + 1u;
    const long long off = P2P_CTRL_BYTES + (static_cast<long long>(s & 1u) * world + rank) * bytes_per_rank;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = src[i];
        for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(peers.buf[p] + off)[i] = v;
    }
    if (p2p_last_cta(reinterpret_cast<unsigned*>(me + P2P_OFF_COUNTER)))
        p2p_publish_and_wait(peers, rank, world, P2P_OFF_FLAGS_AG, P2P_OFF_STEP_AG, s);
}

// gath[step_ag & 1] (all blocks) -> dst
__global__ void __launch_bounds__(256)
p2p_copy_kernel(uint4* __restrict__ dst, long long n16_total, const char* me, int world, long long bytes_per_rank) {
    const unsigned s = *reinterpret_cast<const volatile unsigned*>(me + P2P_OFF_STEP_AG);
    const uint4* src = reinterpret_cast<const uint4*>(me + P2P_CTRL_BYTES + static_cast<long long>(s & 1u) * world * bytes_per_rank);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16_total; i += stride) dst[i] = src[i];
}

// grad [world * n16] -> dcol[s & 1]; then publish / wait
__global__ void __launch_bounds__(256)
p2p_stage_kernel(const uint4* __restrict__ grad, long long n16_total, const P2PPeers peers, int rank, int world, long long bytes_per_rank) {
    char* me = peers.buf[rank];
    const unsigned s = *reinterpret_cast<volatile unsigned*>(me + P2P_OFF_STEP_RS) + 1u;
    uint4* dst = reinterpret_cast<uint4*>(me + P2P_CTRL_BYTES + (2LL + (s & 1u)) * world * bytes_per_rank);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16_total; i += stride) dst[i] = grad[i];
    if (p2p_last_cta(reinterpret_cast<unsigned*>(me + P2P_OFF_COUNTER)))
        p2p_publish_and_wait(peers, rank, world, P2P_OFF_FLAGS_RS, P2P_OFF_STEP_RS, s);
}

// out[i] = sum over ranks q (ascending) of q's dcol[step_rs & 1][block rank][i]
__global__ void __launch_bounds__(256)
p2p_pull_kernel(float4* __restrict__ out, long long n16, const P2PPeers peers, int rank, int world, long long bytes_per_rank) {
    const unsigned s = *reinterpret_cast<const volatile unsigned*>(peers.buf[rank] + P2P_OFF_STEP_RS);
    const long long off = P2P_CTRL_BYTES + ((2LL + (s & 1u)) * world + rank) * bytes_per_rank;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v[P2P_MAX_WORLD];
#pragma unroll
        for (int q = 0; q < P2P_MAX_WORLD; ++q)
            if (q < world) v[q] = reinterpret_cast<const float4*>(peers.buf[q] + off)[i];      // all loads in flight first
#pragma unroll
        for (int q = 0; q < P2P_MAX_WORLD; ++q)
            if (q < world) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        out[i] = acc;
    }
}

static int fill_peers(P2PPeers& p, void* const* peer_bufs, int rank, int world, const char* what) {
    if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || !peer_bufs) {
        set_error("%s: world must be 1..%d and rank inside it", what, P2P_MAX_WORLD);
        return 2;
    }
    for (int q = 0; q < P2P_MAX_WORLD; ++q) p.buf[q] = q < world ? static_cast<char*>(peer_bufs[q]) : nullptr;
    for (int q = 0; q < world; ++q)
        if (!p.buf[q]) { set_error("%s: null peer buffer", what); return 2; }
    return 0;
}

static int p2p_grid(long long n16) {
    const long long want = (n16 + 255) / 256;
    const int cap = 2 * sm_count();
    return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace mm

using namespace mm;

extern "C" long long mm_p2p_workspace_bytes(int world, long long bytes_per_rank) {
    return P2P_CTRL_BYTES + 4LL * world * bytes_per_rank;
}

// cudaMalloc + zero + IPC handle (64 bytes written to `handle_out`)
extern "C" int mm_p2p_alloc(long long bytes, void** ptr_out, void* handle_out) {
    MM_REQUIRE(bytes > 0 && ptr_out && handle_out, 2, "mm_p2p_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
    if (e != cudaSuccess) { set_error("mm_p2p_alloc: cudaMalloc(%lld) failed (%s)", bytes, cudaGetErrorString(e)); return 1; }
    e = cudaMemset(p, 0, static_cast<size_t>(bytes));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        set_error("mm_p2p_alloc: %s", cudaGetErrorString(e));
        cudaFree(p);
        cudaGetLastError();
        return 1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    *ptr_out = p;
    return 0;
}

extern "C" int mm_p2p_open(const void* handle, void** ptr_out) {
    MM_REQUIRE(handle && ptr_out, 2, "mm_p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("mm_p2p_open: cudaIpcOpenMemHandle failed (%s)", cudaGetErrorString(e));
        cudaGetLastError();
        return 1;
    }
    *ptr_out = p;
    return 0;
}

extern "C" int mm_p2p_close(void* ptr) {
    if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) { cudaGetLastError(); return 1; }
    return 0;
}

extern "C" int mm_p2p_free(void* ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); return 1; }
    return 0;
}

// all-gather: src [bytes_per_rank] of this rank -> dst [world * bytes_per_rank] (block q = rank q's src)
extern "C" int mm_p2p_all_gather(const void* src, void* dst, long long bytes_per_rank, void* const* peer_bufs, int rank,
                                 int world, void* stream) {
    P2PPeers p;
    if (int rc = fill_peers(p, peer_bufs, rank, world, "mm_p2p_all_gather")) return rc;
    MM_REQUIRE(src && dst && bytes_per_rank > 0 && bytes_per_rank % 16 == 0, 2, "mm_p2p_all_gather: bytes_per_rank must be a positive multiple of 16");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n16 = bytes_per_rank / 16;
    p2p_put_kernel<<<p2p_grid(n16), 256, 0, st>>>(static_cast<const uint4*>(src), n16, p, rank, world, bytes_per_rank);
    p2p_copy_kernel<<<p2p_grid(n16 * world), 256, 0, st>>>(static_cast<uint4*>(dst), n16 * world, p.buf[rank], world, bytes_per_rank);
    note_launches(2);
    return check_launch("mm_p2p_all_gather");
}

// reduce-scatter(SUM), fp32: grad [world * bytes_per_rank] of this rank -> out [bytes_per_rank] = sum_q grad_q[block rank]
extern "C" int mm_p2p_reduce_scatter_f32(const void* grad, void* out, long long bytes_per_rank, void* const* peer_bufs, int rank,
                                         int world, void* stream) {
    P2PPeers p;
    if (int rc = fill_peers(p, peer_bufs, rank, world, "mm_p2p_reduce_scatter_f32")) return rc;
    MM_REQUIRE(grad && out && bytes_per_rank > 0 && bytes_per_rank % 16 == 0, 2, "mm_p2p_reduce_scatter_f32: bytes_per_rank must be a positive multiple of 16");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n16 = bytes_per_rank / 16;
    p2p_stage_kernel<<<p2p_grid(n16 * world), 256, 0, st>>>(static_cast<const uint4*>(grad), n16 * world, p, rank, world, bytes_per_rank);
    p2p_pull_kernel<<<p2p_grid(n16), 256, 0, st>>>(static_cast<float4*>(out), n16, p, rank, world, bytes_per_rank);
    note_launches(2);
    return check_launch("mm_p2p_reduce_scatter_f32");
}
