"""Thin tensor-level wrappers over the C-ABI (include/medmoe_b200.h).

Each wrapper only checks device / dtype / contiguity and forwards raw pointers plus the
current CUDA stream.  There is deliberately no CPU or PyTorch fallback: a tensor that is
not on a CUDA device raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib
from .plan import DispatchPlan, RowLayout

EPI_RELU, EPI_ZERO_PAD = 1, 2
EPI_PAIR_OK = 8      # plan-based launches: tiles (2j, 2j + 1) share an expert (plan.SEG_ALIGN), CTA pairs may share weight tiles
_P = _lib.ptr
# test hook: run the generic (any scale ratio) backward-combine kernel instead of the token-centric one
FORCE_GENERIC_COMBINE_BWD = False
# test hook: run the CUDA-core forward combine instead of the tensor-core one (combine_mma.cuh)
FORCE_CUDA_CORE_COMBINE_FWD = False


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("medmoe_b200 kernels run on CUDA tensors only (there is no CPU fallback)")


def _st():
    return _lib.stream_ptr()


# ---- per-scale launches ------------------------------------------------------------------------------
# The four per-scale launches of one logical op are independent.  Forking the three small ones onto side streams
# (events, CUDA-graph capturable) was measured on B200: no gain (5.87 vs 5.81 ms/step) — every kernel here is a
# persistent one-CTA-per-SM kernel, so a side launch only gets SMs when the big launch retires.  Kept as a switch.
_SIDE_STREAMS = {}
USE_SIDE_STREAMS = False


def run_scales(fns):
    """Run fns[0] on the current stream and fns[1:] on side streams; join before returning."""
    if not USE_SIDE_STREAMS or len(fns) <= 1:
        for f in fns:
            f()
        return
    cur = torch.cuda.current_stream()
    dev = cur.device
    pool = _SIDE_STREAMS.setdefault(dev, [])
    while len(pool) < len(fns) - 1:
        pool.append(torch.cuda.Stream(device=dev))
    fork = torch.cuda.Event()
    fork.record(cur)
    joins = []
    for f, side in zip(fns[1:], pool):
        side.wait_event(fork)
        with torch.cuda.stream(side):
            f()
            ev = torch.cuda.Event()
            ev.record(side)
        joins.append(ev)
    fns[0]()
    for ev in joins:
        cur.wait_event(ev)


# ---- router ---------------------------------------------------------------------------
NEAR_TIE_TOL = 1e-6    # report |p1 - p2| < 1e-6


def router_topk(x, W1, b1, W2, b2, topk: int):
    """-> (hidden, probs, idx, w, near_tie): near_tie [B] int32 flags the images whose selection margin is < NEAR_TIE_TOL."""
    _need_cuda(x, W1, b1, W2, b2)
    B, D = x.shape
    K = W2.shape[0]
    f32 = dict(dtype=torch.float32, device=x.device)
    hidden = torch.empty(B, 128, **f32)
    probs = torch.empty(B, K, **f32)
    idx = torch.empty(B, topk, dtype=torch.int32, device=x.device)
    w = torch.empty(B, topk, **f32)
    near_tie = torch.empty(B, dtype=torch.int32, device=x.device)
    _lib.call("mm_router_topk", _P(x), B, D, _P(W1), _P(b1), _P(W2), _P(b2), K, topk, _P(hidden), _P(probs), _P(idx),
              _P(w), _P(near_tie), float(NEAR_TIE_TOL), _st())
    return hidden, probs, idx, w, near_tie


def router_bwd(dprobs, probs, hidden, x, W1, W2, need_dx: bool):
    _need_cuda(dprobs, probs, hidden, x, W1, W2)
    B, D = x.shape
    K = W2.shape[0]
    f32 = dict(dtype=torch.float32, device=x.device)
    dlogit = torch.empty(B, K, **f32)
    dhidden = torch.empty(B, 128, **f32)
    dx = torch.empty(B, D, **f32) if need_dx else None
    dW1 = torch.empty(128, D, **f32); db1 = torch.empty(128, **f32)
    dW2 = torch.empty(K, 128, **f32); db2 = torch.empty(K, **f32)
    _lib.call("mm_router_bwd", _P(dprobs), _P(probs), _P(hidden), _P(x), _P(W1), _P(W2), B, D, K, _P(dlogit), _P(dhidden),
              _P(dx), _P(dW1), _P(db1), _P(dW2), _P(db2), _st())
    return dx, dW1, db1, dW2, db2


# ---- dispatch -------------------------------------------------------------------------
def dispatch_rows(feats: Sequence[torch.Tensor], plan: DispatchPlan, widths: Sequence[int], first_scale: int = 0) -> List:
    """[B, P_s, D_s] (fp32|bf16, image order) -> bf16 [region_rows_s, D_s] (expert-sorted, padded).
    `first_scale` = 1 leaves the finest scale out (its consumers address the image-order tensor through the group map):
    the returned list then has None in its place."""
    lay = plan.layout
    _need_cuda(*feats)
    src_f32 = feats[0].dtype == torch.float32
    for f in feats:
        if f.dtype != feats[0].dtype or f.dtype not in (torch.float32, torch.bfloat16) or not f.is_contiguous():
            raise RuntimeError("stage features must be contiguous and all fp32 or all bf16")
    fs = first_scale
    dst = [torch.empty(lay.region_rows[s], widths[s], dtype=torch.bfloat16, device=feats[0].device) for s in range(fs, lay.S)]
    _lib.call("mm_dispatch_rows", _lib.host_ptrs(feats[fs:]), int(src_f32), _lib.host_ptrs(dst), lay.n_items, lay.topk,
              lay.num_experts, lay.S - fs, _lib.host_i32(lay.P[fs:]), _lib.host_i32(widths[fs:]), _lib.host_i32(lay.region_base[fs:]),
              _P(plan.perm), _P(plan.slot_row[fs:]), _P(plan.counts), _P(plan.seg_start[fs:]), _st())
    return [None] * fs + dst


def undispatch_rows(srcs: Sequence[torch.Tensor], plan: DispatchPlan, widths: Sequence[int], out_dtype, first_scale: int = 0) -> List:
    """`first_scale` = 1: srcs[0] is ignored (the finest scale's gradient was stored in image order by its GEMM); None in its place."""
    lay = plan.layout
    fs = first_scale
    _need_cuda(*srcs[fs:])
    dst = [torch.empty(lay.n_images, lay.P[s], widths[s], dtype=out_dtype, device=srcs[fs].device) for s in range(fs, lay.S)]
    _lib.call("mm_undispatch_rows", _lib.host_ptrs(srcs[fs:]), _lib.host_ptrs(dst), int(out_dtype == torch.float32),
              lay.n_images, lay.topk, lay.S - fs, _lib.host_i32(lay.P[fs:]), _lib.host_i32(widths[fs:]),
              _lib.host_i32(lay.region_base[fs:]), _P(plan.inv_perm), _P(plan.slot_row[fs:]), _st())
    return [None] * fs + dst


# The finest scale without a sorted copy (csrc: mm_dispatch_group_map and the *_gather / *_scatter entry points): top-1, bf16
# features, P_0 a multiple of 64, CTA-pair back-to-back kernel available.  Switch kept for A/B measurements and tests.
USE_DIRECT_FINEST = True


def group_map(plan: DispatchPlan) -> torch.Tensor:
    """int32 [region_rows_0 / 64]: first image-order row of every 64-row group of the finest region's sorted row space (-1: padding)."""
    lay = plan.layout
    n_groups = lay.region_rows[0] // 64
    g64 = torch.empty(n_groups, dtype=torch.int32, device=plan.perm.device)
    _lib.call("mm_dispatch_group_map", _P(plan.perm), _P(plan.slot_row), lay.n_items, lay.topk, lay.P[0], lay.region_base[0],
              n_groups, _P(g64), _st())
    return g64


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    _need_cuda(src)
    src = src.contiguous()
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _lib.call("mm_cast_f32_bf16", _P(src), _P(dst), src.numel(), _st())
    return dst


def transpose_cast_bf16(src: torch.Tensor) -> torch.Tensor:
    """fp32 [batch, R, C] -> bf16 [batch, C, R]."""
    _need_cuda(src)
    src = src.contiguous()
    b, R, Cc = src.shape
    dst = torch.empty(b, Cc, R, dtype=torch.bfloat16, device=src.device)
    _lib.call("mm_transpose_cast_f32_bf16", _P(src), _P(dst), b, R, Cc, _st())
    return dst


def pack_expert_params(params: Sequence[torch.Tensor], E: int, S: int, widths: Sequence[int], D: int, H: int, need_T: bool,
                       launch_stream=None):
    """The experts' fp32 master parameters -> stacked kernel operands, in one launch (mm_pack_expert_params).
    params: per expert [conv_s.weight, conv_s.bias] * S + [attn0.weight, attn0.bias, attn2.weight, attn2.bias].
    Returns dict: Wp[s] bf16 [E*D, D_s], WpT[s] bf16 [E*D_s, D] | None, W1 bf16 [E*H, D], W1T bf16 [E*D, H] | None,
    bp[s] fp32 [E, D], b1 [E, H], w2 [E, H], b2 [E].
    `launch_stream`: enqueue the kernel there (the caller forks / joins); the buffers are allocated under the current stream."""
    per = 2 * S + 4
    dev = params[0].device
    keep = []                                     # fp32 / contiguous temporaries stay alive until the launch is enqueued

    def f32(t):
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.detach().float().contiguous()
            keep.append(t)
        return t
    n_w = sum(E * D * w for w in widths) + E * H * D
    wbuf = torch.empty(n_w * (2 if need_T else 1), dtype=torch.bfloat16, device=dev)
    vbuf = torch.empty(S * E * D + 2 * E * H + E, dtype=torch.float32, device=dev)
    out = {"Wp": [], "WpT": [], "bp": []}
    off = 0
    for s in range(S):
        out["Wp"].append(wbuf[off:off + E * D * widths[s]].view(E * D, widths[s])); off += E * D * widths[s]
    out["W1"] = wbuf[off:off + E * H * D].view(E * H, D); off += E * H * D
    if need_T:
        for s in range(S):
            out["WpT"].append(wbuf[off:off + E * D * widths[s]].view(E * widths[s], D)); off += E * D * widths[s]
        out["W1T"] = wbuf[off:off + E * H * D].view(E * D, H); off += E * H * D
    else:
        out["WpT"], out["W1T"] = [None] * S, None
    voff = 0
    for s in range(S):
        out["bp"].append(vbuf[voff:voff + E * D].view(E, D)); voff += E * D
    out["b1"] = vbuf[voff:voff + E * H].view(E, H); voff += E * H
    out["w2"] = vbuf[voff:voff + E * H].view(E, H); voff += E * H
    out["b2"] = vbuf[voff:voff + E]
    src, dst, dstT, rows, cols, kind = [], [], [], [], [], []

    def job(t, d, dT, r, c, k):
        src.append(f32(t).data_ptr()); dst.append(d.data_ptr()); dstT.append(dT.data_ptr() if dT is not None else 0)
        rows.append(r); cols.append(c); kind.append(k)
    for e in range(E):
        base = e * per
        for s in range(S):
            w = widths[s]
            job(params[base + 2 * s], out["Wp"][s][e * D:(e + 1) * D], out["WpT"][s][e * w:(e + 1) * w] if need_T else None, D, w, 0)
            job(params[base + 2 * s + 1], out["bp"][s][e], None, 1, D, 1)
        job(params[base + 2 * S], out["W1"][e * H:(e + 1) * H], out["W1T"][e * D:(e + 1) * D] if need_T else None, H, D, 0)
        job(params[base + 2 * S + 1], out["b1"][e], None, 1, H, 1)
        job(params[base + 2 * S + 2], out["w2"][e], None, 1, H, 1)
        job(params[base + 2 * S + 3], out["b2"][e:e + 1], None, 1, 1, 1)
    n = len(src)
    # (temporaries of non-fp32 parameters are freed right after the call, and the per-call event profiler records on the current
    # stream: both cases keep the launch there)
    st = launch_stream.cuda_stream if (launch_stream is not None and not keep and _lib.PROFILER is None) else _st()
    _lib.call("mm_pack_expert_params", (_lib.C.c_void_p * n)(*src), (_lib.C.c_void_p * n)(*dst), (_lib.C.c_void_p * n)(*dstT),
              _lib.host_i32(rows), _lib.host_i32(cols), _lib.host_i32(kind), n, st)
    del keep
    return out


# ---- grouped GEMMs --------------------------------------------------------------------
def gemm_rows(A: torch.Tensor, W: torch.Tensor, N: int, out: torch.Tensor, *, plan: Optional[DispatchPlan] = None,
              tile_begin: int = 0, tile_count: int = 0, M: int = 0, bias=None, aux=None, gate=None, colsum=None,
              flags: int = 0, out_scale: float = 1.0, tag: str = "", out_g64=None):
    """out[rows, N] = epi(A[rows, K] W_e[N, K]^T); W is the stacked [E * N, K] bf16 weight.
    `out_g64` (group map): `out` is in image order, the store un-permutes (plain bf16 epilogue, plan required)."""
    _need_cuda(A, W, out)
    E = W.shape[0] // N
    if out_g64 is not None:
        assert plan is not None and aux is None and gate is None and colsum is None and out.dtype == torch.bfloat16 and out_scale == 1.0
        _lib.call("mm_grouped_gemm_rows_scatter", _P(A), A.shape[0], A.shape[1], A.stride(0), _P(W), E, N, W.stride(0),
                  _P(plan.tile_info), tile_begin, tile_count, _P(bias), _P(out), out.shape[0], out.stride(0), _P(out_g64), flags,
                  _st(), label=f"{tag}:gemm_rows[K={A.shape[1]},N={N}]")
        return out
    _lib.call("mm_grouped_gemm_rows", _P(A), A.shape[0], A.shape[1], A.stride(0), _P(W), E, N, W.stride(0),
              _P(plan.tile_info) if plan is not None else 0, tile_begin, tile_count, M, _P(bias),
              _P(aux), aux.stride(0) if aux is not None else 0, _P(gate), gate.stride(0) if gate is not None else 0,
              _P(out), out.stride(0), int(out.dtype == torch.float32), _P(colsum), float(out_scale), flags, _st(),
              label=f"{tag}:gemm_rows[K={A.shape[1]},N={N}]")
    return out


# fused E1 -> E4 (csrc/b2b.cuh) for the scales it supports; switch kept for A/B measurements and the parity tests
USE_B2B_FWD = True


def b2b_pairs_available() -> bool:
    """CTA-pair back-to-back kernel not switched off (MEDMOE_B2B_DEBUG bit 0)."""
    import os
    return not (int(os.environ.get("MEDMOE_B2B_DEBUG", "0") or 0) & 1)


def expert_b2b_supported(K1: int, D: int, H: int) -> bool:
    return USE_B2B_FWD and bool(_lib.call("mm_expert_b2b_fwd_supported", K1, D, H))


def expert_b2b_fwd(f: torch.Tensor, Wp: torch.Tensor, bias1: torch.Tensor, W1: torch.Tensor, bias2: torch.Tensor,
                   Y: torch.Tensor, Z: torch.Tensor, *, plan: DispatchPlan, tile_begin: int, tile_count: int, tag: str = "",
                   pairs: bool = True, f_g64=None):
    """Y = ReLU(f Wp_e^T + bias1_e), Z = Y W1_e^T + bias2_e over the tiles of one scale region, in one kernel (Y is not
    re-read).  Wp: stacked [E * D, K1] bf16, W1: stacked [E * H, D] bf16; f / Y / Z start at the region's first row.
    `pairs`: the plan's 256-row segment alignment (plan.SEG_ALIGN) lets CTA pairs share the weight tiles (same results)."""
    _need_cuda(f, Wp, bias1, W1, bias2, Y, Z)
    D, H = Y.shape[1], Z.shape[1]
    E = Wp.shape[0] // D
    if f_g64 is not None:      # f is the image-order [B * P, K1] tensor, addressed through the group map
        _lib.call("mm_expert_b2b_fwd_gather", _P(f), f.shape[0], f.shape[1], f.stride(0), _P(Wp), E, D, Wp.stride(0), _P(bias1),
                  _P(W1), H, W1.stride(0), _P(bias2), _P(plan.tile_info), tile_begin, tile_count, _P(Y), Y.stride(0), _P(Z),
                  Z.stride(0), 1, _P(f_g64), _st(), label=f"{tag}:expert_b2b[K1={f.shape[1]}]")
        return Y, Z
    _lib.call("mm_expert_b2b_fwd", _P(f), f.shape[0], f.shape[1], f.stride(0), _P(Wp), E, D, Wp.stride(0), _P(bias1), _P(W1),
              H, W1.stride(0), _P(bias2), _P(plan.tile_info), tile_begin, tile_count, _P(Y), Y.stride(0), _P(Z), Z.stride(0),
              int(pairs and tile_begin % 2 == 0), _st(), label=f"{tag}:expert_b2b[K1={f.shape[1]}]")
    return Y, Z


def gemm_rows_rank1(A: torch.Tensor, W: torch.Tensor, N: int, out: torch.Tensor, *, plan: DispatchPlan, tile_begin: int,
                    tile_count: int, row_coef: torch.Tensor, row_vec: torch.Tensor, vecs: torch.Tensor, gate: torch.Tensor,
                    aux: Optional[torch.Tensor] = None, colsum=None, tag: str = "", flags: int = EPI_PAIR_OK):
    """out[row] = (A[row] W_e^T + row_coef[row] * vecs[row_vec[row]] [+ aux[row]]) * [gate[row] > 0]
    (the rank-1 aux is never materialised)."""
    _need_cuda(A, W, out, row_coef, row_vec, vecs, gate, aux)
    E = W.shape[0] // N
    _lib.call("mm_grouped_gemm_rows_rank1", _P(A), A.shape[0], A.shape[1], A.stride(0), _P(W), E, N, W.stride(0),
              _P(plan.tile_info), tile_begin, tile_count, _P(row_coef), _P(row_vec), _P(vecs), vecs.stride(0), _P(aux),
              aux.stride(0) if aux is not None else 0, _P(gate), gate.stride(0), _P(out), out.stride(0), _P(colsum), flags, _st(),
              label=f"{tag}:gemm_rows_rank1[K={A.shape[1]},N={N}]")
    return out


def gemm_wgrad(A: torch.Tensor, Bm: torch.Tensor, out: torch.Tensor, plan: DispatchPlan, chunk_begin: int,
               chunk_count: int, tile_base: int, tag: str = "", colsum: Optional[torch.Tensor] = None, b_g64=None):
    """out[e] (+)= A_e^T B_e over the rows of expert e; out fp32 [E, N1, N2] must be pre-zeroed.
    colsum (fp32 [E, N1], pre-zeroed): also accumulate the column sums of A per expert (bias gradient).
    `b_g64` (group map): Bm is in image order and is addressed through the map (needs colsum, tile_base 0)."""
    _need_cuda(A, Bm, out, colsum)
    if b_g64 is not None:
        _lib.call("mm_grouped_gemm_wgrad_colsum_gather", _P(A), A.shape[0], A.shape[1], A.stride(0), _P(Bm), Bm.shape[0],
                  Bm.shape[1], Bm.stride(0), _P(plan.chunks), chunk_begin, chunk_count, tile_base, _P(out), _P(colsum), _P(b_g64),
                  _st(), label=f"{tag}:gemm_wgrad[N1={A.shape[1]},N2={Bm.shape[1]}]")
        return out
    if colsum is None:
        _lib.call("mm_grouped_gemm_wgrad", _P(A), A.shape[0], A.shape[1], A.stride(0), _P(Bm), Bm.shape[0], Bm.shape[1],
                  Bm.stride(0), _P(plan.chunks), chunk_begin, chunk_count, tile_base, _P(out), _st(),
                  label=f"{tag}:gemm_wgrad[N1={A.shape[1]},N2={Bm.shape[1]}]")
    else:
        _lib.call("mm_grouped_gemm_wgrad_colsum", _P(A), A.shape[0], A.shape[1], A.stride(0), _P(Bm), Bm.shape[0],
                  Bm.shape[1], Bm.stride(0), _P(plan.chunks), chunk_begin, chunk_count, tile_base, _P(out), _P(colsum), _st(),
                  label=f"{tag}:gemm_wgrad[N1={A.shape[1]},N2={Bm.shape[1]}]")
    return out


# ---- combine --------------------------------------------------------------------------
def combine_fwd(Y, Z, w2, b2, plan: DispatchPlan, D: int, gate, out_dtype):
    lay = plan.layout
    _need_cuda(Y, Z, w2, b2)
    B, P = lay.n_images, lay.P[0]
    dev = Y.device
    nblk = _lib.call("mm_combine_num_token_blocks", P)
    beta = torch.empty(lay.n_items, P, 4, dtype=torch.float32, device=dev)
    out = torch.empty(B, P, D, dtype=out_dtype, device=dev)
    gpart = torch.empty(B, nblk, D, dtype=torch.float32, device=dev)
    gfeat = torch.empty(B, D, dtype=torch.float32, device=dev)
    tile0 = plan.tile_info[lay.tile_base[0]:lay.tile_base[0] + lay.region_tiles[0]]
    _lib.call("mm_interp_softmax_combine_fwd", _P(Y), _P(Z), _P(w2), _P(b2), B, lay.topk, P, _lib.host_i32(lay.P), D,
              _P(plan.inv_perm), _P(plan.slot_expert), _P(plan.slot_row), _P(gate), _P(beta), _P(out),
              int(out_dtype == torch.float32), _P(gpart), _P(gfeat), _P(plan.perm), _P(plan.seg_start), _P(plan.offsets),
              _P(tile0), lay.region_tiles[0], lay.region_base[0], lay.num_experts, Y.shape[0],
              int(FORCE_CUDA_CORE_COMBINE_FWD), _st())
    return out, gfeat, beta


def combine_bwd(Y, Z, w2, plan: DispatchPlan, D: int, gate, beta, dlocal, dglobal, need_dgate: bool,
                force_generic: bool = False, red=None):
    lay = plan.layout
    _need_cuda(Y, Z, w2, beta, dlocal, dglobal)
    B, P, K = lay.n_images, lay.P[0], lay.num_experts
    dev = Y.device
    f32 = dict(dtype=torch.float32, device=dev)
    nrb = _lib.call("mm_combine_num_part_blocks", P, _lib.host_i32(lay.P))
    nruns = _lib.call("mm_combine_num_runs", P)
    mom_u = torch.empty(lay.n_items, nruns, 2, D, **f32)
    mom_z = torch.empty(lay.n_items, _lib.call("mm_combine_bwd_z_scratch_floats", P, _lib.host_i32(lay.P), D), **f32)
    dlogit = torch.empty(lay.n_items, P, 8, **f32)   # dlogit [.., 4] (generic path) or the two dbeta halves [.., 2, 4]
    dgate = torch.zeros(lay.n_items, **f32) if need_dgate else None
    dUT = torch.empty(lay.total_rows, D, dtype=torch.bfloat16, device=dev)
    dZ = torch.empty(lay.total_rows, D // 2, dtype=torch.bfloat16, device=dev)
    part = torch.empty(lay.n_items, nrb, D + 1, **f32)
    if red is None:
        red = torch.empty(K, D + 1, **f32)
    dl_f32 = dlocal is not None and dlocal.dtype == torch.float32
    _lib.call("mm_interp_softmax_combine_bwd", _P(Y), _P(Z), _P(w2), B, lay.topk, P, _lib.host_i32(lay.P), D, K,
              _P(plan.perm), _P(plan.inv_perm), _P(plan.slot_expert), _P(plan.slot_row), _P(plan.counts),
              _P(plan.seg_start), _P(plan.offsets), _P(gate), _P(beta), _P(dlocal), int(dl_f32), _P(dglobal),
              _P(dlogit), _P(dgate), _P(dUT), _P(dZ), _P(part), _P(red), _P(mom_u), _P(mom_z),
              int(force_generic or FORCE_GENERIC_COMBINE_BWD), _st())
    H = D // 2
    return dUT, dZ, red[:, :H], red[:, H:2 * H], red[:, 2 * H], dgate


def combine_bwd_global_supported(plan: DispatchPlan, D: int) -> bool:
    lay = plan.layout
    return bool(_lib.call("mm_combine_bwd_global_supported", lay.P[0], _lib.host_i32(lay.P), D))


def combine_bwd_tc_supported(plan: DispatchPlan, D: int) -> bool:
    lay = plan.layout
    return bool(_lib.call("mm_combine_bwd_tc_supported", lay.P[0], _lib.host_i32(lay.P), D))


def combine_bwd_tc(Y, Z, w2, plan: DispatchPlan, D: int, gate, beta, dlocal, dglobal, need_dgate: bool, red=None):
    """Backward of the combine on the tensor-core / rank-1 path.  dlocal: bf16 [B, P, D] or None; dglobal: fp32 [B, D] or None.
    Returns (row_coef, row_img, dUT, dZ, dw2, db1, db2, dgate): d fused / d Y = dUT (local part, None without dlocal)
    + row_coef[row] * dglobal[row_img[row]] (rebuilt inside the dY GEMM, None without dglobal)."""
    lay = plan.layout
    _need_cuda(Y, Z, w2, beta, dlocal, dglobal)
    B, P, K = lay.n_images, lay.P[0], lay.num_experts
    dev = Y.device
    f32 = dict(dtype=torch.float32, device=dev)
    nrb = _lib.call("mm_combine_num_part_blocks", P, _lib.host_i32(lay.P))
    zscr = torch.empty(lay.n_items, _lib.call("mm_combine_bwd_z_scratch_floats", P, _lib.host_i32(lay.P), D), **f32)
    row_dot = row_coef = row_img = dbeta_loc = dUT = mom_u = None
    if dglobal is not None:
        row_dot = torch.empty(lay.total_rows, **f32)
        row_coef = torch.empty(lay.total_rows, **f32)
        row_img = torch.empty(lay.total_rows, dtype=torch.int32, device=dev)
    if dlocal is not None:
        if dlocal.dtype != torch.bfloat16 or not dlocal.is_contiguous():
            raise RuntimeError("combine_bwd_tc: dlocal must be a contiguous bf16 [B, P, D] tensor")
        dbeta_loc = torch.empty(lay.n_items, P, 4, **f32)
        dUT = torch.empty(lay.total_rows, D, dtype=torch.bfloat16, device=dev)
        mom_u = torch.empty(lay.n_items, _lib.call("mm_combine_num_runs", P), 2, D, **f32)
    dgate = torch.zeros(lay.n_items, **f32) if need_dgate else None
    dZ = torch.empty(lay.total_rows, D // 2, dtype=torch.bfloat16, device=dev)
    part = torch.empty(lay.n_items, nrb, D + 1, **f32)
    if red is None:
        red = torch.empty(K, D + 1, **f32)
    tile0 = plan.tile_info[lay.tile_base[0]:lay.tile_base[0] + lay.region_tiles[0]]
    _lib.call("mm_interp_softmax_combine_bwd_tc", _P(Y), _P(Z), _P(w2), B, lay.topk, P, _lib.host_i32(lay.P), D, K,
              _P(plan.perm), _P(plan.inv_perm), _P(plan.slot_expert), _P(plan.slot_row), _P(plan.counts),
              _P(plan.seg_start), _P(plan.offsets), _P(tile0), lay.region_tiles[0], lay.region_base[0], Y.shape[0],
              _P(gate), _P(beta), _P(dlocal), _P(dglobal), _P(row_dot), _P(row_coef), _P(row_img), _P(dbeta_loc), _P(dUT),
              _P(mom_u), _P(dgate), _P(dZ), _P(part), _P(red), _P(zscr), _st())
    H = D // 2
    return row_coef, row_img, dUT, dZ, red[:, :H], red[:, H:2 * H], red[:, 2 * H], dgate


# ---- losses ---------------------------------------------------------------------------
def gloria_fwd(img, txt, temp: float, eps: float):
    _need_cuda(img, txt)
    B, D = img.shape
    ws = torch.empty(_lib.call("mm_gloria_workspace_floats", B), dtype=torch.float32, device=img.device)
    loss = torch.empty((), dtype=torch.float32, device=img.device)
    _lib.call("mm_gloria_global_fwd", _P(img), _P(txt), B, D, float(temp), float(eps), _P(ws), _P(loss), _st())
    return loss, ws


def gloria_bwd(img, txt, temp: float, eps: float, ws, gout, need_dimg: bool, need_dtxt: bool):
    B, D = img.shape
    dimg = torch.empty_like(img) if need_dimg else None
    dtxt = torch.empty_like(txt) if need_dtxt else None
    _lib.call("mm_gloria_global_bwd", _P(img), _P(txt), B, D, float(temp), float(eps), _P(ws), _P(gout), _P(dimg),
              _P(dtxt), _st())
    return dimg, dtxt


def infonce_fwd(a, b_all, scale_exp, label0: int, row_w):
    _need_cuda(a, b_all, scale_exp)
    R, D = a.shape
    N = b_all.shape[0]
    f32 = dict(dtype=torch.float32, device=a.device)
    logits = torch.empty(R, N, **f32); lse = torch.empty(R, **f32); picked = torch.empty(R, **f32)
    loss = torch.empty((), **f32)
    _lib.call("mm_infonce_fwd", _P(a), _P(b_all), R, N, D, _P(scale_exp), label0, _P(row_w), _P(logits), _P(lse),
              _P(picked), _P(loss), _st())
    return loss, logits, lse


def infonce_bwd(a, b_all, scale_exp, label0: int, row_w, logits, lse, gout, gmul: float, dscale, accumulate_dscale: bool):
    R, D = a.shape
    N = b_all.shape[0]
    f32 = dict(dtype=torch.float32, device=a.device)
    dlogits = torch.empty(R, N, **f32); row_tmp = torch.empty(R, **f32)
    da = torch.empty(R, D, **f32); db_all = torch.empty(N, D, **f32)
    _lib.call("mm_infonce_bwd", _P(a), _P(b_all), R, N, D, _P(scale_exp), label0, _P(row_w), _P(logits), _P(lse),
              _P(gout), float(gmul), _P(dlogits), _P(row_tmp), _P(da), _P(db_all), _P(dscale), int(accumulate_dscale),
              _st())
    return da, db_all


def infonce_fused_supported(R: int, N: int, D: int) -> bool:
    return bool(_lib.call("mm_infonce_fused_supported", R, N, D))


def infonce_fused_fwd(a, b, all_a, all_b, scale_exp, label0: int, row_w, smoothing: float, want_logits: bool):
    """Both directions of the InfoNCE in one fused tcgen05 kernel (+ the bf16 split pre-pass).
    Returns (loss [2], logits_a | None, logits_b | None, lse [2, R], workspace)."""
    _need_cuda(a, b, all_a, all_b, scale_exp, row_w)
    R, D = a.shape
    N = all_b.shape[0]
    dev = a.device
    f32 = dict(dtype=torch.float32, device=dev)
    ws = torch.empty(_lib.call("mm_infonce_fused_workspace_bytes", R, N, D) + 1024, dtype=torch.uint8, device=dev)
    ws = ws[(-ws.data_ptr()) % 1024:]                       # 1 KB aligned view (torch allocations are 512 B aligned)
    logits_a = torch.empty(R, N, **f32) if want_logits else None
    logits_b = torch.empty(R, N, **f32) if want_logits else None
    lse = torch.empty(2, R, **f32)
    loss = torch.empty(2, **f32)
    _lib.call("mm_infonce_fused_fwd", _P(a), _P(b), _P(all_a), _P(all_b), R, N, D, _P(scale_exp), label0, _P(row_w),
              float(smoothing), _P(ws), _P(logits_a), _P(logits_b), _P(lse), _P(loss), _st())
    return loss, logits_a, logits_b, lse, ws


def infonce_fused_bwd(R: int, N: int, D: int, scale_exp, label0: int, row_w, smoothing: float, ws, lse, g_a, g_b, aliased: bool):
    """-> (da, db, dall_a | None, dall_b | None, dscale [1]).  `aliased` (single rank: all_a is a): the column-side gradients are
    accumulated into da / db in place."""
    dev = lse.device
    f32 = dict(dtype=torch.float32, device=dev)
    rows = 2 * R if aliased else 2 * R + 2 * N
    buf = torch.zeros(rows, D, **f32)                      # one memset for all four gradient buffers
    da, db = buf[:R], buf[R:2 * R]
    dall_a, dall_b = (da, db) if aliased else (buf[2 * R:2 * R + N], buf[2 * R + N:])
    dscale = torch.empty(1, **f32)
    _lib.call("mm_infonce_fused_bwd", R, N, D, _P(scale_exp), label0, _P(row_w), float(smoothing), _P(ws), _P(lse), _P(g_a),
              _P(g_b), _P(da), _P(db), _P(dall_a), _P(dall_b), _P(dscale), 0, _st())
    return da, db, (None if aliased else dall_a), (None if aliased else dall_b), dscale


def l2_normalize_fwd(x, eps: float = 1e-12):
    _need_cuda(x)
    R, D = x.shape
    y = torch.empty_like(x); n = torch.empty(R, dtype=torch.float32, device=x.device)
    _lib.call("mm_l2_normalize_fwd", _P(x), R, D, float(eps), _P(y), _P(n), _st())
    return y, n


def l2_normalize_bwd(dy, y, n, eps: float = 1e-12):
    R, D = y.shape
    dx = torch.empty_like(y)
    _lib.call("mm_l2_normalize_bwd", _P(dy), _P(y), _P(n), R, D, float(eps), _P(dx), _st())
    return dx


def zeroshot_argmax(img, txt, eps: float = 1e-8, return_sim: bool = False):
    _need_cuda(img, txt)
    M, D = img.shape
    Cn = txt.shape[0]
    pred = torch.empty(M, dtype=torch.int64, device=img.device)
    sim = torch.empty(M, Cn, dtype=torch.float32, device=img.device) if return_sim else None
    _lib.call("mm_zeroshot_argmax", _P(img), _P(txt), M, Cn, D, float(eps), _P(pred), _P(sim), _st())
    return (pred, sim) if return_sim else pred
