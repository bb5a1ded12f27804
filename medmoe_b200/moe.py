"""`MoE` — drop-in for the reference's modality-specialised mixture-of-experts block.

Mirrors `src/models/components/swin.py` of the reference:
  * `Expert(hidden_dims, output_dim)`                       swin.py:11-30   (parameter container, same submodule names)
  * `MoE(num_experts=6, hidden_dims=[96,192,384,768], output_dim=768, router_input_dim=768)`   swin.py:83-92
  * `MoE.forward(multi_scale_feats, swin_feat) -> (global_feat, local_feat, router_probs)`     swin.py:94-117
with identical `state_dict` keys (`experts.{e}.proj_convs.{s}.0.weight`, `experts.{e}.attn_proj.{0,2}.*`,
`router.{0,2}.*`), identical default initialisation (the same nn modules are constructed in the
same order, so a given torch seed yields the same weights as the reference) and the same
autograd-visible behaviour (zero — not None — gradients for experts no image selected; the
router is trained only through the returned probabilities).

The arithmetic runs in hand-written sm_100a kernels behind the C-ABI (`medmoe_b200._lib`):
router gate -> dispatch (counting sort by expert) -> grouped tcgen05 GEMMs -> fused
interpolate/softmax/combine, and the matching backward.  bf16 operands, fp32 accumulation;
router, softmaxes and reductions in fp32.  There is no CPU path.

`topk > 1` is an extension the reference does not have: out = sum_j g_j expert_j(x)
with g the renormalised top-k router probabilities; `topk=1` (default) is the reference.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import ops
from .plan import build_plan, make_layout


class Expert(nn.Module):
    """Parameter container with the reference's layout (swin.py:11-30).

    The per-expert arithmetic is executed by `MoE.forward` on the fused CUDA path; calling an
    expert on its own routes every image to it through the same kernels."""

    def __init__(self, hidden_dims: Sequence[int] = (96, 192, 384, 768), output_dim: int = 768):
        super().__init__()
        self.output_dim = output_dim
        self.num_scales = len(hidden_dims)
        self.proj_convs = nn.ModuleList([
            nn.Sequential(nn.Conv1d(dim, output_dim, kernel_size=1), nn.ReLU()) for dim in hidden_dims
        ])
        self.attn_proj = nn.Sequential(
            nn.Linear(output_dim, output_dim // 2),
            nn.ReLU(),
            nn.Linear(output_dim // 2, 1),
        )

    def forward(self, multi_scale_feats):
        B = multi_scale_feats[0].shape[0]
        item_expert = torch.zeros(B, 1, dtype=torch.int32, device=multi_scale_feats[0].device)
        out, _ = _run_experts([self], list(multi_scale_feats), item_expert, None, 1)
        return out


_SIDE_STREAMS: dict = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


def _expert_param_list(experts: Sequence[Expert]) -> List[torch.Tensor]:
    ps: List[torch.Tensor] = []
    for ex in experts:
        for seq in ex.proj_convs:
            ps += [seq[0].weight, seq[0].bias]
        ps += [ex.attn_proj[0].weight, ex.attn_proj[0].bias, ex.attn_proj[2].weight, ex.attn_proj[2].bias]
    return ps


class _ExpertsFunction(torch.autograd.Function):
    """(feats..., expert params...) -> (fused [B, P, D], global_feat [B, D]) for a given routing."""

    @staticmethod
    def forward(ctx, item_expert, gate, topk, num_experts, n_scales, owner, *tensors):
        feats = list(tensors[:n_scales])
        params = tensors[n_scales:]
        per = 2 * n_scales + 4
        assert len(params) == num_experts * per
        dev = feats[0].device
        B = feats[0].shape[0]
        P = [f.shape[1] for f in feats]
        widths = [f.shape[2] for f in feats]
        if P[0] != max(P):
            raise RuntimeError("stage features must be ordered finest first (reference swin.py:139)")
        E, S = num_experts, n_scales
        if S != 4:
            raise RuntimeError("medmoe_b200 supports exactly 4 feature scales (reference default hidden_dims)")
        D = params[0].shape[0]
        H = D // 2
        in_dtype = feats[0].dtype
        feat_needs_grad = [f.requires_grad for f in feats]
        param_needs_grad = any(p.requires_grad for p in params)

        # ---- bf16 shadows of the fp32 master weights, stacked over experts (+ transposes for dgrad): ONE launch, on a side
        #      stream: it depends on nothing but the parameters, so it runs under the counting sort and the row permutation
        #      (fork / join by events: capturable in a CUDA graph) ----
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(cur)
        pk = ops.pack_expert_params(params, E, S, widths, D, H, need_T=param_needs_grad or any(feat_needs_grad),
                                    launch_stream=side)
        Wp, W1, bp, b1, w2, b2 = pk["Wp"], pk["W1"], pk["bp"], pk["b1"], pk["w2"], pk["b2"]

        layout = make_layout(B, topk, E, P)
        plan = build_plan(item_expert.reshape(-1).contiguous(), layout)
        feats_c = [f.contiguous() for f in feats]
        # The finest scale (53 % of the permuted bytes) needs no sorted copy: P_0 is a multiple of 64 and segments start on 256-row
        # boundaries, so its consumers address the image-order tensor through a 64-row group map (top-1, bf16 features, CTA-pair
        # back-to-back kernel).  fs[0] / the sorted d f_0 then do not exist.
        direct0 = (ops.USE_DIRECT_FINEST and topk == 1 and in_dtype == torch.bfloat16 and P[0] % 64 == 0 and
                   ops.expert_b2b_supported(widths[0], D, H) and ops.b2b_pairs_available() and layout.region_tiles[0] >= 2)
        g64 = ops.group_map(plan) if direct0 else None
        feat0_2d = feats_c[0].view(B * P[0], widths[0]) if direct0 else None
        fs = ops.dispatch_rows(feats_c, plan, widths, first_scale=1 if direct0 else 0)
        cur.wait_stream(side)

        Y = torch.empty(layout.total_rows, D, dtype=torch.bfloat16, device=dev)
        Z = torch.empty(layout.total_rows, H, dtype=torch.bfloat16, device=dev)
        # E1: Y_s = ReLU(f_s W_s^T + b_s);  E4 at native resolution: Z = Y W1^T + b1 (the lerp commutes with the affine map).
        # Narrow scales run both in ONE back-to-back kernel (Y never re-read); the others as two grouped GEMMs, with E4
        # as a single launch over the contiguous tile range the fused kernel did not cover.
        fused_s = [s for s in range(S) if ops.expert_b2b_supported(widths[s], D, H)]
        n_fused = 0
        while n_fused < S and n_fused in fused_s:      # a prefix of the regions, so that the rest stays one tile range
            n_fused += 1
        for s in range(n_fused):
            r0, nr = layout.region_base[s], layout.region_rows[s]
            ops.expert_b2b_fwd(feat0_2d if (direct0 and s == 0) else fs[s], Wp[s], bp[s], W1, b1, Y[r0:r0 + nr], Z[r0:r0 + nr],
                               plan=plan, tile_begin=layout.tile_base[s], tile_count=layout.region_tiles[s], tag=f"E1E4.s{s}",
                               f_g64=g64 if s == 0 else None)

        def e1(s):
            r0 = layout.region_base[s]
            return lambda: ops.gemm_rows(fs[s], Wp[s], D, Y[r0:r0 + layout.region_rows[s]], plan=plan,
                                         tile_begin=layout.tile_base[s], tile_count=layout.region_tiles[s], bias=bp[s],
                                         flags=ops.EPI_RELU | ops.EPI_ZERO_PAD | ops.EPI_PAIR_OK, tag=f"E1.s{s}")
        ops.run_scales([e1(s) for s in range(n_fused, S)])
        if n_fused < S:
            t0 = layout.tile_base[n_fused]
            r0 = layout.region_base[n_fused]
            ops.gemm_rows(Y[r0:], W1, H, Z[r0:], plan=plan, tile_begin=t0, tile_count=layout.total_tiles - t0, bias=b1,
                          flags=ops.EPI_ZERO_PAD | ops.EPI_PAIR_OK, tag="E4")
        gate_flat = gate.reshape(-1).float().contiguous() if gate is not None else None
        fused, gfeat, beta = ops.combine_fwd(Y, Z, w2, b2, plan, D, gate_flat, in_dtype)

        ctx.plan, ctx.layout = plan, layout
        ctx.dims = (B, E, S, D, H, widths, topk, per, in_dtype)
        ctx.saved = (fs, Y, Z, beta, w2, pk["W1T"], pk["WpT"], gate_flat)
        ctx.direct0 = (g64, feat0_2d)
        if owner is not None:
            owner.last_direct_finest = direct0
        ctx.feat_needs_grad = feat_needs_grad
        ctx.param_needs_grad = param_needs_grad
        ctx.gate_needs_grad = gate is not None and gate.requires_grad
        ctx.owner = owner
        ctx.set_materialize_grads(False)
        return fused, gfeat      # global_feat stays fp32 (the token mean is accumulated in fp32; only local_feat follows the input dtype)

    @staticmethod
    def backward(ctx, dfused, dglobal):
        plan, layout = ctx.plan, ctx.layout
        B, E, S, D, H, widths, topk, per, in_dtype = ctx.dims
        fs, Y, Z, beta, w2, W1T, WsT, gate_flat = ctx.saved
        g64, feat0_2d = ctx.direct0
        dev = Y.device
        n_in = 6 + S + E * per
        if dfused is None and dglobal is None:
            return (None,) * n_in
        if dfused is not None:
            if dfused.dtype not in (torch.float32, torch.bfloat16):
                dfused = dfused.float()
            dfused = dfused.contiguous()
        dglobal32 = dglobal.float().contiguous() if dglobal is not None else None

        # ONE fp32 buffer holds every expert-parameter gradient (one memset instead of ten fills; the per-parameter gradients
        # handed to autograd are contiguous views of it, and a DDP-style hook can all-reduce it as a single bucket while
        # the rest of this backward still runs):  dW1 [E,H,D] | dWp_s [E,D,D_s] | dbp_s [E,D] | red [E, D+1] = (dw2, db1, db2)
        sizes = [E * H * D] + [E * D * w for w in widths] + [E * D] * S + [E * (D + 1)]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        views, off = [], 0
        for n in sizes:
            views.append(flat[off:off + n]); off += n
        dW1 = views[0].view(E, H, D)
        dWp = [views[1 + s].view(E, D, widths[s]) for s in range(S)]
        dbp = [views[1 + S + s].view(E, D) for s in range(S)]      # conv bias gradients = column sums of dPre (ones block of dWp)
        red = views[1 + 2 * S].view(E, D + 1)

        need_dpre = ctx.param_needs_grad or any(ctx.feat_needs_grad)
        use_tc = not ops.FORCE_GENERIC_COMBINE_BWD and (
            ops.combine_bwd_global_supported(plan, D) if dfused is None else ops.combine_bwd_tc_supported(plan, D))
        if use_tc:
            # tensor-core / rank-1 path: dF = dlocal + dglobal / P.  The per-image constant part of d fused / d Y is rank-1
            # (rebuilt in the dY epilogue, never written); the local part goes through a tcgen05 GEMM (dbeta) and dUT.
            dlocal16 = None
            if dfused is not None:
                dlocal16 = dfused if dfused.dtype == torch.bfloat16 else ops.cast_bf16(dfused)
            row_coef, row_img, dUT, dZ, dw2, db1, db2, dgate = ops.combine_bwd_tc(Y, Z, w2, plan, D, gate_flat, beta, dlocal16,
                                                                                  dglobal32, ctx.gate_needs_grad, red=red)
            # one launch over the whole row space: every scale shares W1, and the rank-1 tables are indexed by global row
            if not need_dpre:
                dPre = None          # frozen experts and features: only the gate weights (top-k > 1) want a gradient
            elif dglobal32 is not None:
                dPre = dUT if dUT is not None else torch.empty(layout.total_rows, D, dtype=torch.bfloat16, device=dev)
                ops.gemm_rows_rank1(dZ, W1T, D, dPre, plan=plan, tile_begin=0, tile_count=layout.total_tiles, row_coef=row_coef,
                                    row_vec=row_img, vecs=dglobal32, gate=Y, aux=dUT, tag="dY")
            else:
                ops.gemm_rows(dZ, W1T, D, dUT, plan=plan, tile_begin=0, tile_count=layout.total_tiles, aux=dUT, gate=Y,
                              flags=ops.EPI_ZERO_PAD | ops.EPI_PAIR_OK, tag="dY")
                dPre = dUT
        else:
            dUT, dZ, dw2, db1, db2, dgate = ops.combine_bwd(Y, Z, w2, plan, D, gate_flat, beta, dfused, dglobal32,
                                                            ctx.gate_needs_grad, red=red)
            # dPre = (dUT + dZ W1) * [Y > 0], in place over dUT
            if need_dpre:
                ops.gemm_rows(dZ, W1T, D, dUT, plan=plan, tile_begin=0, tile_count=layout.total_tiles, aux=dUT, gate=Y,
                              flags=ops.EPI_ZERO_PAD | ops.EPI_PAIR_OK, tag="dY")
            dPre = dUT

        # weight gradients first: once they are complete the flat bucket can travel (NCCL on a side stream) while dX runs
        grads_params: List = [None] * (E * per)
        if ctx.param_needs_grad:
            ops.gemm_wgrad(dZ, Y, dW1, plan, 0, layout.total_chunks, 0, tag="dW1")

            def dwp(s):
                r0, nr = layout.region_base[s], layout.region_rows[s]
                if s == 0 and g64 is not None:      # B operand = the image-order stage feature, through the group map
                    return lambda: ops.gemm_wgrad(dPre[r0:r0 + nr], feat0_2d, dWp[s], plan, layout.chunk_base[s], layout.chunk_cap[s],
                                                  layout.tile_base[s], tag=f"dWp.s{s}", colsum=dbp[s], b_g64=g64)
                return lambda: ops.gemm_wgrad(dPre[r0:r0 + nr], fs[s], dWp[s], plan, layout.chunk_base[s], layout.chunk_cap[s],
                                              layout.tile_base[s], tag=f"dWp.s{s}", colsum=dbp[s])
            ops.run_scales([dwp(s) for s in range(S)])
            for e in range(E):
                for s in range(S):
                    grads_params[e * per + 2 * s] = dWp[s][e].unsqueeze(-1)       # Conv1d weight [D, D_s, 1]
                    grads_params[e * per + 2 * s + 1] = dbp[s][e]
                grads_params[e * per + 2 * S] = dW1[e]
                grads_params[e * per + 2 * S + 1] = db1[e]
                grads_params[e * per + 2 * S + 2] = dw2[e].unsqueeze(0)           # Linear(H, 1) weight [1, H]
                grads_params[e * per + 2 * S + 3] = db2[e].reshape(1)
            owner = ctx.owner
            if owner is not None:
                owner.last_flat_grad = flat
                if owner.grad_ready_hook is not None:
                    owner.grad_ready_hook(flat)

        grads_feats: List = [None] * S
        if any(ctx.feat_needs_grad):
            d0 = g64 is not None          # the finest scale's gradient is stored in image order by its GEMM (no sorted copy, no un-permute)
            dfs = [None if (d0 and s == 0) else torch.empty(layout.region_rows[s], widths[s], dtype=torch.bfloat16, device=dev)
                   for s in range(S)]
            dfeat0 = torch.empty(B * layout.P[0], widths[0], dtype=torch.bfloat16, device=dev) if d0 else None

            def dx(s):   # df_s = dPre_s W_s
                r0, nr = layout.region_base[s], layout.region_rows[s]
                if d0 and s == 0:
                    return lambda: ops.gemm_rows(dPre[r0:r0 + nr], WsT[s], widths[s], dfeat0, plan=plan, tile_begin=layout.tile_base[s],
                                                 tile_count=layout.region_tiles[s], tag=f"dX.s{s}", out_g64=g64)
                return lambda: ops.gemm_rows(dPre[r0:r0 + nr], WsT[s], widths[s], dfs[s], plan=plan, flags=ops.EPI_PAIR_OK,
                                             tile_begin=layout.tile_base[s], tile_count=layout.region_tiles[s], tag=f"dX.s{s}")
            ops.run_scales([dx(s) for s in range(S)])
            outs = ops.undispatch_rows(dfs, plan, widths, in_dtype, first_scale=1 if d0 else 0)
            if d0:
                outs[0] = dfeat0.view(B, layout.P[0], widths[0])
            grads_feats = [o if need else None for o, need in zip(outs, ctx.feat_needs_grad)]
        dgate_out = dgate.view(B, topk) if dgate is not None else None
        return (None, dgate_out, None, None, None, None, *grads_feats, *grads_params)


def _run_experts(experts, feats, item_expert, gate, topk, owner=None):
    params = _expert_param_list(experts)
    return _ExpertsFunction.apply(item_expert, gate, topk, len(experts), len(feats), owner, *feats, *params)


class _RouterFunction(torch.autograd.Function):
    """swin_feat -> (probs [B, K] fp32, topk_idx [B, k] int32, topk_w [B, k]); probs is differentiable."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, topk):
        x32 = x.float().contiguous()
        W1c, b1c, W2c, b2c = (t.float().contiguous() for t in (W1, b1, W2, b2))
        hidden, probs, idx, w, near_tie = ops.router_topk(x32, W1c, b1c, W2c, b2c, topk)
        ctx.save_for_backward(x32, W1c, W2c, hidden, probs)
        ctx.x_dtype = x.dtype
        ctx.need_dx = x.requires_grad
        ctx.mark_non_differentiable(idx, w, near_tie)
        ctx.set_materialize_grads(False)
        return probs, idx, w, near_tie

    @staticmethod
    def backward(ctx, dprobs, _didx, _dw, _dnt):
        if dprobs is None:
            return (None,) * 6
        x32, W1c, W2c, hidden, probs = ctx.saved_tensors
        dx, dW1, db1, dW2, db2 = ops.router_bwd(dprobs.float().contiguous(), probs, hidden, x32, W1c, W2c, ctx.need_dx)
        return (dx.to(ctx.x_dtype) if dx is not None else None), dW1, db1, dW2, db2, None


class MoE(nn.Module):
    def __init__(self, num_experts: int = 6, hidden_dims: Sequence[int] = (96, 192, 384, 768), output_dim: int = 768,
                 router_input_dim: int = 768, topk: int = 1):
        super().__init__()
        hidden_dims = list(hidden_dims)
        self.experts = nn.ModuleList([Expert(hidden_dims, output_dim) for _ in range(num_experts)])
        self.router = nn.Sequential(
            nn.Linear(router_input_dim, 128),
            nn.ReLU(),
            nn.Linear(128, num_experts),
        )
        self.topk = int(topk)
        self.last_top_expert = None   # int32 [B, topk] of the most recent forward (device tensor, no sync)
        self.last_near_tie = None     # int32 [B]: 1 where the routing margin was < ops.NEAR_TIE_TOL (device tensor, no sync)
        # Every expert-parameter gradient of a backward pass lives in ONE flat fp32 buffer (`last_flat_grad`; the .grad
        # tensors are views of it).  `grad_ready_hook(flat)` is called from inside the backward as soon as that buffer is
        # complete — before the input gradients are computed — so that data-parallel training can start its all-reduce
        # early (medmoe_b200.distributed.OverlappedGradSync).
        self.last_flat_grad = None
        self.grad_ready_hook = None
        self.last_direct_finest = False   # the most recent forward addressed the finest stage feature in image order (no sorted copy)

    def near_tie_count(self) -> int:
        """Images of the most recent forward whose expert choice hangs on a probability gap < 1e-6 (synchronises).
        The reference takes `argmax` of fp32 probabilities (swin.py:99-100); the router here is fp32 too, so the
        assignment is identical unless an image is flagged here (documented near-ties)."""
        return 0 if self.last_near_tie is None else int(self.last_near_tie.sum().item())

    def forward(self, multi_scale_feats, swin_feat):
        """multi_scale_feats: list of 4 tensors [B, P_s, D_s] (finest first); swin_feat [B, router_input_dim].
        Returns (global_feat [B, D], local_feat [B, D, sqrt(P), sqrt(P)], router_probs [B, K])."""
        feats = list(multi_scale_feats)
        if not feats[0].is_cuda:
            raise RuntimeError("medmoe_b200.MoE runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if any(f.dtype != feats[0].dtype for f in feats):
            # autocast hands over a mix (fp32 from LayerNorm, bf16 from Linear): the kernels compute in bf16 anyway
            feats = [f.to(torch.bfloat16) for f in feats]
        with torch.cuda.device(feats[0].device):     # the C-ABI launches on the current device's stream
            probs, idx, w, near_tie = _RouterFunction.apply(swin_feat, self.router[0].weight, self.router[0].bias,
                                                            self.router[2].weight, self.router[2].bias, self.topk)
            self.last_top_expert, self.last_near_tie = idx, near_tie
            gate = None
            if self.topk > 1:
                # extension: renormalised top-k probabilities, differentiable w.r.t. the router (tiny torch ops)
                sel = probs.gather(1, idx.long())
                gate = sel / sel.sum(dim=1, keepdim=True)
            fused, global_feat = _run_experts(list(self.experts), feats, idx, gate, self.topk, owner=self)
        B, P, D = fused.shape
        Hh = Ww = int(P ** 0.5)                                       # swin.py:111
        local_feat = fused.transpose(1, 2).reshape(B, D, Hh, Ww)      # a stride view, as in the reference
        return global_feat, local_feat, probs
