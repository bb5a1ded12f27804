"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement, in plain PyTorch ops on fp32/fp64 tensors, of the reference's MoE block:
`Expert.forward` (reference src/models/components/swin.py:32-80) and `MoE.forward`
(swin.py:94-117).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.

Two executions of the same arithmetic are provided:

* `moe_forward_dense`   — the reference's as-written schedule: every expert runs on every
  image, outputs are stacked [B, K, P, D] and one is gathered per image (swin.py:105-108);
  interpolation through `F.interpolate(mode="linear", align_corners=False)` (swin.py:42).
  This is what the CPU baseline times.
* `moe_forward_sparse`  — route first, run only the selected expert(s); interpolation by
  the closed form of SURVEY §8a row a4; generalised to top-k (k=1 reduces exactly to the
  reference gather).  This is the schedule the CUDA path uses.

Parity pinning: the reference ships no tests or golden vectors ("parity unpinned" by the
reference itself).  The oracle is pinned instead against outputs of the reference's own
classes run in the build container: tests/golden/*.npz (made by tests/golden/make_golden.py
from /root/reference) and, when /root/reference is present, live in tests/test_oracle_vs_reference.py.

Parameters are passed as a flat dict keyed by the reference's state_dict names
(`experts.{e}.proj_convs.{s}.0.weight`, ... `router.2.bias`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


def num_experts_of(params: Params) -> int:
    return params["router.2.weight"].shape[0]


def init_params(num_experts: int = 6, hidden_dims: Sequence[int] = (96, 192, 384, 768), output_dim: int = 768,
                router_input_dim: int = 768, seed: int = 0, dtype=torch.float32) -> Params:
    """Random parameters with nn.Linear / nn.Conv1d default statistics (U(-1/sqrt(fan_in), +))
    (reference swin.py:83-92 builds these modules with PyTorch defaults).  Not bit-identical to
    the reference's RNG consumption — tests copy state_dicts when they need identical weights."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)

    p: Params = {}
    for e in range(num_experts):
        for s, d in enumerate(hidden_dims):
            p[f"experts.{e}.proj_convs.{s}.0.weight"] = uni((output_dim, d, 1), d)
            p[f"experts.{e}.proj_convs.{s}.0.bias"] = uni((output_dim,), d)
        h = output_dim // 2
        p[f"experts.{e}.attn_proj.0.weight"] = uni((h, output_dim), output_dim)
        p[f"experts.{e}.attn_proj.0.bias"] = uni((h,), output_dim)
        p[f"experts.{e}.attn_proj.2.weight"] = uni((1, h), h)
        p[f"experts.{e}.attn_proj.2.bias"] = uni((1,), h)
    p["router.0.weight"] = uni((128, router_input_dim), router_input_dim)
    p["router.0.bias"] = uni((128,), router_input_dim)
    p["router.2.weight"] = uni((num_experts, 128), 128)
    p["router.2.bias"] = uni((num_experts,), 128)
    return p


# --------------------------------------------------------------------------------------
# interpolation closed form (ATen upsample_linear1d, align_corners=False)
# --------------------------------------------------------------------------------------
def lerp_indices(p_src: int, p_dst: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """i0, i1 (int64 [p_dst]) and lambda (float32 [p_dst]) such that
    out[j] = (1 - lam[j]) * x[i0[j]] + lam[j] * x[i1[j]]."""
    scale = torch.tensor(p_src, dtype=torch.float32) / torch.tensor(p_dst, dtype=torch.float32)
    j = torch.arange(p_dst, dtype=torch.float32)
    src = torch.clamp(scale * (j + 0.5) - 0.5, min=0.0)
    i0 = torch.clamp(src.floor().long(), max=p_src - 1)
    lam = torch.clamp(src - i0.to(torch.float32), 0.0, 1.0)
    i1 = i0 + (i0 < p_src - 1).long()
    return i0, i1, lam


def lerp_rows(x: torch.Tensor, p_dst: int) -> torch.Tensor:
    """x [..., P_src, C] -> [..., p_dst, C] along the token axis."""
    i0, i1, lam = lerp_indices(x.shape[-2], p_dst)
    lam = lam.to(x.dtype).unsqueeze(-1)
    return (1 - lam) * x.index_select(-2, i0) + lam * x.index_select(-2, i1)


# --------------------------------------------------------------------------------------
# router (swin.py:98-100) and top-k gate (extension; k = 1 is the reference)
# --------------------------------------------------------------------------------------
def router_probs(params: Params, swin_feat: torch.Tensor) -> torch.Tensor:
    h = torch.relu(F.linear(swin_feat, params["router.0.weight"], params["router.0.bias"]))
    return torch.softmax(F.linear(h, params["router.2.weight"], params["router.2.bias"]), dim=-1)


def topk_gate(probs: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """indices [B, k] (first maximum wins, like torch.argmax) and gate weights [B, k]
    (1.0 for k = 1, else selected probs renormalised to sum 1)."""
    B, K = probs.shape
    work = probs.detach().clone()
    idx = []
    for _ in range(k):
        i = torch.argmax(work, dim=-1)
        idx.append(i)
        work[torch.arange(B), i] = -1.0
    idx = torch.stack(idx, dim=1)
    if k == 1:
        return idx, torch.ones(B, 1, dtype=probs.dtype)
    sel = probs.gather(1, idx)
    return idx, sel / sel.sum(dim=1, keepdim=True)


# --------------------------------------------------------------------------------------
# one expert
# --------------------------------------------------------------------------------------
def expert_forward_as_written(params: Params, e: int, feats: List[torch.Tensor]) -> torch.Tensor:
    """swin.py:32-80 op for op: conv1d(k=1)+ReLU, F.interpolate, stack/permute, attn MLP, softmax, weighted sum."""
    max_len = max(f.shape[1] for f in feats)
    ups = []
    for s, f in enumerate(feats):
        y = torch.relu(F.conv1d(f.transpose(1, 2), params[f"experts.{e}.proj_convs.{s}.0.weight"],
                                params[f"experts.{e}.proj_convs.{s}.0.bias"]))          # [B, D, P_s]
        ups.append(F.interpolate(y, size=max_len, mode="linear", align_corners=False))     # [B, D, P]
    stacked = torch.stack(ups, dim=0).permute(1, 3, 0, 2)                                  # [B, P, S, D]
    B, P, S, D = stacked.shape
    x = stacked.reshape(B * P * S, D)
    h = torch.relu(F.linear(x, params[f"experts.{e}.attn_proj.0.weight"], params[f"experts.{e}.attn_proj.0.bias"]))
    logit = F.linear(h, params[f"experts.{e}.attn_proj.2.weight"], params[f"experts.{e}.attn_proj.2.bias"]).view(B, P, S)
    beta = torch.softmax(logit, dim=-1)
    return (stacked * beta.unsqueeze(-1)).sum(dim=2)                                       # [B, P, D]


def _stored(x: torch.Tensor, dtype) -> torch.Tensor:
    """Value as it reads back from a `dtype` buffer; the rounding is invisible to autograd (straight-through)."""
    return x if dtype is None else x + (x.to(dtype).to(x.dtype) - x).detach()


def expert_forward_storage_aware(params: Params, e: int, feats: List[torch.Tensor], storage=torch.bfloat16) -> torch.Tensor:
    """The closed form below with the two HBM storage points of the CUDA path made explicit: Y = ReLU(conv) and
    Z = Y W1^T + b1 (first attention Linear, evaluated at native resolution because the lerp commutes with the affine map)
    are rounded to `storage` before they are interpolated.  The ReLU gate of the attention hidden layer is then decided on
    the same values the kernels see, which is what the gradient of attn_proj.0 is sensitive to (tests/test_moe_gpu.py:
    a unit within bf16 rounding of zero flips its gate, a 100 % error on that entry).  Mathematically identical to
    `expert_forward_closed_form` for storage=None."""
    P = max(f.shape[1] for f in feats)
    W1, b1 = params[f"experts.{e}.attn_proj.0.weight"], params[f"experts.{e}.attn_proj.0.bias"]
    us, hs = [], []
    for s, f in enumerate(feats):
        w = params[f"experts.{e}.proj_convs.{s}.0.weight"].squeeze(-1)
        y = _stored(torch.relu(f @ w.t() + params[f"experts.{e}.proj_convs.{s}.0.bias"]), storage)   # [B, P_s, D]
        z = _stored(y @ W1.t() + b1, storage)                                                       # [B, P_s, H]
        us.append(lerp_rows(y, P))
        hs.append(torch.relu(lerp_rows(z, P)))
    U = torch.stack(us, dim=2)                                                             # [B, P, S, D]
    h = torch.stack(hs, dim=2)                                                             # [B, P, S, H]
    logit = (h @ params[f"experts.{e}.attn_proj.2.weight"].t()).squeeze(-1) + params[f"experts.{e}.attn_proj.2.bias"]
    beta = torch.softmax(logit, dim=-1)
    return (U * beta.unsqueeze(-1)).sum(dim=2)


def expert_forward_closed_form(params: Params, e: int, feats: List[torch.Tensor]) -> torch.Tensor:
    """Same expert with the projection as a matmul and the interpolation in closed form
    (token-major throughout, no [B, D, P] transposes) — the schedule the CUDA path restates."""
    P = max(f.shape[1] for f in feats)
    us = []
    for s, f in enumerate(feats):
        w = params[f"experts.{e}.proj_convs.{s}.0.weight"].squeeze(-1)                     # [D, D_s]
        y = torch.relu(f @ w.t() + params[f"experts.{e}.proj_convs.{s}.0.bias"])           # [B, P_s, D]
        us.append(lerp_rows(y, P))                                                         # [B, P, D]
    U = torch.stack(us, dim=2)                                                             # [B, P, S, D]
    h = torch.relu(U @ params[f"experts.{e}.attn_proj.0.weight"].t() + params[f"experts.{e}.attn_proj.0.bias"])
    logit = (h @ params[f"experts.{e}.attn_proj.2.weight"].t()).squeeze(-1) + params[f"experts.{e}.attn_proj.2.bias"]
    beta = torch.softmax(logit, dim=-1)                                                    # [B, P, S]
    return (U * beta.unsqueeze(-1)).sum(dim=2)


# --------------------------------------------------------------------------------------
# the MoE block
# --------------------------------------------------------------------------------------
def _finish(fused: torch.Tensor, probs: torch.Tensor):
    B, P, D = fused.shape
    H = W = int(P ** 0.5)                                                                  # swin.py:111
    global_feat = fused.mean(dim=1)
    local_feat = fused.transpose(1, 2).reshape(B, D, H, W)
    return global_feat, local_feat, probs


def moe_forward_dense(params: Params, feats: List[torch.Tensor], swin_feat: torch.Tensor):
    """Reference schedule (swin.py:94-117): all K experts on all images, stack, gather top-1."""
    probs = router_probs(params, swin_feat)
    top = torch.argmax(probs, dim=-1)
    outs = torch.stack([expert_forward_as_written(params, e, feats) for e in range(num_experts_of(params))], dim=1)
    fused = outs[torch.arange(outs.size(0)), top, :, :]
    return _finish(fused, probs)


def moe_forward_sparse(params: Params, feats: List[torch.Tensor], swin_feat: torch.Tensor, topk: int = 1,
                       closed_form: bool = True, storage=None):
    """Route first, run only the selected experts on their images.  topk > 1 is the documented
    extension: fused = sum_j g_j * expert_{idx_j}(x), g = renormalised top-k probabilities.
    `storage=torch.bfloat16` evaluates the experts with the CUDA path's storage points (expert_forward_storage_aware)."""
    probs = router_probs(params, swin_feat)
    idx, gate = topk_gate(probs, topk)
    if topk > 1:
        sel = probs.gather(1, idx)
        gate = sel / sel.sum(dim=1, keepdim=True)          # differentiable w.r.t. the router
    B = swin_feat.shape[0]
    P = max(f.shape[1] for f in feats)
    D = params["experts.0.attn_proj.0.weight"].shape[1]
    fused = torch.zeros(B, P, D, dtype=feats[0].dtype)
    run = expert_forward_closed_form if closed_form else expert_forward_as_written
    if storage is not None:
        run = lambda p_, e_, f_: expert_forward_storage_aware(p_, e_, f_, storage)      # noqa: E731
    for e in range(num_experts_of(params)):
        for j in range(topk):
            rows = (idx[:, j] == e).nonzero().flatten()
            if rows.numel() == 0:
                continue
            out = run(params, e, [f.index_select(0, rows) for f in feats])
            fused = fused.index_add(0, rows, out * gate[rows, j].view(-1, 1, 1).to(out.dtype))
    return _finish(fused, probs), idx


def router_ce(probs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """medmoe_module.py:235-237: cross-entropy applied to the already-softmaxed router output."""
    return F.cross_entropy(probs, labels)
