"""Embedding exchange for the global contrastive loss — mirror of the reference's
`src/utils/distributed.py` (`BackpropType`, `gather_tensor`, `concat_gather_all_gpu`, `get_rank`).

One process per GPU; the collectives are NCCL over NVLink 5 / NVSwitch through
`torch.distributed` (all-gather forward, reduce-scatter(SUM) backward — the same pair the
reference gets from `torch.distributed.nn.functional.all_gather`, src/utils/distributed.py:47-48).
Messages are tiny ([B_loc, 768] per rank), i.e. latency-bound: one fused
`all_gather_into_tensor` per tensor instead of a list all-gather + `torch.cat`.
On backends without reduce-scatter (gloo, used by the CPU tests) the backward falls back to
all-reduce + slice, which is arithmetically identical.
"""
from __future__ import annotations

from enum import Enum
from typing import List

import torch
import torch.distributed as dist


class BackpropType(Enum):
    """How gradients flow through the gather (reference src/utils/distributed.py:16-25)."""
    GLOBAL = 0   # to every worker (all-gather fwd, reduce-scatter bwd)
    LOCAL = 1    # only into this worker's own slice
    NONE = 2     # no gradient


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if is_distributed() else 0


def get_world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


class _AllGatherCat(torch.autograd.Function):
    """x [B, ...] on every rank -> concatenation [W * B, ...]; backward = reduce-scatter(SUM)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        world = dist.get_world_size()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = grad_out.contiguous()
        rank = dist.get_rank()
        grad_in = torch.empty((ctx.rows,) + tuple(grad_out.shape[1:]), dtype=grad_out.dtype, device=grad_out.device)
        if dist.get_backend() == "gloo":
            # gloo has no reduce-scatter: all-reduce then take the local slice (same sum).  The path is chosen from the
            # backend up front, so a failing NCCL collective raises instead of silently diverging from its peers.
            dist.all_reduce(grad_out, op=dist.ReduceOp.SUM)
            grad_in = grad_out[rank * ctx.rows:(rank + 1) * ctx.rows].clone()
        else:
            dist.reduce_scatter_tensor(grad_in, grad_out, op=dist.ReduceOp.SUM)
        return grad_in


def all_gather_cat(tensor: torch.Tensor, backprop_type: BackpropType = BackpropType.GLOBAL) -> torch.Tensor:
    """Concatenation over ranks along dim 0 (== torch.cat(gather_tensor(...)) of the reference)."""
    if not is_distributed():
        return tensor
    if backprop_type == BackpropType.GLOBAL:
        return _AllGatherCat.apply(tensor)
    with torch.no_grad():
        world, rank = dist.get_world_size(), dist.get_rank()
        src = tensor.detach().contiguous()
        out = torch.empty((world * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        dist.all_gather_into_tensor(out, src)
    if backprop_type == BackpropType.LOCAL:
        # gradients only into this worker's own block (reference :55-56 re-inserts the local tensor)
        B = tensor.shape[0]
        out = torch.cat([out[:rank * B], tensor, out[(rank + 1) * B:]], dim=0)
    return out


def gather_tensor(tensor: torch.Tensor, backprop_type: BackpropType = BackpropType.GLOBAL) -> List[torch.Tensor]:
    """List of every rank's tensor (reference src/utils/distributed.py:28-58)."""
    world = get_world_size()
    return list(all_gather_cat(tensor, backprop_type).chunk(world, dim=0))


def concat_gather_all_gpu(tensor: torch.Tensor, backprop_type: BackpropType = BackpropType.GLOBAL, dim: int = 0) -> torch.Tensor:
    """Reference src/utils/distributed.py:61-82."""
    if not is_distributed():
        return tensor
    if dim == 0:
        return all_gather_cat(tensor, backprop_type)
    return torch.cat(gather_tensor(tensor, backprop_type), dim=dim)
