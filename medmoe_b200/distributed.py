"""Embedding exchange for the global contrastive loss — mirror of the reference's
`src/utils/distributed.py` (`BackpropType`, `gather_tensor`, `concat_gather_all_gpu`, `get_rank`).

One process per GPU.  The exchange is an all-gather forward and a reduce-scatter(SUM) backward — the pair the
reference gets from `torch.distributed.nn.functional.all_gather`, src/utils/distributed.py:47-48.  Messages are
tiny ([B_loc, 2 * 768] fp32 per rank), i.e. pure latency, so on one node (NCCL backend, <= 8 ranks, fp32) it runs
over NVLink PEER MEMORY with the library's own kernels (csrc/p2p.cu: peer stores + system-scope flags forward,
peer loads backward; `PeerExchange`), not through NCCL: ~10 us per exchange instead of ~80 us of NCCL launch and
protocol latency, CUDA-graph capturable, deterministic summation order.  Everything else (other dtypes, several
nodes, `MEDMOE_P2P_EXCHANGE=0`) uses one fused NCCL `all_gather_into_tensor` / `reduce_scatter_tensor` per tensor;
on backends without reduce-scatter (gloo, used by the CPU tests) the backward is all-reduce + slice, which is
arithmetically identical.
"""
from __future__ import annotations

import ctypes
import os
import socket
from enum import Enum
from typing import List, Optional

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


class PeerExchange:
    """All-gather / reduce-scatter(SUM) of one fixed message size between the ranks of one node over NVLink peer memory
    (csrc/p2p.cu).  Construction is COLLECTIVE (every rank of the default group, outside CUDA-graph capture): each rank
    allocates one workspace with cudaMalloc, the 64-byte CUDA IPC handles travel through `all_gather_object`, peers map
    them.  `PeerExchange.get(nbytes, device)` returns the cached instance, or None when the exchange cannot be used (then
    every rank takes the NCCL path: the decision is made from values that are equal on all ranks)."""

    MAX_WORLD = 8
    _cache: dict = {}
    _disabled_reason: Optional[str] = None
    last_backend: str = "none"           # "p2p" or "nccl": what the most recent exchange of this process used

    def __init__(self, nbytes: int, device: torch.device):
        from . import _lib
        self.nbytes, self.device = nbytes, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        total = _lib.call("mm_p2p_workspace_bytes", self.world, nbytes)
        handle = (ctypes.c_ubyte * 64)()
        own = ctypes.c_void_p()
        err = None
        with torch.cuda.device(device):
            try:
                _lib.call("mm_p2p_alloc", total, ctypes.byref(own), handle)
            except RuntimeError as ex:
                err = str(ex)
            infos = [None] * self.world
            dist.all_gather_object(infos, (socket.gethostname(), bytes(handle), err))
            self.ptrs: List[int] = [0] * self.world
            if err is None and all(i[2] is None for i in infos) and len({i[0] for i in infos}) == 1:
                for q, (_, h, _) in enumerate(infos):
                    if q == self.rank:
                        self.ptrs[q] = own.value
                        continue
                    p = ctypes.c_void_p()
                    try:
                        _lib.call("mm_p2p_open", (ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p))
                        self.ptrs[q] = p.value
                    except RuntimeError as ex:
                        err = str(ex)
                        break
            else:
                err = err or "a peer could not allocate its workspace, or the ranks span several hosts"
            oks = [None] * self.world
            dist.all_gather_object(oks, err)
            bad = [e for e in oks if e is not None]
            if bad:
                for q, p in enumerate(self.ptrs):
                    if p and q != self.rank:
                        _lib.load().mm_p2p_close(ctypes.c_void_p(p))
                if own.value:
                    _lib.load().mm_p2p_free(own)
                raise RuntimeError(bad[0])
            self._peer_array = (ctypes.c_void_p * self.world)(*self.ptrs)
            torch.cuda.synchronize(device)
            dist.barrier()

    @classmethod
    def get(cls, nbytes: int, device: torch.device) -> Optional["PeerExchange"]:
        if cls._disabled_reason is not None or os.environ.get("MEDMOE_P2P_EXCHANGE", "1") == "0":
            return None
        if dist.get_backend() != "nccl" or dist.get_world_size() > cls.MAX_WORLD or nbytes % 16 != 0:
            return None
        key = (nbytes, device.index)
        ex = cls._cache.get(key)
        if ex is None:
            if torch.cuda.is_current_stream_capturing():
                return None              # cannot rendezvous inside a capture: warm up eagerly first (bench.py does)
            try:
                ex = cls._cache[key] = PeerExchange(nbytes, device)
            except RuntimeError as err:   # raised on every rank alike (the error state is all-gathered)
                cls._disabled_reason = str(err)
                return None
        return ex

    def all_gather(self, x: torch.Tensor) -> torch.Tensor:
        from . import _lib
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        with torch.cuda.device(self.device):
            _lib.call("mm_p2p_all_gather", x.data_ptr(), out.data_ptr(), self.nbytes, self._peer_array, self.rank, self.world,
                      _lib.stream_ptr(), label="p2p_all_gather")
        PeerExchange.last_backend = "p2p"
        return out

    def reduce_scatter(self, g: torch.Tensor, rows: int) -> torch.Tensor:
        from . import _lib
        out = torch.empty((rows,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        with torch.cuda.device(self.device):
            _lib.call("mm_p2p_reduce_scatter_f32", g.data_ptr(), out.data_ptr(), self.nbytes, self._peer_array, self.rank,
                      self.world, _lib.stream_ptr(), label="p2p_reduce_scatter")
        return out


def _peer_exchange_for(x: torch.Tensor) -> Optional[PeerExchange]:
    if not x.is_cuda or x.dtype != torch.float32:
        return None
    return PeerExchange.get(x.numel() * x.element_size(), x.device)


class _AllGatherCat(torch.autograd.Function):
    """x [B, ...] on every rank -> concatenation [W * B, ...]; backward = reduce-scatter(SUM)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        world = dist.get_world_size()
        ctx.rows = x.shape[0]
        ctx.peer = _peer_exchange_for(x)
        if ctx.peer is not None:
            return ctx.peer.all_gather(x)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        PeerExchange.last_backend = "nccl"
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = grad_out.contiguous()
        if ctx.peer is not None and grad_out.dtype == torch.float32:
            return ctx.peer.reduce_scatter(grad_out, ctx.rows)
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
        peer = _peer_exchange_for(src)
        if peer is not None:
            out = peer.all_gather(src)
        else:
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



class OverlappedGradSync:
    """DDP-style gradient averaging for a data-parallel `medmoe_b200.MoE` (experts replicated) that overlaps
    the collective with the backward pass.

    The MoE backward writes every expert-parameter gradient into ONE flat fp32 buffer and calls `grad_ready_hook` as soon as
    the weight-gradient GEMMs are done — before the input-gradient GEMMs (dX), the un-permute and the router backward.  The
    hook starts the NCCL all-reduce (AVG) of that bucket on a communication stream, so it travels over NVLink while those
    kernels run; `finish()` (after `loss.backward()`) joins the stream and averages the few remaining parameters (router,
    logit_scale: ~0.1 M values) in one more small bucket.  Everything is event-ordered and CUDA-graph capturable.
    The reference gets the same arithmetic from Lightning's DDP (configs/trainer/ddp.yaml:4), after the whole backward.
    """

    def __init__(self, moe, other_params=(), process_group=None):
        self.moe = moe
        self.group = process_group
        self.expert_params = [p for ex in moe.experts for p in ex.parameters()]
        ids = {id(p) for p in self.expert_params}
        self.other_params = [p for p in list(moe.parameters()) + list(other_params) if id(p) not in ids]
        self.comm = None
        self._flat = None
        moe.grad_ready_hook = self._on_ready

    def _on_ready(self, flat: torch.Tensor) -> None:
        if not is_distributed():
            return
        if any(p.grad is not None for p in self.expert_params):
            return                      # gradient accumulation into existing .grad: fall back to the late all-reduce in finish()
        cur = torch.cuda.current_stream(flat.device)
        if self.comm is None:
            self.comm = torch.cuda.Stream(device=flat.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        self._flat = flat

    def finish(self) -> None:
        """Call after backward: waits for the early all-reduce and averages whatever it did not cover."""
        if not is_distributed():
            return
        late = list(self.other_params)
        flat = self._flat
        self._flat = None
        if flat is not None:
            torch.cuda.current_stream(flat.device).wait_stream(self.comm)
            lo = flat.data_ptr()
            hi = lo + flat.numel() * flat.element_size()
            # autograd normally keeps the views of the flat buffer as .grad (no copy); anything else is averaged late
            late += [p for p in self.expert_params if p.grad is not None and not (lo <= p.grad.data_ptr() < hi)]
        else:
            late += self.expert_params
        grads = [p.grad for p in late if p.grad is not None]
        if grads:
            bucket = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(bucket, op=dist.ReduceOp.AVG, group=self.group)
            torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(bucket, grads)))
