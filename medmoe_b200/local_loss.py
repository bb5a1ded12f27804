"""GLORIALocalContrastiveLoss on the B200 kernels (reference src/losses.py:954-1026 + attention_fn :698-736).

The reference loops over the B captions in Python, repeats each caption's words B times and runs two `bmm`s and two
softmaxes per caption.  Here all B x B (image, caption) pairs are six large GEMMs on the tcgen05 grouped kernels plus four
streaming passes (csrc/local_loss.cu):

    forward   S   = ctx  words^T                      dense rows GEMM      [B*Ppad, 768] x [N, 768]^T     (N = captions * Wp)
              E   = exp(temp1 * softmax_w S)          mm_local_softmax_exp_fwd
              wcU = E_b^T ctx_b per image             grouped wgrad GEMM   (images play the role of experts)
              cos, sim = cosine + log-sum-exp         mm_local_cos_lse_fwd
    backward  dwcU, dwords(direct)                    mm_local_cos_lse_bwd
              dE  = ctx_b dwcU_b^T                    grouped rows GEMM
              dS                                      mm_local_softmax_exp_bwd (in place over dE)
              dctx = [E | dS] [dwcU_b ; words]        ONE grouped rows GEMM over K = 2N (E and dS share a buffer), fp32 out
              dwords = dS^T ctx + direct part         dense wgrad GEMM

The softmax over patches is `E / colsum(E)`; the cosine that consumes the attended context is scale free, so the column sums
are never formed (they only matter for the returned attention maps, which are computed for the B matching pairs alone).
Captions may have up to 128 words (the reference tokenises to max_length 25, configs/model/med-moe.yaml:40; up to 32 words take
the fused score-softmax epilogue).  Captions are processed in column blocks so that the fp32 score matrix of a block stays below `SCORE_BYTES_BUDGET`.
The two cross-entropies over the B x B similarity matrix are ordinary torch ops (65 k elements).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib, ops

_P = _lib.ptr
SCORE_BYTES_BUDGET = 6 << 30
FUSE_SCORE_SOFTMAX = True       # captions of at most 32 words: first softmax in the score GEMM's epilogue (tests switch it off)


def _st():
    return _lib.stream_ptr()


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class GLORIALocalContrastiveLossOutput(OrderedDict):
    loss0: Tensor
    loss1: Tensor
    att_maps: List[Tensor]


_TABLE_CACHE: dict = {}


def _cap_len_tensor(lens, caps: int, dev) -> Tensor:
    """int32 [caps] on the device, cached per length tuple: repeated steps (and CUDA graph capture) do no host-to-device copy."""
    key = ("len", tuple(lens), caps, str(dev))
    t = _TABLE_CACHE.get(key)
    if t is None:
        if len(_TABLE_CACHE) > 64:
            _TABLE_CACHE.clear()
        host = torch.zeros(caps, dtype=torch.int32)
        host[:len(lens)] = torch.tensor(lens, dtype=torch.int32)
        t = _TABLE_CACHE[key] = host.to(dev)
    return t


def _tables(B: int, tiles_per_image: int, dev):
    """tile -> image table for the grouped row GEMMs, per-image and whole-range chunk lists for the wgrad GEMMs (cached)."""
    key = ("tab", B, tiles_per_image, str(dev))
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = _build_tables(B, tiles_per_image, dev)
    return _TABLE_CACHE[key]


def _build_tables(B: int, tiles_per_image: int, dev):
    tiles = B * tiles_per_image
    t = torch.arange(tiles, dtype=torch.int32)
    tile_info = torch.stack([t // tiles_per_image, torch.full_like(t, 128)], dim=1).contiguous().to(dev)
    img_chunks, all_chunks = [], []
    for b in range(B):
        for c in range(0, tiles_per_image, 64):
            img_chunks.append((b, b * tiles_per_image + c, min(64, tiles_per_image - c), 0))
    for c in range(0, tiles, 64):
        all_chunks.append((0, c, min(64, tiles - c), 0))
    mk = lambda rows: torch.tensor(rows, dtype=torch.int32).contiguous().to(dev)   # noqa: E731
    return (SimpleNamespace(tile_info=tile_info), SimpleNamespace(chunks=mk(img_chunks)), len(img_chunks),
            SimpleNamespace(chunks=mk(all_chunks)), len(all_chunks))


class _LocalSimilarity(torch.autograd.Function):
    """(tokens [B, P, D], words [B, L, D]) -> sim [B, B] with sim[b, i] = log sum_w exp(temp2 * cos(word_iw, attended context))."""

    @staticmethod
    def forward(ctx, tokens: Tensor, words: Tensor, cap_lens: Sequence[int], temp1: float, temp2: float, agg_mean: bool):
        if not tokens.is_cuda:
            raise RuntimeError("medmoe_b200 runs on CUDA tensors only; there is no CPU fallback")
        dev = tokens.device
        B, P, D = tokens.shape
        L = words.shape[1]
        lens = [max(0, min(int(n), L)) for n in cap_lens]
        if len(lens) != words.shape[0] or words.shape[0] != B:
            raise RuntimeError("GLORIALocalContrastiveLoss expects one caption (and one length) per image")
        Wp = _round_up(max(max(lens), 1), 8)
        Lc = min(L, Wp)
        caps = _round_up(B, 16)                      # N = caps * Wp is a multiple of 128 (GEMM tile widths)
        N = caps * Wp
        Ppad = _round_up(P, 128)
        tpi = Ppad // 128
        rows = B * Ppad

        ctx16 = torch.zeros(B, Ppad, D, dtype=torch.bfloat16, device=dev)
        ctx16[:, :P].copy_(tokens)
        ctx16 = ctx16.view(rows, D)
        words32 = torch.zeros(caps, Wp, D, dtype=torch.float32, device=dev)
        words32[:B, :Lc].copy_(words[:, :Lc])
        words32 = words32.view(N, D)
        words16 = words32.to(torch.bfloat16)
        cap_len = _cap_len_tensor(lens, caps, dev)
        img_tiles, img_chunks, n_img_chunks, all_chunks, n_all_chunks = _tables(B, tpi, dev)

        # caption blocks: 16 captions at least, as many as keep the fp32 scores of a block under the budget
        per_cap = rows * Wp * 4
        cb = max(16, min(caps, (SCORE_BYTES_BUDGET // max(per_cap, 1)) // 16 * 16))
        fused_scores = Wp == 32 and FUSE_SCORE_SOFTMAX     # one epilogue chunk per caption: the scores never leave the SM
        if fused_scores:
            cb = caps
        blocks = [(c0, min(caps, c0 + cb)) for c0 in range(0, caps, cb)]

        sim = torch.empty(B, caps, dtype=torch.float32, device=dev)
        # [E | dE]: the backward's d ctx GEMM multiplies both halves in one pass (K = 2N), so they share a buffer
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        ES = torch.empty(rows, 2 * N if need_grad else N, dtype=torch.bfloat16, device=dev)
        wcUs, coss = [], []
        for c0, c1 in blocks:
            Nb = (c1 - c0) * Wp
            wblk = words16[c0 * Wp:c1 * Wp]
            E = ES[:, c0 * Wp:c1 * Wp]
            if fused_scores:
                _lib.call("mm_local_scores_softmax_exp", _P(ctx16), rows, D, D, _P(wblk), c1 - c0, D, _P(cap_len[c0:c1]),
                          float(temp1), _P(E), ES.stride(0), _st(), label=f"LL.S+softmax_exp:gemm_rows[K={D},N={Nb}]")
            else:
                S = torch.empty(rows, Nb, dtype=torch.float32, device=dev)
                ops.gemm_rows(ctx16, wblk, Nb, S, M=rows, tag="LL.S")
                _lib.call("mm_local_softmax_exp_fwd", _P(S), Nb, _P(E), ES.stride(0), rows, c1 - c0, Wp, _P(cap_len[c0:c1]),
                          float(temp1), _st(), label="LL.softmax_exp")
                del S
            wcU = torch.zeros(B, Nb, D, dtype=torch.float32, device=dev)
            ops.gemm_wgrad(E, ctx16, wcU, img_chunks, 0, n_img_chunks, 0, tag="LL.wc")
            cosv = torch.empty(B, Nb, dtype=torch.float32, device=dev)
            _lib.call("mm_local_cos_lse_fwd", _P(wcU), _P(words32[c0 * Wp:c1 * Wp]), B, c1 - c0, Wp, D, _P(cap_len[c0:c1]),
                      float(temp2), int(agg_mean), _P(cosv), _P(sim[:, c0:]), caps, _st(), label="LL.cos_lse")
            wcUs.append(wcU); coss.append(cosv)

        ctx.geom = (B, P, D, L, Lc, Wp, caps, N, Ppad, rows, blocks, float(temp1), float(temp2), bool(agg_mean))
        ctx.tables = (img_tiles, all_chunks, n_all_chunks)
        ctx.saved = (ctx16, words32, words16, cap_len, sim, ES, wcUs, coss)
        ctx.in_dtypes = (tokens.dtype, words.dtype)
        return sim[:, :B].clone()

    @staticmethod
    def backward(ctx, dsim: Tensor):
        B, P, D, L, Lc, Wp, caps, N, Ppad, rows, blocks, temp1, temp2, agg_mean = ctx.geom
        img_tiles, all_chunks, n_all_chunks = ctx.tables
        ctx16, words32, words16, cap_len, sim, ES, wcUs, coss = ctx.saved
        dev = ctx16.device
        dsim_pad = torch.zeros(B, caps, dtype=torch.float32, device=dev)
        dsim_pad[:, :B].copy_(dsim)
        dw_direct = torch.empty(N, D, dtype=torch.float32, device=dev)
        tiles = rows // 128
        ld = ES.stride(0)
        # per-image weights of the d ctx GEMM: [dwcU_b^T | words^T], matching the [E | dS] columns of ES
        Wcat = torch.empty(B, D, 2 * N, dtype=torch.bfloat16, device=dev)
        Wcat[:, :, N:] = words16.t().contiguous()          # transpose the 12 MB once, then a row-contiguous broadcast
        for (c0, c1), wcU, cosv in zip(blocks, wcUs, coss):
            nc = c1 - c0
            Nb = nc * Wp
            dwcU = torch.empty(B, Nb, D, dtype=torch.bfloat16, device=dev)
            _lib.call("mm_local_cos_lse_bwd", _P(dsim_pad[:, c0:]), caps, _P(sim[:, c0:]), caps, _P(cosv), _P(wcU),
                      _P(words32[c0 * Wp:c1 * Wp]), B, nc, Wp, D, _P(cap_len[c0:c1]), temp2, int(agg_mean), _P(dwcU),
                      _P(Wcat[:, :, c0 * Wp:]), 2 * N, _P(dw_direct[c0 * Wp:c1 * Wp]), _st(), label="LL.cos_lse_bwd")
            # dE[(b, p), n] = <ctx[(b, p)], dwcU[b, n]>: every image multiplies its own [Nb, D] matrix
            dE = ES[:, N + c0 * Wp:N + c1 * Wp]
            ops.gemm_rows(ctx16, dwcU.view(B * Nb, D), Nb, dE, plan=img_tiles, tile_begin=0, tile_count=tiles, tag="LL.dE")
            _lib.call("mm_local_softmax_exp_bwd", _P(ES[:, c0 * Wp:c1 * Wp]), ld, _P(dE), ld, rows, nc, Wp, _P(cap_len[c0:c1]),
                      temp1, _st(), label="LL.softmax_exp_bwd")
            del dwcU
        dS = ES[:, N:]
        # d ctx = E dwcU_b + dS words in one grouped GEMM over K = 2N
        dctx = torch.empty(rows, D, dtype=torch.float32, device=dev)
        ops.gemm_rows(ES, Wcat.view(B * D, 2 * N), D, dctx, plan=img_tiles, tile_begin=0, tile_count=tiles, tag="LL.dctx")
        del Wcat
        # d words = dS^T ctx (+ the direct part through the cosine)
        dwords = torch.zeros(1, N, D, dtype=torch.float32, device=dev)
        ops.gemm_wgrad(dS, ctx16, dwords, all_chunks, 0, n_all_chunks, 0, tag="LL.dwords")
        dwords = dwords[0] + dw_direct
        tok_dtype, word_dtype = ctx.in_dtypes
        d_tokens = dctx.view(B, Ppad, D)[:, :P].to(tok_dtype)
        d_words = torch.zeros(B, L, D, dtype=word_dtype, device=dev)
        d_words[:, :Lc] = dwords.view(caps, Wp, D)[:B, :Lc].to(word_dtype)
        return d_tokens, d_words, None, None, None, None


def local_similarities(img_features: Tensor, words_emb: Tensor, cap_lens: Sequence[int], temp1: float = 4.0,
                       temp2: float = 5.0, agg: str = "sum") -> Tensor:
    """sim [B, B] before temp3 (rows: images, columns: captions).  img_features [B, D, H, W], words_emb [B, D, L]."""
    B, D = img_features.shape[:2]
    tokens = img_features.permute(0, 2, 3, 1).reshape(B, -1, D)       # a view when local_feat comes from medmoe_b200.MoE
    words = words_emb.permute(0, 2, 1)
    return _LocalSimilarity.apply(tokens, words, list(cap_lens), float(temp1), float(temp2), agg != "sum")


@torch.no_grad()
def attention_maps(img_features: Tensor, words_emb: Tensor, cap_lens: Sequence[int], temp1: float = 4.0) -> List[Tensor]:
    """att_maps[i] = attention of caption i's words over image i's patches, [1, cap_len_i, H, W] (losses.py:987-989)."""
    B, D, ih, iw = img_features.shape
    ctxt = img_features.reshape(B, D, -1).float()
    maps = []
    scores = torch.bmm(ctxt.transpose(1, 2), words_emb.float())       # [B, P, L]: only the B matching pairs
    for i in range(B):
        n = max(0, min(int(cap_lens[i]), words_emb.shape[2]))
        a = torch.softmax(scores[i, :, :n], dim=-1).t()
        maps.append(torch.softmax(a * temp1, dim=-1).reshape(1, n, ih, iw))
    return maps


class GLORIALocalContrastiveLoss(nn.Module):
    """Same call signature and outputs as the reference class (losses.py:954-1026); `idx` and `probs` are ignored there too."""

    def __init__(self, return_att_maps: bool = True):
        super().__init__()
        self.return_att_maps = return_att_maps

    def forward(self, img_features: Tensor, words_emb: Tensor, cap_lens: Sequence[int], temp1: float = 4.0,
                temp2: float = 5.0, temp3: float = 10.0, agg: str = "sum", idx: Optional[int] = None,
                probs: Optional[Tensor] = None) -> GLORIALocalContrastiveLossOutput:
        sim = local_similarities(img_features, words_emb, cap_lens, temp1, temp2, agg) * temp3
        labels = torch.arange(img_features.shape[0], device=sim.device)
        loss0 = F.cross_entropy(sim, labels)                  # losses.py:1019
        loss1 = F.cross_entropy(sim.t(), labels)              # losses.py:1020
        att = attention_maps(img_features, words_emb, cap_lens, temp1) if self.return_att_maps else []
        return GLORIALocalContrastiveLossOutput(loss0=loss0, loss1=loss1, att_maps=att)
