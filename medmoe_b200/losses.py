"""Global image-text contrastive losses — drop-ins for the reference's `src/losses.py`.

  * `GLORIAGlobalContrastiveLoss().forward(cnn_code, rnn_code, temp3=10.0, idx=None, probs=None)`
        reference losses.py:757-794 — the configured loss (configs/model/med-moe_pretraining.yaml:29-31):
        cosine similarity x temp3, CE over rows + CE over columns (summed, no /2), local batch only.
  * `contrastive_loss_with_temperature(embeddings_a, embeddings_b, logit_scale, mask=None,
        backprop_type=GLOBAL, cross_entropy_kwargs=None) -> ContrastiveLossOutput`
        reference losses.py:527-592 (+ `_gather_embeddings_and_labels`, :503-524): learnable
        temperature, embeddings all-gathered across ranks, labels offset by rank, mean of the two CEs.
  * `FLAVAGlobalContrastiveLoss(logit_scale=None, ...).forward(image_sequence, text_sequence, mask=None)`
        reference losses.py:248-301.
  * `zero_shot_predict(image_embeddings, text_embeddings)` — BASELINE config 5.

All arithmetic (similarity GEMM, log-sum-exp, cross-entropy, and the whole backward) runs in
the sm_100a kernels of csrc/loss.cu through the C-ABI; the all-gather / reduce-scatter of the
embeddings is NCCL via `medmoe_b200.distributed`.  fp32 throughout.  No CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, Optional, Union

import torch
from torch import Tensor, nn

from . import ops
from .distributed import BackpropType, all_gather_cat, get_rank, is_distributed

DEFAULT_LOGIT_SCALE = math.log(1 / 0.07)
# test hook: run the fp32 CUDA-core InfoNCE kernels (csrc/loss.cu) instead of the fused tensor-core ones
FORCE_UNFUSED_INFONCE = False


# --------------------------------------------------------------------------------------
# GLORIA global loss
# --------------------------------------------------------------------------------------
class _GloriaGlobalFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, temp, eps):
        img32, txt32 = img.float().contiguous(), txt.float().contiguous()
        loss, ws = ops.gloria_fwd(img32, txt32, temp, eps)
        ctx.save_for_backward(img32, txt32, ws)
        ctx.cfg = (temp, eps, img.dtype, txt.dtype, img.requires_grad, txt.requires_grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        img32, txt32, ws = ctx.saved_tensors
        temp, eps, dt_i, dt_t, need_i, need_t = ctx.cfg
        dimg, dtxt = ops.gloria_bwd(img32, txt32, temp, eps, ws, gout.float().contiguous(), need_i, need_t)
        return (dimg.to(dt_i) if dimg is not None else None, dtxt.to(dt_t) if dtxt is not None else None, None, None)


class GLORIAGlobalContrastiveLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.eps = 1e-8       # losses.py:763
        self.temp3 = 10.0     # losses.py:764 (the forward argument is what is used, as in the reference)

    def forward(self, cnn_code: Tensor, rnn_code: Tensor, temp3: float = 10.0, idx: int = None, probs: Tensor = None) -> Tensor:
        # idx / probs are accepted and ignored, exactly like the reference (losses.py:771-772)
        if cnn_code.dim() == 3 and cnn_code.shape[0] == 1:
            cnn_code, rnn_code = cnn_code[0], rnn_code[0]
        if cnn_code.dim() != 2 or cnn_code.shape != rnn_code.shape:
            raise RuntimeError("GLORIAGlobalContrastiveLoss expects cnn_code and rnn_code of the same [B, D] shape")
        if not cnn_code.is_cuda:
            raise RuntimeError("medmoe_b200 losses run on CUDA tensors only; there is no CPU fallback")
        return _GloriaGlobalFunction.apply(cnn_code, rnn_code, float(temp3), float(self.eps))


# --------------------------------------------------------------------------------------
# FLAVA / CLIP-style loss with temperature
# --------------------------------------------------------------------------------------
@dataclass
class ContrastiveLossOutput:
    loss: Tensor
    logits_a: Tensor
    logits_b: Tensor
    loss_a: Tensor
    loss_b: Tensor


@dataclass
class FLAVAGlobalContrastiveLossOutput:
    text_embedding: Tensor
    image_embedding: Tensor
    logit_scale: Tensor
    image_logits: Tensor
    text_logits: Tensor
    image_loss: Tensor
    text_loss: Tensor
    loss: Tensor


class _InfoNCEFunction(torch.autograd.Function):
    """(a, b, all_a, all_b, logit_scale[, row_w]) -> (loss_a, loss_b, logits_a, logits_b).

    loss_a = sum_r w_r CE(exp(scale) a_r . all_b^T, label0 + r); loss_b symmetric.  Gradients are
    returned for a, b, all_a, all_b separately so that the all-gather's own backward (reduce-scatter)
    routes the cross-rank part.  Widths that are multiples of 192 (768 in every configuration) run the fused
    tensor-core kernels (csrc/infonce_fused.cu: one launch per pass for both directions, logits only written
    when `want_logits`); other widths take the fp32 CUDA-core kernels of csrc/loss.cu."""

    @staticmethod
    def forward(ctx, a, b, all_a, all_b, logit_scale, label0, row_w, smoothing=0.0, want_logits=True):
        a32, b32 = a.float().contiguous(), b.float().contiguous()
        aliased = all_a is a and all_b is b
        alla32 = a32 if all_a is a else all_a.float().contiguous()
        allb32 = b32 if all_b is b else all_b.float().contiguous()
        scale_exp = torch.exp(logit_scale.detach().float()).reshape(1).contiguous()
        R, D = a32.shape
        N = allb32.shape[0]
        ctx.fused = (not FORCE_UNFUSED_INFONCE) and a32.shape == b32.shape and alla32.shape == allb32.shape and \
            ops.infonce_fused_supported(R, N, D)
        ctx.cfg = (label0, row_w, a.dtype, b.dtype, all_a.dtype, all_b.dtype, logit_scale.dtype, logit_scale.shape, smoothing, aliased)
        ctx.set_materialize_grads(False)
        if ctx.fused:
            loss, logits_a, logits_b, lse, ws = ops.infonce_fused_fwd(a32, b32, alla32, allb32, scale_exp, label0, row_w,
                                                                      smoothing, want_logits)
            ctx.save_for_backward(scale_exp, lse, ws)
            ctx.dims = (R, N, D)
            loss_a, loss_b = loss[0], loss[1]
            if not want_logits:
                return loss_a, loss_b, None, None
            ctx.mark_non_differentiable(logits_a, logits_b)
            return loss_a, loss_b, logits_a, logits_b
        if smoothing:
            raise NotImplementedError("label_smoothing needs the fused InfoNCE kernels (embedding width a multiple of 192)")
        loss_a, logits_a, lse_a = ops.infonce_fwd(a32, allb32, scale_exp, label0, row_w)
        loss_b, logits_b, lse_b = ops.infonce_fwd(b32, alla32, scale_exp, label0, row_w)
        ctx.save_for_backward(a32, b32, alla32, allb32, scale_exp, logits_a, logits_b, lse_a, lse_b)
        ctx.mark_non_differentiable(logits_a, logits_b)
        return loss_a, loss_b, logits_a, logits_b

    @staticmethod
    def backward(ctx, g_a, g_b, _gla, _glb):
        label0, row_w, dt_a, dt_b, dt_alla, dt_allb, dt_s, shape_s, smoothing, aliased = ctx.cfg

        def cast(t, dt):
            return t.to(dt) if t is not None else None

        def scalar(g):
            return g.float().reshape(1).contiguous() if g is not None else None
        if ctx.fused:
            scale_exp, lse, ws = ctx.saved_tensors
            R, N, D = ctx.dims
            da, db, dall_a, dall_b, dscale = ops.infonce_fused_bwd(R, N, D, scale_exp, label0, row_w, smoothing, ws, lse,
                                                                   scalar(g_a), scalar(g_b), aliased)
            return (cast(da, dt_a), cast(db, dt_b), cast(dall_a, dt_alla), cast(dall_b, dt_allb),
                    dscale.reshape(shape_s).to(dt_s), None, None, None, None)
        a32, b32, alla32, allb32, scale_exp, logits_a, logits_b, lse_a, lse_b = ctx.saved_tensors
        dev = a32.device
        dscale = torch.zeros(1, dtype=torch.float32, device=dev)
        da = db = dall_a = dall_b = None
        if g_a is not None:
            da, dall_b = ops.infonce_bwd(a32, allb32, scale_exp, label0, row_w, logits_a, lse_a, scalar(g_a), 1.0, dscale, True)
        if g_b is not None:
            db, dall_a = ops.infonce_bwd(b32, alla32, scale_exp, label0, row_w, logits_b, lse_b, scalar(g_b), 1.0, dscale, True)
        return (cast(da, dt_a), cast(db, dt_b), cast(dall_a, dt_alla), cast(dall_b, dt_allb),
                dscale.reshape(shape_s).to(dt_s), None, None, None, None)


def contrastive_loss_with_temperature(
    embeddings_a: Tensor,
    embeddings_b: Tensor,
    logit_scale: Union[nn.Parameter, Tensor],
    mask: Optional[Tensor] = None,
    backprop_type: BackpropType = BackpropType.GLOBAL,
    cross_entropy_kwargs: Optional[Dict[str, Any]] = None,
    return_logits: bool = True,
) -> ContrastiveLossOutput:
    """`cross_entropy_kwargs` (losses.py:579-583): `label_smoothing` is computed by the fused kernels; the other
    F.cross_entropy options are accepted at their default values only.  `return_logits=False` (extension) skips
    writing the two [B, N] logit matrices — the fused kernels then never materialise them (fields are None)."""
    smoothing = 0.0
    for k, v in (cross_entropy_kwargs or {}).items():
        if k == "label_smoothing":
            smoothing = float(v)
        elif (k, v) not in (("reduction", "mean"), ("ignore_index", -100), ("weight", None), ("size_average", None),
                            ("reduce", None)):
            raise NotImplementedError(f"cross_entropy_kwargs[{k!r}]={v!r} is not supported by the fused InfoNCE kernels")
    if not embeddings_a.is_cuda:
        raise RuntimeError("medmoe_b200 losses run on CUDA tensors only; there is no CPU fallback")
    B = embeddings_a.shape[0]
    if is_distributed():
        if embeddings_a.shape == embeddings_b.shape and embeddings_a.dtype == embeddings_b.dtype:
            # both embeddings travel in ONE all-gather (and one reduce-scatter in backward): the collectives of this loss are
            # latency-bound (2 x 0.8 MB), so halving their number is what counts
            D = embeddings_a.shape[1]
            both = all_gather_cat(torch.cat([embeddings_a, embeddings_b], dim=1), backprop_type)
            all_a, all_b = both[:, :D], both[:, D:]
        else:
            all_a = all_gather_cat(embeddings_a, backprop_type)
            all_b = all_gather_cat(embeddings_b, backprop_type)
        label0 = B * get_rank()                                   # losses.py:515-518
    else:
        all_a, all_b, label0 = embeddings_a, embeddings_b, 0
    row_w = None
    if mask is not None:
        m = mask.to(device=embeddings_a.device, dtype=torch.float32)
        row_w = (m / m.sum()).contiguous()                        # mean over the selected rows (losses.py:574-577)
    with torch.cuda.device(embeddings_a.device):
        loss_a, loss_b, logits_a, logits_b = _InfoNCEFunction.apply(embeddings_a, embeddings_b, all_a, all_b, logit_scale,
                                                                    label0, row_w, smoothing, return_logits)
    if mask is not None and logits_a is not None:
        logits_a, logits_b = logits_a[mask], logits_b[mask]
    return ContrastiveLossOutput(loss=(loss_a + loss_b) / 2, logits_a=logits_a, logits_b=logits_b, loss_a=loss_a,
                                 loss_b=loss_b)


class _L2NormalizeFunction(torch.autograd.Function):
    """F.normalize(x, dim=-1) (eps 1e-12) on [R, D]."""

    @staticmethod
    def forward(ctx, x):
        x32 = x.float().contiguous()
        y, n = ops.l2_normalize_fwd(x32)
        ctx.save_for_backward(y, n)
        ctx.dt = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        y, n = ctx.saved_tensors
        return ops.l2_normalize_bwd(dy.float().contiguous(), y, n).to(ctx.dt)


class FLAVAGlobalContrastiveLoss(nn.Module):
    def __init__(
        self,
        logit_scale: Union[float, nn.Parameter] = None,
        image_embedding_size: int = 768,
        text_embedding_size: int = 768,
        projection_size: int = 768,
        image_embedding_index: int = 0,
        text_embedding_index: int = 0,
    ):
        super().__init__()
        if logit_scale is None:
            logit_scale = DEFAULT_LOGIT_SCALE
        if isinstance(logit_scale, nn.Parameter):
            self.logit_scale = logit_scale
        else:
            self.logit_scale = nn.Parameter(logit_scale * torch.ones([]))
        # extension: set to False when nothing reads image_logits / text_logits (training): the fused kernels then keep the
        # logits on chip (the output fields are None)
        self.return_logits = True

    def forward(self, image_sequence: Tensor, text_sequence: Tensor, mask: Optional[Tensor] = None) -> FLAVAGlobalContrastiveLossOutput:
        text_embedding = _L2NormalizeFunction.apply(text_sequence)
        image_embedding = _L2NormalizeFunction.apply(image_sequence)
        self.logit_scale.data.clamp_(0, 4.6052)                  # losses.py:281
        out = contrastive_loss_with_temperature(
            embeddings_a=image_embedding, embeddings_b=text_embedding, logit_scale=self.logit_scale, mask=mask,
            backprop_type=BackpropType.GLOBAL, return_logits=self.return_logits)
        return FLAVAGlobalContrastiveLossOutput(
            loss=out.loss, image_logits=out.logits_a, text_logits=out.logits_b, image_loss=out.loss_a,
            text_loss=out.loss_b, text_embedding=text_embedding, image_embedding=image_embedding,
            logit_scale=self.logit_scale.data)


def zero_shot_predict(image_embeddings: Tensor, text_embeddings: Tensor, return_similarity: bool = False):
    """pred[m] = argmax_c cos(image_m, text_c) (first maximum wins); fp32 kernel, exact argmax."""
    if not image_embeddings.is_cuda:
        raise RuntimeError("medmoe_b200 runs on CUDA tensors only; there is no CPU fallback")
    return ops.zeroshot_argmax(image_embeddings.float().contiguous(), text_embeddings.float().contiguous(),
                               return_sim=return_similarity)
