"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's global contrastive objectives:

* `gloria_global_loss`      — GLORIAGlobalContrastiveLoss.forward, reference src/losses.py:766-794
                              (the configured loss: configs/model/med-moe_pretraining.yaml:29-31).
* `contrastive_loss_with_temperature` / `flava_global_loss`
                            — src/losses.py:527-592 and :268-301, with the all-gather of
                              :503-524 restated as "rank r owns rows [r*B, (r+1)*B) of the
                              concatenated problem" (the preferred multi-rank oracle).
* `zero_shot_predict`       — SURVEY §8a row Z (src/eval_zs.py is empty in the reference):
                              fp64 cosine similarity + first-max argmax.

Parity pinning: see oracle/moe_oracle.py — pinned by tests/golden/*.npz generated from the
reference's own classes, and live against /root/reference when it is present.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

DEFAULT_LOGIT_SCALE = math.log(1 / 0.07)   # losses.py:595
LOGIT_SCALE_MAX = 4.6052                   # losses.py:281


def gloria_global_loss(cnn_code: torch.Tensor, rnn_code: torch.Tensor, temp3: float = 10.0, eps: float = 1e-8) -> torch.Tensor:
    """loss = CE(S, arange) + CE(S^T, arange), S = temp3 * (I T^T) / max(|I| |T|^T, eps)  (losses.py:782-794)."""
    n_i = cnn_code.norm(dim=-1, keepdim=True)
    n_t = rnn_code.norm(dim=-1, keepdim=True)
    scores = (cnn_code @ rnn_code.t()) / (n_i @ n_t.t()).clamp(min=eps) * temp3
    labels = torch.arange(cnn_code.shape[0])
    return F.cross_entropy(scores, labels) + F.cross_entropy(scores.t(), labels)


def contrastive_loss_with_temperature(emb_a: torch.Tensor, emb_b: torch.Tensor, logit_scale: torch.Tensor,
                                      all_a: Optional[torch.Tensor] = None, all_b: Optional[torch.Tensor] = None,
                                      rank: int = 0, mask: Optional[torch.Tensor] = None, label_smoothing: float = 0.0):
    """One rank's view of losses.py:527-592.  all_a / all_b are the concatenated embeddings of
    every rank (None = single process); labels are local_batch * rank + arange (losses.py:516);
    `label_smoothing` is the one `cross_entropy_kwargs` entry restated (losses.py:579-583)."""
    temperature = torch.exp(logit_scale)
    if all_a is None:
        all_a, all_b = emb_a, emb_b
    B = emb_a.shape[0]
    labels = B * rank + torch.arange(B)
    logits_a = emb_a @ all_b.t() * temperature
    logits_b = emb_b @ all_a.t() * temperature
    if mask is not None:
        logits_a, logits_b, labels = logits_a[mask], logits_b[mask], labels[mask]
    loss_a = F.cross_entropy(logits_a, labels, label_smoothing=label_smoothing)
    loss_b = F.cross_entropy(logits_b, labels, label_smoothing=label_smoothing)
    return (loss_a + loss_b) / 2, logits_a, logits_b, loss_a, loss_b


def flava_global_loss(image_seq: torch.Tensor, text_seq: torch.Tensor, logit_scale: torch.Tensor,
                      mask: Optional[torch.Tensor] = None):
    """FLAVAGlobalContrastiveLoss.forward single-process (losses.py:268-301): normalise, clamp the
    scale to [0, 4.6052], then the function above."""
    t = F.normalize(text_seq, dim=-1)
    i = F.normalize(image_seq, dim=-1)
    with torch.no_grad():
        logit_scale.clamp_(0, LOGIT_SCALE_MAX)      # in-place on .data, no gradient effect (losses.py:281)
    return contrastive_loss_with_temperature(i, t, logit_scale, mask=mask)


def flava_multi_rank(a_parts: List[torch.Tensor], b_parts: List[torch.Tensor], logit_scale: torch.Tensor):
    """World-size W emulation on concatenated embeddings: returns per-rank losses and their mean
    (what DDP's gradient averaging optimises).  Gradients w.r.t. a_parts/b_parts obtained from
    the mean are the DDP-averaged gradients a real run produces (all_gather_with_backprop's
    backward sums the contributions of every rank's loss, src/utils/distributed.py:47-48)."""
    all_a, all_b = torch.cat(a_parts), torch.cat(b_parts)
    losses = [contrastive_loss_with_temperature(a, b, logit_scale, all_a, all_b, rank=r)[0]
              for r, (a, b) in enumerate(zip(a_parts, b_parts))]
    return losses, torch.stack(losses).mean()


def zero_shot_predict(img: torch.Tensor, txt: torch.Tensor, eps: float = 1e-8) -> Tuple[torch.Tensor, torch.Tensor]:
    """pred[m] = argmax_c cos(img_m, txt_c) in fp64; returns (pred int64 [M], sim fp64 [M, C])."""
    i, t = img.double(), txt.double()
    sim = (i @ t.t()) / (i.norm(dim=-1, keepdim=True) @ t.norm(dim=-1, keepdim=True).t()).clamp(min=eps)
    return torch.argmax(sim, dim=-1), sim
