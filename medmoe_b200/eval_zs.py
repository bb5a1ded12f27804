"""Zero-shot classification driver around the `mm_zeroshot_argmax` kernel.

The reference ships no code for it (`src/eval_zs.py` is empty; `configs/eval_zs.yaml:16` names 5 classes); the semantics
are the paper's: one text embedding per class (the mean of its L2-normalised prompt embeddings when several prompts are
given), prediction = argmax of the cosine similarity with the image's global embedding (MoE `global_feat`).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Union

import torch
from torch import Tensor

from .losses import zero_shot_predict


def class_embeddings(prompt_embeddings: Tensor) -> Tensor:
    """[C, D] stays as is; [C, n_prompts, D] -> mean of the unit-norm prompt embeddings of each class."""
    if prompt_embeddings.dim() == 2:
        return prompt_embeddings
    if prompt_embeddings.dim() != 3:
        raise ValueError("class prompt embeddings must be [C, D] or [C, n_prompts, D]")
    unit = torch.nn.functional.normalize(prompt_embeddings.float(), dim=-1)
    return unit.mean(dim=1)


@torch.no_grad()
def zero_shot_evaluate(image_embeddings: Union[Tensor, Iterable[Tensor]], prompt_embeddings: Tensor,
                       labels: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """image_embeddings: [M, D] tensor or an iterable of [m_i, D] batches (CUDA); labels: int [M] (optional).

    Returns {"pred": int64 [M]} and, with labels, "accuracy" (0-dim) and "per_class_accuracy" [C] (NaN for absent classes)."""
    txt = class_embeddings(prompt_embeddings).cuda()
    batches = [image_embeddings] if isinstance(image_embeddings, Tensor) else list(image_embeddings)
    pred = torch.cat([zero_shot_predict(b, txt) for b in batches]) if batches else torch.empty(0, dtype=torch.long, device=txt.device)
    out = {"pred": pred}
    if labels is not None:
        labels = labels.to(pred.device).long()
        if labels.numel() != pred.numel():
            raise ValueError(f"{labels.numel()} labels for {pred.numel()} images")
        hit = (pred == labels).float()
        C = txt.shape[0]
        n = torch.zeros(C, device=pred.device).index_add_(0, labels, torch.ones_like(hit))
        ok = torch.zeros(C, device=pred.device).index_add_(0, labels, hit)
        out["accuracy"] = hit.mean() if hit.numel() else torch.tensor(float("nan"), device=pred.device)
        out["per_class_accuracy"] = ok / n          # 0/0 -> NaN for classes without samples
    return out
