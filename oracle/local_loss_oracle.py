"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's word-patch attention loss:

* `attention`          — attention_fn, reference src/losses.py:698-736
* `gloria_local_loss`  — GLORIALocalContrastiveLoss.forward, src/losses.py:954-1026 (per-caption Python loop with
                         `.repeat`), restated per (image, caption) pair in one batched computation.

Parity pinning: tests/golden/local_loss.npz is produced by the UNMODIFIED reference class
(tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against it, and
tests/test_oracle_vs_reference.py against the live class when /root/reference is present.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F


def attention(query: torch.Tensor, context: torch.Tensor, temp1: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """query [B, D, W], context [B, D, P] -> (weighted context [B, D, W], attention [B, W, P])   (losses.py:698-736)."""
    attn = torch.bmm(context.transpose(1, 2), query)            # [B, P, W]
    attn = torch.softmax(attn, dim=-1)                          # over the words        (:717)
    attn = attn.transpose(1, 2)                                 # [B, W, P]
    attn = torch.softmax(attn * temp1, dim=-1)                  # over the patches      (:725-726)
    weighted = torch.bmm(context, attn.transpose(1, 2))         # [B, D, W]             (:733)
    return weighted, attn


def similarities(img_features: torch.Tensor, words_emb: torch.Tensor, cap_lens: Sequence[int], temp1: float = 4.0,
                 temp2: float = 5.0, agg: str = "sum", eps: float = 1e-8) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """sim[b, i] (before temp3) and the attention maps of the matching pairs.  img_features [B, D, H, W], words_emb [B, D, L]."""
    B, D = img_features.shape[:2]
    ih, iw = img_features.shape[2:]
    context = img_features.reshape(B, D, -1)
    cols, att_maps = [], []
    for i in range(words_emb.shape[0]):
        n = int(cap_lens[i])
        word = words_emb[i, :, :n].unsqueeze(0).expand(B, D, n)                 # :979-980
        wc, attn = attention(word, context, temp1)
        att_maps.append(attn[i].reshape(1, n, ih, iw))                          # :987-989
        w12 = (word * wc).sum(1)
        den = (word.norm(dim=1) * wc.norm(dim=1)).clamp(min=eps)                # cosine_similarity, :690-696
        row = torch.exp(temp2 * (w12 / den))                                    # :998
        row = row.sum(1, keepdim=True) if agg == "sum" else row.mean(1, keepdim=True)
        cols.append(torch.log(row))                                             # :1003
    return torch.cat(cols, 1), att_maps


def gloria_local_loss(img_features: torch.Tensor, words_emb: torch.Tensor, cap_lens: Sequence[int], temp1: float = 4.0,
                      temp2: float = 5.0, temp3: float = 10.0, agg: str = "sum"):
    """-> (loss0, loss1, att_maps)   (losses.py:1007-1021)."""
    sim, att_maps = similarities(img_features, words_emb, cap_lens, temp1, temp2, agg)
    sim = sim * temp3
    labels = torch.arange(img_features.shape[0], device=sim.device)
    return F.cross_entropy(sim, labels), F.cross_entropy(sim.t(), labels), att_maps
