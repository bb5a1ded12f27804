"""Text-tower glue without per-token host round trips.

`BertEncoder.aggregate_tokens` (src/models/components/text_encoder.py:32-90) merges word pieces into words with a Python
double loop that calls `.item()` on every token id (B x L device syncs per step).  The same result is a segmented sum:
a token starts a new word unless its piece begins with "##", everything after the first [SEP] is dropped, and the word
embeddings are the sums of their pieces, left-aligned and zero-padded to the original length.  Quirks kept on purpose:
[CLS] and [SEP] are words of their own; a caption without [SEP] loses its last word (the loop only flushes on [SEP]).
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Sequence, Tuple

import torch
from torch import Tensor


def continuation_table(idxtoword: Mapping[int, str], vocab_size: int = 0) -> Tensor:
    """bool [vocab]: True where the word piece continues the previous one (starts with "##")."""
    n = max(vocab_size, max(idxtoword) + 1 if idxtoword else 0)
    table = torch.zeros(n, dtype=torch.bool)
    cont = [i for i, w in idxtoword.items() if w.startswith("##")]
    if cont:
        table[torch.tensor(cont)] = True
    return table


def aggregate_tokens(embeddings: Tensor, caption_ids: Tensor, is_continuation: Tensor, sep_id: int) -> Tuple[Tensor, Tensor]:
    """embeddings [B, layers, L, D], caption_ids [B, L] -> (word embeddings [B, layers, L, D], number of words [B]).

    Runs wherever its inputs live; no host synchronisation."""
    B, n_layers, L, D = embeddings.shape
    ids = caption_ids.to(embeddings.device)
    cont = is_continuation.to(embeddings.device)[ids]                       # [B, L]
    starts = ~cont
    starts[:, 0] = True                                                     # a leading "##" piece still opens the first word
    is_sep = ids == sep_id
    seen_sep = torch.cumsum(is_sep.int(), dim=1)
    keep = (seen_sep - is_sep.int()) == 0                                   # tokens up to and including the first [SEP]
    word = torch.cumsum(starts.int(), dim=1) - 1                            # word index of every token
    has_sep = is_sep.any(dim=1)
    n_words = (word * keep).amax(dim=1) + 1
    n_words = torch.where(has_sep, n_words, n_words - 1)                    # no [SEP]: the last word is never flushed
    keep = keep & (word < n_words.unsqueeze(1))
    # pieces that are dropped go to a scratch slot L that is cut off afterwards
    slot = torch.where(keep, word, torch.full_like(word, L)).long()
    out = embeddings.new_zeros(B, n_layers, L + 1, D)
    out.scatter_add_(2, slot.view(B, 1, L, 1).expand(B, n_layers, L, D), embeddings)
    return out[:, :, :L, :], n_words


def sentences_from_ids(caption_ids: Tensor, idxtoword: Mapping[int, str]) -> List[List[str]]:
    """The word strings the reference returns next to the embeddings (host side, for logging / attention maps only)."""
    sents = []
    for row in caption_ids.tolist():
        words: List[str] = []
        bank: List[str] = []
        for tok in row:
            w = idxtoword[tok]
            if w == "[SEP]":
                words.append("".join(bank))
                words.append(w)
                break
            if w.startswith("##"):
                bank.append(w[2:])
            elif not bank:
                bank.append(w)
            else:
                words.append("".join(bank))
                bank = [w]
        sents.append(words + ["[PAD]"] * (len(row) - len(words)))
    return sents
