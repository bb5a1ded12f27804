"""Word-piece aggregation (text_encoder.py:32-90) as a segmented sum: equal to the reference's Python loop."""
import os
import random

import pytest
import torch

from medmoe_b200 import text as mmtext

VOCAB = ["[PAD]", "[CLS]", "[SEP]", "heart", "##s", "##ize", "normal", "lung", "##s", "clear", "no", "effusion", "##al", "pleur"]
IDX = dict(enumerate(VOCAB))
CLS, SEP = 1, 2


def _loop_reference(embeddings, caption_ids):
    """Plain restatement of the reference loop (used when the reference checkout is absent)."""
    B, n_layers, L, D = embeddings.shape
    emb = embeddings.permute(0, 2, 1, 3)
    out = torch.zeros(B, L, n_layers, D)
    for b in range(B):
        agg, bank, nbank = [], [], 0
        for t in range(L):
            w = IDX[int(caption_ids[b, t])]
            if w == "[SEP]":
                agg.append(torch.stack(bank).sum(0))
                agg.append(emb[b, t])
                break
            if not w.startswith("##"):
                if nbank == 0:
                    bank.append(emb[b, t]); nbank += 1
                else:
                    agg.append(torch.stack(bank).sum(0))
                    bank, nbank = [emb[b, t]], 1
            else:
                bank.append(emb[b, t]); nbank += 1
        for i, e in enumerate(agg):
            out[b, i] = e
    return out.permute(0, 2, 1, 3)


def _captions(B, L, seed, with_sep=True):
    rng = random.Random(seed)
    ids = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        n = rng.randint(1, L - 2)
        row = [CLS] + [rng.randrange(3, len(VOCAB)) for _ in range(n)]
        if with_sep or b % 2 == 0:
            row.append(SEP)
        ids[b, :len(row)] = torch.tensor(row)
    return ids


@pytest.mark.parametrize("seed,with_sep", [(0, True), (1, True), (2, False)])
def test_aggregate_tokens_equals_the_loop(seed, with_sep):
    B, n_layers, L, D = 5, 4, 12, 8
    ids = _captions(B, L, seed, with_sep)
    if not with_sep:                       # rows without [SEP] must not contain pad ids that look like words
        ids[ids == 0] = 3
    emb = torch.randn(B, n_layers, L, D, generator=torch.Generator().manual_seed(seed))
    table = mmtext.continuation_table(IDX)
    got, n_words = mmtext.aggregate_tokens(emb, ids, table, SEP)
    ref = _loop_reference(emb, ids)
    assert torch.allclose(got, ref, atol=1e-6)
    for b in range(B):
        assert got[b, :, int(n_words[b]):].abs().sum() == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")
def test_aggregate_tokens_equals_the_reference_method():
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("ref_text_encoder", "/root/reference/src/models/components/text_encoder.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    stub = types.SimpleNamespace(idxtoword=IDX)
    B, n_layers, L, D = 4, 4, 10, 6
    ids = _captions(B, L, 7)
    emb = torch.randn(B, n_layers, L, D, generator=torch.Generator().manual_seed(7))
    ref, sents = mod.BertEncoder.aggregate_tokens(stub, emb, ids)
    got, _ = mmtext.aggregate_tokens(emb, ids, mmtext.continuation_table(IDX), SEP)
    assert torch.allclose(got, ref, atol=1e-6)
    assert mmtext.sentences_from_ids(ids, IDX) == sents


def test_aggregate_tokens_equals_the_committed_reference_golden():
    """tests/golden/text_aggregate.npz was produced by the reference method (make_golden_text.py); this runs everywhere."""
    import numpy as np
    from tests.util import GOLDEN
    z = np.load(f"{GOLDEN}/text_aggregate.npz")
    vocab = [str(w) for w in z["vocab"]]
    got, n_words = mmtext.aggregate_tokens(torch.from_numpy(z["embeddings"]), torch.from_numpy(z["caption_ids"]),
                                           mmtext.continuation_table(dict(enumerate(vocab))), int(z["sep_id"]))
    assert torch.allclose(got, torch.from_numpy(z["aggregated"]), atol=1e-6)
    assert n_words.tolist() == z["n_words"].tolist()
