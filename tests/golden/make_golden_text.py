"""Golden vectors of `BertEncoder.aggregate_tokens` from the UNMODIFIED reference method
(src/models/components/text_encoder.py:32-90).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_text.py        -> tests/golden/text_aggregate.npz

The method is called unbound on a stub that only carries `idxtoword` (the method reads nothing else), so neither the
HuggingFace weights nor the tokenizer are needed.  Captions: [CLS] pieces... [SEP] pad..., word pieces with and
without "##" continuation, one caption per length class incl. a single-word caption and a full-length one.
"""
import importlib.util
import os
import random
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
VOCAB = ["[PAD]", "[CLS]", "[SEP]", "heart", "##s", "##ize", "normal", "lung", "##s", "clear", "no", "effusion", "##al", "pleur",
         "cardio", "##megaly", "mild", "##ly", "enlarged", "opacity"]


def main():
    spec = importlib.util.spec_from_file_location("ref_text_encoder", "/root/reference/src/models/components/text_encoder.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    idx = dict(enumerate(VOCAB))
    stub = types.SimpleNamespace(idxtoword=idx)
    B, n_layers, L, D = 12, 4, 25, 16          # L = the configured text.max_length (configs/model/med-moe.yaml:40)
    rng = random.Random(2024)
    ids = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        n = [1, L - 2][b] if b < 2 else rng.randint(2, L - 2)
        row = [1] + [rng.randrange(3, len(VOCAB)) for _ in range(n)] + [2]
        ids[b, :len(row)] = torch.tensor(row)
    emb = torch.randn(B, n_layers, L, D, generator=torch.Generator().manual_seed(7))
    ref, sents = mod.BertEncoder.aggregate_tokens(stub, emb, ids)
    n_words = np.array([len([w for w in s if w != "[PAD]"]) for s in sents])
    np.savez_compressed(os.path.join(HERE, "text_aggregate.npz"), embeddings=emb.numpy(), caption_ids=ids.numpy(),
                        aggregated=ref.numpy(), n_words=n_words, vocab=np.array(VOCAB), sep_id=np.array(2))
    print("wrote text_aggregate.npz", ref.shape, n_words.tolist())


if __name__ == "__main__":
    main()
