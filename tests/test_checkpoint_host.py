"""Host logic: finding the MoE block inside reference checkpoints (SURVEY §8b state_dict keys, §8f row 4)."""
import os

import pytest
import torch

import medmoe_b200
from medmoe_b200 import checkpoint


def _moe(K=3):
    torch.manual_seed(0)
    return medmoe_b200.MoE(num_experts=K)


@pytest.mark.parametrize("prefix", ["model.image_encoder.model.moe.", "model.moe.", "moe.", ""])
def test_extract_and_load_by_prefix(prefix, tmp_path):
    src, dst = _moe(), _moe()
    with torch.no_grad():
        for p in src.parameters():
            p.add_(1.0)
    sd = {prefix + k: v.clone() for k, v in src.state_dict().items()}
    sd["model.text_encoder.model.embeddings.word_embeddings.weight"] = torch.zeros(4, 4)       # unrelated towers
    sd["model.image_encoder.model.model.encoder.layers.0.blocks.0.attention.self.query.weight"] = torch.zeros(2, 2)
    ckpt = {"state_dict": sd, "epoch": 3}
    path = tmp_path / "ref.ckpt"
    torch.save(ckpt, path)
    res = checkpoint.load_reference_checkpoint(dst, str(path))
    assert not res.missing_keys and not res.unexpected_keys
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k


def test_expert_count_mismatch_raises():
    sd = {"moe." + k: v for k, v in _moe(4).state_dict().items()}
    with pytest.raises(RuntimeError, match="4 experts"):
        checkpoint.load_reference_checkpoint(_moe(3), sd)
    with pytest.raises(KeyError):
        checkpoint.extract_moe_state_dict({"foo.weight": torch.zeros(1)})


def test_medclip_rename_follows_reference():
    sd = {"vision_model.encoder.w": torch.zeros(1), "text_model.x": torch.zeros(1)}
    assert list(checkpoint.rename_medclip_vision_keys(sd)) == ["model.encoder.w"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")
def test_keys_equal_the_reference_modules():
    from oracle import reference_shim
    ref = reference_shim.load_moe_module().MoE(num_experts=3)
    sd = {"model.image_encoder.model.moe." + k: v for k, v in ref.state_dict().items()}
    dst = _moe(3)
    res = checkpoint.load_reference_checkpoint(dst, sd)
    assert not res.missing_keys and not res.unexpected_keys
    assert set(dst.state_dict()) == set(ref.state_dict())
