"""CPU, build container only (needs /root/reference): the oracle and the drop-in boundary against the
reference's own unmodified classes, live."""
import pytest
import torch

import medmoe_b200
from oracle import loss_oracle as lo
from oracle import moe_oracle as mo
from oracle import reference_shim as rs

pytestmark = pytest.mark.reference


def test_state_dict_keys_shapes_and_default_init_match_reference():
    swin = rs.load_moe_module()
    torch.manual_seed(7)
    ref = swin.MoE()
    torch.manual_seed(7)
    mine = medmoe_b200.MoE()
    sd_r, sd_m = ref.state_dict(), mine.state_dict()
    assert list(sd_r.keys()) == list(sd_m.keys())
    for k in sd_r:
        assert sd_r[k].shape == sd_m[k].shape and torch.equal(sd_r[k], sd_m[k]), k   # same RNG stream => same weights
    mine.load_state_dict(sd_r)                       # checkpoints load
    assert sum(p.numel() for p in mine.parameters()) == 8_527_244


def test_survey_anchors_and_oracle_on_reference_shapes():
    """SURVEY §8c sanity anchors, re-derived: seed 0 weights, seed 1 inputs, B = 4, 224^2 token counts."""
    swin = rs.load_moe_module()
    torch.manual_seed(0)
    moe = swin.MoE()
    torch.manual_seed(1)
    feats = [torch.randn(4, p, d) for p, d in zip([3136, 784, 196, 49], [96, 192, 384, 768])]
    sw = torch.randn(4, 768)
    g, l, probs = moe(feats, sw)
    assert torch.argmax(probs, -1).tolist() == [2, 1, 5, 1]
    assert abs(g.sum().item() - 707.869141) < 1e-2
    params = {k: v.detach() for k, v in moe.state_dict().items()}
    (g2, l2, p2), idx = mo.moe_forward_sparse(params, feats, sw)
    assert idx[:, 0].tolist() == [2, 1, 5, 1] and torch.equal(p2, probs)
    assert (g2 - g).abs().max().item() < 1e-5 and (l2 - l).abs().max().item() < 1e-5


def test_activate_swaps_the_reference_symbols():
    swin = rs.load_moe_module()
    losses = rs.load_losses_module()
    orig = (swin.MoE, swin.Expert, losses.GLORIAGlobalContrastiveLoss, losses.FLAVAGlobalContrastiveLoss,
            losses.contrastive_loss_with_temperature)
    try:
        medmoe_b200.activate()
        assert swin.MoE is medmoe_b200.MoE and losses.GLORIAGlobalContrastiveLoss is medmoe_b200.GLORIAGlobalContrastiveLoss
        assert losses.contrastive_loss_with_temperature is medmoe_b200.contrastive_loss_with_temperature
    finally:
        (swin.MoE, swin.Expert, losses.GLORIAGlobalContrastiveLoss, losses.FLAVAGlobalContrastiveLoss,
         losses.contrastive_loss_with_temperature) = orig


def test_loss_oracle_live():
    L = rs.load_losses_module()
    torch.manual_seed(3)
    I, T = torch.randn(20, 768), torch.randn(20, 768)
    assert torch.equal(L.GLORIAGlobalContrastiveLoss()(I, T, temp3=7.0), lo.gloria_global_loss(I, T, 7.0))
    out = L.FLAVAGlobalContrastiveLoss()(I, T)
    mine = lo.flava_global_loss(I, T, torch.tensor(lo.DEFAULT_LOGIT_SCALE))
    assert torch.equal(out.loss, mine[0]) and torch.equal(out.image_logits, mine[1])
