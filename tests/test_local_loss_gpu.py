"""GPU parity of medmoe_b200.GLORIALocalContrastiveLoss through the C-ABI kernels.

Floating-point path with bf16 GEMM operands and fp32 accumulation; tolerances are norm-wise relative errors against the
fp32 reference (golden vectors of the reference class) / the fp32 oracle: 2e-2 for gradients, 5e-3 for the losses.
"""
import os

import numpy as np
import pytest
import torch

import medmoe_b200
from medmoe_b200.local_loss import GLORIALocalContrastiveLoss, local_similarities
from oracle import local_loss_oracle as lo
from tests.util import rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "local_loss.npz")


@pytest.mark.parametrize("agg", ["sum", "mean"])
def test_local_loss_matches_reference_golden(agg):
    g = np.load(GOLDEN)
    img = torch.tensor(g["img"]).cuda().requires_grad_(True)
    words = torch.tensor(g["words"]).cuda().requires_grad_(True)
    cap_lens = g["cap_lens"].tolist()
    out = GLORIALocalContrastiveLoss()(img, words, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg=agg)
    (out.loss0 + out.loss1).backward()
    assert abs(out.loss0.item() - float(g[f"loss0_{agg}"])) < 5e-3 * max(1.0, abs(float(g[f"loss0_{agg}"])))
    assert abs(out.loss1.item() - float(g[f"loss1_{agg}"])) < 5e-3 * max(1.0, abs(float(g[f"loss1_{agg}"])))
    assert rel_err(img.grad.cpu(), torch.tensor(g[f"d_img_{agg}"])) < 2e-2
    assert rel_err(words.grad.cpu(), torch.tensor(g[f"d_words_{agg}"])) < 2e-2
    assert len(out.att_maps) == len(cap_lens)
    assert out.att_maps[3].shape == (1, cap_lens[3], 6, 6)
    assert torch.allclose(out.att_maps[0].cpu(), torch.tensor(g[f"att0_{agg}"]), atol=1e-4)
    assert torch.allclose(out.att_maps[3].cpu(), torch.tensor(g[f"att3_{agg}"]), atol=1e-4)


@pytest.mark.parametrize("B,D,H,L,blocks", [(12, 768, 14, 20, 1), (37, 256, 7, 11, 3), (3, 768, 56, 25, 1)])
def test_local_loss_vs_oracle(B, D, H, L, blocks, monkeypatch):
    """Ragged captions, P not a multiple of 128, several caption blocks (forced by a small score budget)."""
    from medmoe_b200 import local_loss as ll
    g = torch.Generator().manual_seed(B)
    img = (torch.randn(B, D, H, H, generator=g) * 0.3).to(torch.bfloat16).float()
    words = (torch.randn(B, D, L, generator=g) * 0.3).to(torch.bfloat16).float()
    cap_lens = [int(x) for x in torch.randint(1, L + 1, (B,), generator=g)]
    cap_lens[0] = L
    if blocks > 1:
        rows = B * ((H * H + 127) // 128 * 128)
        monkeypatch.setattr(ll, "SCORE_BYTES_BUDGET", rows * 8 * 4 * 16)      # 16 captions of 8 words per block
    ri, rw = img.clone().requires_grad_(True), words.clone().requires_grad_(True)
    l0, l1, _ = lo.gloria_local_loss(ri, rw, cap_lens)
    (l0 + 2.0 * l1).backward()
    di, dw = img.cuda().requires_grad_(True), words.cuda().requires_grad_(True)
    out = GLORIALocalContrastiveLoss(return_att_maps=False)(di, dw, cap_lens)
    (out.loss0 + 2.0 * out.loss1).backward()
    assert abs(out.loss0.item() - l0.item()) < 5e-3 * max(1.0, abs(l0.item()))
    assert abs(out.loss1.item() - l1.item()) < 5e-3 * max(1.0, abs(l1.item()))
    assert rel_err(di.grad.cpu(), ri.grad) < 2e-2
    assert rel_err(dw.grad.cpu(), rw.grad) < 2e-2
    # words beyond a caption's length get exactly zero gradient
    for i, n in enumerate(cap_lens):
        assert dw.grad[i, :, n:].abs().sum().item() == 0.0


def test_local_similarities_take_the_moe_local_feat_view():
    """local_feat of medmoe_b200.MoE is a [B, D, H, W] stride view of the token-major buffer; the loss consumes it as is (bf16)."""
    B, D, H, L = 4, 768, 8, 6
    g = torch.Generator(device="cuda").manual_seed(0)
    fused = (torch.randn(B, H * H, D, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    local = fused.transpose(1, 2).reshape(B, D, H, H)
    words = torch.randn(B, D, L, device="cuda", generator=g) * 0.3
    sim = local_similarities(local, words, [L] * B)
    ref, _ = lo.similarities(local.float().cpu(), words.to(torch.bfloat16).float().cpu(), [L] * B)
    assert rel_err(sim.cpu(), ref) < 5e-3


def test_fused_score_softmax_epilogue_agrees_with_the_two_pass_path(monkeypatch):
    """Captions of <= 32 words: the first softmax runs in the score GEMM's epilogue; same result as GEMM + separate pass."""
    from medmoe_b200 import local_loss as ll
    B, D, H, L = 9, 768, 12, 30
    g = torch.Generator(device="cuda").manual_seed(2)
    img = torch.randn(B, D, H, H, device="cuda", generator=g) * 0.3
    words = torch.randn(B, D, L, device="cuda", generator=g) * 0.3
    cap_lens = [30, 1, 17, 25, 32, 8, 30, 29, 2]
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(ll, "FUSE_SCORE_SOFTMAX", fused)
        x, w = img.clone().requires_grad_(True), words.clone().requires_grad_(True)
        sim = local_similarities(x, w, cap_lens)
        sim.square().sum().backward()
        outs.append((sim.detach(), x.grad, w.grad))
    assert rel_err(outs[0][0], outs[1][0]) < 1e-4
    assert rel_err(outs[0][1], outs[1][1]) < 2e-3
    assert rel_err(outs[0][2], outs[1][2]) < 2e-3


@pytest.mark.parametrize("B,D,H,L,lens", [(1, 128, 3, 4, [4]), (2, 256, 7, 40, [1, 40]), (5, 768, 4, 128, [128, 3, 64, 65, 96])])
def test_local_loss_edge_shapes_vs_oracle(B, D, H, L, lens):
    """A single pair, one-word captions, the longest supported captions (128 words: four lanes groups), bf16 inputs."""
    g = torch.Generator().manual_seed(100 + B)
    img = (torch.randn(B, D, H, H, generator=g) * 0.3).to(torch.bfloat16)
    words = (torch.randn(B, D, L, generator=g) * 0.3).to(torch.bfloat16)
    ri, rw = img.float().requires_grad_(True), words.float().requires_grad_(True)
    l0, l1, _ = lo.gloria_local_loss(ri, rw, lens, agg="mean")
    (l0 + l1).backward()
    di, dw = img.cuda().requires_grad_(True), words.cuda().requires_grad_(True)
    out = GLORIALocalContrastiveLoss()(di, dw, lens, agg="mean")
    (out.loss0 + out.loss1).backward()
    assert di.grad.dtype == torch.bfloat16 and dw.grad.dtype == torch.bfloat16
    assert abs(out.loss0.item() - l0.item()) < 5e-3 * max(1.0, abs(l0.item()))
    assert abs(out.loss1.item() - l1.item()) < 5e-3 * max(1.0, abs(l1.item()))
    if B > 1:      # B = 1: both cross-entropies are identically zero, and so are the gradients
        assert rel_err(di.grad.float().cpu(), ri.grad) < 3e-2
        assert rel_err(dw.grad.float().cpu(), rw.grad) < 3e-2
    assert [m.shape[1] for m in out.att_maps] == lens


def test_local_loss_rejects_captions_longer_than_128_words():
    img = torch.randn(2, 128, 4, 4, device="cuda")
    words = torch.randn(2, 128, 130, device="cuda")
    with pytest.raises(RuntimeError, match="128"):
        GLORIALocalContrastiveLoss()(img, words, [130, 5])
