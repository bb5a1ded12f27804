"""CPU: the oracle restatements against the golden vectors produced by the reference itself
(tests/golden/make_golden.py).  fp32 on both sides: bit-exact where the op sequence is the
same (dense schedule, losses), <= 1e-5 relative for the closed-form / sparse schedule."""
import pytest
import torch

from oracle import loss_oracle as lo
from oracle import moe_oracle as mo
from tests.util import golden_params, load_golden, rel_err


@pytest.mark.parametrize("case", ["moe_small", "moe_k6"])
def test_moe_dense_restatement_is_bit_exact(case):
    g = load_golden(case)
    params = golden_params(g)
    feats = [g[f"feat{s}"] for s in range(4)]
    gf, lf, probs = mo.moe_forward_dense(params, feats, g["swin_feat"])
    assert torch.equal(probs, g["probs"])
    assert torch.equal(torch.argmax(probs, -1), g["top_expert"])
    assert torch.equal(gf, g["global_feat"])
    assert torch.equal(lf.contiguous(), g["local_feat"])


@pytest.mark.parametrize("case", ["moe_small", "moe_k6"])
def test_moe_sparse_closed_form_and_gradients(case):
    g = load_golden(case)
    params = {k: v.clone().requires_grad_(True) for k, v in golden_params(g).items()}
    feats = [g[f"feat{s}"].clone().requires_grad_(True) for s in range(4)]
    sw = g["swin_feat"].clone().requires_grad_(True)
    (gf, lf, probs), idx = mo.moe_forward_sparse(params, feats, sw)
    assert idx.flatten().tolist() == g["top_expert"].tolist()
    assert rel_err(gf, g["global_feat"]) < 1e-5 and rel_err(lf, g["local_feat"]) < 1e-5
    obj = (gf * g["cot_global"]).sum() + (lf * g["cot_local"]).sum() + 2.0 * mo.router_ce(probs, g["labels"])
    obj.backward()
    for s in range(4):
        assert rel_err(feats[s].grad, g[f"d_feat{s}"]) < 1e-4
    assert rel_err(sw.grad, g["d_swin_feat"]) < 1e-5
    used = set(g["top_expert"].tolist())
    for k, p in params.items():
        gn = g["gradnorm." + k].item()
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if k.startswith("experts.") and int(k.split(".")[1]) not in used:
            assert gn == 0.0 and float(got.abs().max()) == 0.0, f"{k}: idle expert must have an all-zero gradient"
        elif k.endswith("attn_proj.2.bias"):
            # d/db2 of a softmax over scales is identically zero (shift invariance): only rounding noise
            assert float(got.abs().max()) < 1e-5 and gn < 1e-5
        else:
            assert abs(got.double().norm().item() - gn) <= 1e-4 * gn, k
            if "grad." + k in g and g["grad." + k].numel():
                assert rel_err(got, g["grad." + k]) < 1e-4, k


def test_idle_experts_have_zero_not_none_grads_in_reference():
    g = load_golden("moe_small")
    idle = [e for e in range(3) if e not in g["top_expert"].tolist()]
    assert idle, "fixture should contain an idle expert"
    for k in g:
        if k.startswith("gradnone.experts."):
            assert not bool(g[k]), "reference produces zero tensors, not None, for unselected experts (SURVEY §3.2)"
    for e in idle:
        assert g[f"gradnorm.experts.{e}.attn_proj.0.weight"].item() == 0.0


def test_gloria_global_loss_bit_exact():
    g = load_golden("losses")
    for B in (16, 37):
        I = g[f"gloria{B}.img"].clone().requires_grad_(True)
        T = g[f"gloria{B}.txt"].clone().requires_grad_(True)
        loss = lo.gloria_global_loss(I, T, 10.0)
        assert torch.equal(loss, g[f"gloria{B}.loss"])
        loss.backward()
        assert rel_err(I.grad, g[f"gloria{B}.dimg"]) < 1e-6 and rel_err(T.grad, g[f"gloria{B}.dtxt"]) < 1e-6
    assert abs(g["gloria16.loss"].item() - 5.819103717803955) < 1e-12   # SURVEY §8c anchor


def test_flava_single_process():
    g = load_golden("losses")
    I = g["flava.img"].clone().requires_grad_(True)
    T = g["flava.txt"].clone().requires_grad_(True)
    s = torch.tensor(lo.DEFAULT_LOGIT_SCALE, requires_grad=True)
    loss, la, lb, loss_a, loss_b = lo.flava_global_loss(I, T, s)
    assert torch.equal(loss, g["flava.loss"]) and torch.equal(la, g["flava.image_logits"])
    assert torch.equal(loss_a, g["flava.image_loss"]) and torch.equal(loss_b, g["flava.text_loss"])
    loss.backward()
    assert rel_err(I.grad, g["flava.dimg"]) < 1e-6 and rel_err(T.grad, g["flava.dtxt"]) < 1e-6
    assert abs(s.grad.item() - g["flava.dscale"].item()) < 1e-6
    # mask path
    I2 = g["flava.img"].clone().requires_grad_(True)
    T2 = g["flava.txt"].clone().requires_grad_(True)
    s2 = torch.tensor(lo.DEFAULT_LOGIT_SCALE, requires_grad=True)
    loss2, la2, *_ = lo.flava_global_loss(I2, T2, s2, mask=g["flava_mask.mask"])
    assert torch.equal(loss2, g["flava_mask.loss"]) and torch.equal(la2, g["flava_mask.image_logits"])


def test_flava_multi_rank_restatement():
    """rank r = rows [r*B, (r+1)*B) of the concatenated problem == the reference's gather + label offset."""
    g = load_golden("losses")
    W = 3
    a = [g[f"flava_mr.a{r}"].clone().requires_grad_(True) for r in range(W)]
    b = [g[f"flava_mr.b{r}"].clone().requires_grad_(True) for r in range(W)]
    s = g["flava_mr.scale"].clone().requires_grad_(True)
    losses, mean = lo.flava_multi_rank(a, b, s)
    assert torch.allclose(torch.stack(losses), g["flava_mr.losses"], rtol=0, atol=1e-6)
    mean.backward()
    for r in range(W):
        assert rel_err(a[r].grad, g[f"flava_mr.da{r}"]) < 1e-5 and rel_err(b[r].grad, g[f"flava_mr.db{r}"]) < 1e-5
    assert abs(s.grad.item() - g["flava_mr.dscale"].item()) < 1e-5


def test_lerp_closed_form_matches_aten():
    for p_src, p_dst in [(3136, 3136), (784, 3136), (196, 3136), (49, 3136), (2304, 9216), (144, 9216), (5, 17), (1, 64)]:
        x = torch.randn(2, p_src, 8)
        ref = torch.nn.functional.interpolate(x.transpose(1, 2), size=p_dst, mode="linear", align_corners=False).transpose(1, 2)
        assert torch.allclose(mo.lerp_rows(x, p_dst), ref, rtol=0, atol=1e-6)


def test_zero_shot_oracle():
    torch.manual_seed(0)
    img, txt = torch.randn(64, 768), torch.randn(5, 768)
    pred, sim = lo.zero_shot_predict(img, txt)
    ref = torch.nn.functional.cosine_similarity(img.double()[:, None], txt.double()[None], dim=-1)
    assert torch.allclose(sim, ref, atol=1e-12) and torch.equal(pred, ref.argmax(-1))


def test_local_loss_oracle_matches_reference_golden():
    """oracle/local_loss_oracle.py against the committed outputs of the reference's GLORIALocalContrastiveLoss."""
    import os
    import numpy as np
    from oracle import local_loss_oracle as lo
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "local_loss.npz"))
    for agg in ("sum", "mean"):
        img = torch.tensor(g["img"]).requires_grad_(True)
        words = torch.tensor(g["words"]).requires_grad_(True)
        l0, l1, att = lo.gloria_local_loss(img, words, g["cap_lens"].tolist(), agg=agg)
        (l0 + l1).backward()
        assert abs(l0.item() - float(g[f"loss0_{agg}"])) < 1e-5 and abs(l1.item() - float(g[f"loss1_{agg}"])) < 1e-5
        assert torch.allclose(img.grad, torch.tensor(g[f"d_img_{agg}"]), atol=1e-6)
        assert torch.allclose(words.grad, torch.tensor(g[f"d_words_{agg}"]), atol=1e-6)
        assert torch.allclose(att[3], torch.tensor(g[f"att3_{agg}"]), atol=1e-6)


def test_flava_label_smoothing_oracle_matches_reference_golden():
    """cross_entropy_kwargs={"label_smoothing": 0.1} through the reference function (make_golden_smoothing.py), incl. the mask."""
    from oracle import loss_oracle as lo
    g = load_golden("losses_smoothing")
    for tag, m in (("", None), ("mask.", g["mask"])):
        a = g["a"].clone().requires_grad_(True); b = g["b"].clone().requires_grad_(True); s = g["scale"].clone().requires_grad_(True)
        loss, _, _, la, lb = lo.contrastive_loss_with_temperature(a, b, s, mask=m, label_smoothing=0.1)
        loss.backward()
        assert abs(loss.item() - g[tag + "loss"].item()) < 1e-6 and abs(la.item() - g[tag + "loss_a"].item()) < 1e-6
        assert torch.allclose(a.grad, g[tag + "da"], atol=1e-7) and torch.allclose(b.grad, g[tag + "db"], atol=1e-7)
        assert abs(s.grad.item() - g[tag + "dscale"].item()) < 1e-5


def test_storage_aware_expert_reduces_to_the_closed_form():
    """`expert_forward_storage_aware` (the oracle variant that rounds Y and Z at the CUDA path's two storage points, used to pin
    the attn_proj.0 gradient) is the closed form exactly when nothing is rounded, stays within bf16 rounding of it otherwise,
    and its rounding is invisible to autograd (straight-through): gradients exist and are finite for every parameter."""
    import torch
    from oracle import moe_oracle as mo
    params = mo.init_params(2, [8, 16, 32, 64], 64, 64, seed=3)
    torch.manual_seed(4)
    feats = [torch.randn(3, n, d) for n, d in zip([64, 16, 4, 1], [8, 16, 32, 64])]
    ref = mo.expert_forward_closed_form(params, 1, feats)
    same = mo.expert_forward_storage_aware(params, 1, feats, storage=None)
    assert (ref - same).abs().max().item() <= 1e-6 * ref.abs().max().item()
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    stored = mo.expert_forward_storage_aware(pr, 1, feats, storage=torch.bfloat16)
    rel = ((stored.detach() - ref).norm() / ref.norm()).item()
    assert 0.0 < rel < 1e-2, rel
    stored.square().sum().backward()
    for k, v in pr.items():
        if k.startswith("experts.1.") and not k.endswith("attn_proj.2.bias"):
            assert v.grad is not None and torch.isfinite(v.grad).all() and v.grad.abs().max() > 0, k
    # and through the routed block
    (gf, lf, probs), idx = mo.moe_forward_sparse(params, feats, torch.randn(3, 64), storage=torch.bfloat16)
    assert gf.shape == (3, 64) and torch.isfinite(lf).all()
