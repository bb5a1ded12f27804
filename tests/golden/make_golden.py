"""Generate the committed golden vectors from the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Writes small .npz fixtures next to this file.  Every array is produced by the reference's own
`MoE` (src/models/components/swin.py), `GLORIAGlobalContrastiveLoss`,
`FLAVAGlobalContrastiveLoss` / `contrastive_loss_with_temperature` (src/losses.py) on seeded
CPU fp32 inputs — these files are what pins the oracle (the reference ships no tests).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_shim as rs  # noqa: E402


def npify(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def round_expert_weights_(moe):
    """Round the GEMM operands (conv / first attention Linear weights) to bf16-representable fp32 values."""
    with torch.no_grad():
        for ex in moe.experts:
            for seq in ex.proj_convs:
                seq[0].weight.copy_(seq[0].weight.to(torch.bfloat16).float())
            ex.attn_proj[0].weight.copy_(ex.attn_proj[0].weight.to(torch.bfloat16).float())


def moe_case(name, K, hidden, D, Ps, B, seed, store_params=True, round_bf16=False):
    swin = rs.load_moe_module()
    torch.manual_seed(seed)
    moe = swin.MoE(num_experts=K, hidden_dims=hidden, output_dim=D, router_input_dim=D)
    if round_bf16:
        round_expert_weights_(moe)
    torch.manual_seed(seed + 1)
    feats = [torch.randn(B, p, d) for p, d in zip(Ps, hidden)]
    if round_bf16:   # the reference (fp32 arithmetic) evaluated at the operands the bf16 kernels actually see
        feats = [f.to(torch.bfloat16).float() for f in feats]
    feats = [f.requires_grad_(True) for f in feats]
    sw = torch.randn(B, D, requires_grad=True)
    labels = torch.randint(0, K, (B,))
    cg = torch.randn(B, D)                        # cotangent of global_feat
    P = max(Ps)
    cl = torch.randn(B, D, int(P ** 0.5), int(P ** 0.5)) / P   # cotangent of local_feat
    g, l, probs = moe(feats, sw)
    ce = torch.nn.functional.cross_entropy(probs, labels)      # medmoe_module.py:235-237
    obj = (g * cg).sum() + (l * cl).sum() + 2.0 * ce
    obj.backward()
    out = {"global_feat": g, "local_feat": l.contiguous(), "probs": probs, "top_expert": torch.argmax(probs, -1),
           "router_ce": ce, "labels": labels, "cot_global": cg, "cot_local": cl, "swin_feat": sw, "d_swin_feat": sw.grad}
    for s, f in enumerate(feats):
        out[f"feat{s}"] = f
        out[f"d_feat{s}"] = f.grad
    out["init_seed"] = torch.tensor(seed)
    out["round_bf16"] = torch.tensor(round_bf16)
    for k, v in moe.state_dict().items():
        if store_params:
            out["param." + k] = v
        else:   # weights are re-created from `init_seed` (same nn modules, same order => same RNG stream); checksum pins it
            out["psum." + k] = v.double().sum()
    for k, p in moe.named_parameters():
        out["gradnone." + k] = torch.tensor(p.grad is None)
        g = p.grad if p.grad is not None else torch.zeros(0)
        out["gradnorm." + k] = g.double().norm()
        if store_params or g.numel() <= 4096:
            out["grad." + k] = g
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(out))
    print(name, "top_expert", out["top_expert"].tolist())


def loss_cases():
    L = rs.load_losses_module()
    torch.manual_seed(0)
    out = {}
    for B in (16, 37):
        I = torch.randn(B, 768, requires_grad=True)
        T = torch.randn(B, 768, requires_grad=True)
        loss = L.GLORIAGlobalContrastiveLoss()(I, T, temp3=10.0)
        loss.backward()
        out.update({f"gloria{B}.img": I, f"gloria{B}.txt": T, f"gloria{B}.loss": loss, f"gloria{B}.dimg": I.grad,
                    f"gloria{B}.dtxt": T.grad})
    # FLAVA single process (incl. the mask path)
    I = torch.randn(24, 768, requires_grad=True)
    T = torch.randn(24, 768, requires_grad=True)
    fl = L.FLAVAGlobalContrastiveLoss()
    o = fl(I, T)
    o.loss.backward()
    out.update({"flava.img": I, "flava.txt": T, "flava.loss": o.loss, "flava.image_loss": o.image_loss,
                "flava.text_loss": o.text_loss, "flava.image_logits": o.image_logits, "flava.text_logits": o.text_logits,
                "flava.dimg": I.grad, "flava.dtxt": T.grad, "flava.dscale": fl.logit_scale.grad,
                "flava.logit_scale": fl.logit_scale.data})
    mask = torch.rand(24) > 0.4
    I2 = I.detach().clone().requires_grad_(True)
    T2 = T.detach().clone().requires_grad_(True)
    fl2 = L.FLAVAGlobalContrastiveLoss()
    o2 = fl2(I2, T2, mask=mask)
    o2.loss.backward()
    out.update({"flava_mask.mask": mask, "flava_mask.loss": o2.loss, "flava_mask.dimg": I2.grad, "flava_mask.dtxt": T2.grad,
                "flava_mask.dscale": fl2.logit_scale.grad, "flava_mask.image_logits": o2.image_logits})

    # FLAVA multi-rank: run the reference function once per emulated rank with torch.distributed
    # faked in-process (is_initialized/get_rank patched, gather_tensor returning every rank's tensor,
    # the local one with gradient) — i.e. the reference's own label/offset/gather logic, W = 3.
    import torch.distributed as dist
    W, Bl = 3, 8
    a_parts = [torch.randn(Bl, 768) for _ in range(W)]
    b_parts = [torch.randn(Bl, 768) for _ in range(W)]
    a_parts = [torch.nn.functional.normalize(x, dim=-1) for x in a_parts]
    b_parts = [torch.nn.functional.normalize(x, dim=-1) for x in b_parts]
    scale = torch.tensor(2.0, requires_grad=True)
    leafs_a = [x.clone().requires_grad_(True) for x in a_parts]
    leafs_b = [x.clone().requires_grad_(True) for x in b_parts]
    orig = (dist.is_available, dist.is_initialized, dist.get_rank, L.gather_tensor)
    losses = []
    try:
        dist.is_available = lambda: True
        dist.is_initialized = lambda: True
        for r in range(W):
            dist.get_rank = lambda r=r: r
            pool = {id(leafs_a[r]): leafs_a, id(leafs_b[r]): leafs_b}
            L.gather_tensor = lambda t, bp, pool=pool: list(pool[id(t)])   # GLOBAL backprop: every rank's live tensor
            o = L.contrastive_loss_with_temperature(leafs_a[r], leafs_b[r], scale)
            losses.append(o.loss)
            out[f"flava_mr.logits_a{r}"] = o.logits_a
    finally:
        dist.is_available, dist.is_initialized, dist.get_rank, L.gather_tensor = orig
    torch.stack(losses).mean().backward()      # DDP averages gradients over ranks
    out.update({"flava_mr.scale": scale.detach(), "flava_mr.dscale": scale.grad,
                "flava_mr.losses": torch.stack(losses)})
    for r in range(W):
        out[f"flava_mr.a{r}"] = a_parts[r]; out[f"flava_mr.b{r}"] = b_parts[r]
        out[f"flava_mr.da{r}"] = leafs_a[r].grad; out[f"flava_mr.db{r}"] = leafs_b[r].grad
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **npify(out))
    print("losses: gloria16", out["gloria16.loss"].item(), "flava", out["flava.loss"].item())


if __name__ == "__main__" and os.environ.get("GOLDEN_ONLY", "") != "local":
    assert rs.available(), "needs /root/reference"
    # small-width MoE (keeps the fixture ~2 MB) incl. experts that receive no image
    moe_case("moe_small", K=3, hidden=[32, 64, 128, 256], D=256, Ps=[64, 16, 4, 1], B=6, seed=0)
    # reference default widths, tiny token counts, 6 experts (several idle -> zero grads)
    moe_case("moe_k6", K=6, hidden=[96, 192, 384, 768], D=768, Ps=[16, 4, 1, 1], B=4, seed=2, store_params=False)
    # the same two cases with bf16-representable GEMM operands: isolates kernel arithmetic from the
    # ReLU-gate flips that operand rounding causes (tests/test_moe_gpu.py explains the tolerances)
    moe_case("moe_small_bf16w", K=3, hidden=[32, 64, 128, 256], D=256, Ps=[64, 16, 4, 1], B=6, seed=0, round_bf16=True)
    moe_case("moe_k6_bf16w", K=6, hidden=[96, 192, 384, 768], D=768, Ps=[16, 4, 1, 1], B=4, seed=2, store_params=False,
             round_bf16=True)
    loss_cases()


def local_loss_case():
    """GLORIALocalContrastiveLoss (src/losses.py:954-1026) on seeded inputs, ragged caption lengths."""
    losses = rs.load_losses_module()
    torch.manual_seed(21)
    B, D, H, L = 5, 128, 6, 9
    img = (0.5 * torch.randn(B, D, H, H)).requires_grad_(True)
    words = (0.5 * torch.randn(B, D, L)).requires_grad_(True)
    cap_lens = [9, 4, 7, 1, 6]
    out = {}
    for agg in ("sum", "mean"):
        img.grad = None
        words.grad = None
        res = losses.GLORIALocalContrastiveLoss()(img, words, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg=agg)
        (res.loss0 + res.loss1).backward()
        out.update({f"loss0_{agg}": res.loss0, f"loss1_{agg}": res.loss1, f"d_img_{agg}": img.grad.clone(),
                    f"d_words_{agg}": words.grad.clone(), f"att0_{agg}": res.att_maps[0], f"att3_{agg}": res.att_maps[3]})
    out.update({"img": img, "words": words, "cap_lens": torch.tensor(cap_lens)})
    np.savez_compressed(os.path.join(HERE, "local_loss.npz"), **npify(out))
    print("local_loss.npz", float(out["loss0_sum"]), float(out["loss1_sum"]))


if __name__ == "__main__":       # GOLDEN_ONLY=local regenerates just this fixture
    local_loss_case()
