"""Diagnostic: per-tensor relative errors of medmoe_b200.MoE vs the oracle, with localisation.
Run on the GPU box; prints to stdout and gpurun_out/moe_diag.log."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medmoe_b200  # noqa: E402
from oracle import moe_oracle as mo  # noqa: E402

LOG = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.append(s)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(K, hidden, D, Ps, B, seed, use_g=True, use_l=True, use_ce=True, bf16_oracle_inputs=False):
    params = mo.init_params(K, hidden, D, D, seed=seed)
    torch.manual_seed(seed + 1)
    feats = [torch.randn(B, p, d) for p, d in zip(Ps, hidden)]
    sw = torch.randn(B, D)
    labels = torch.randint(0, K, (B,))
    P = max(Ps)
    cg = torch.randn(B, D) * float(use_g)
    cl = torch.randn(B, D, int(P ** 0.5), int(P ** 0.5)) / P * float(use_l)
    if bf16_oracle_inputs:   # oracle on bf16-rounded weights and inputs: isolates intermediate-rounding effects
        params = {k: (v.to(torch.bfloat16).float() if k.startswith("experts") and "weight" in k and "attn_proj.2" not in k else v)
                  for k, v in params.items()}
        feats = [f.to(torch.bfloat16).float() for f in feats]
    pr = {k: v.clone().double().requires_grad_(True) for k, v in params.items()}
    fr = [f.clone().double().requires_grad_(True) for f in feats]
    sr = sw.clone().double().requires_grad_(True)
    (gf, lf, probs), idx = mo.moe_forward_sparse(pr, fr, sr)
    obj = (gf * cg.double()).sum() + (lf * cl.double()).sum() + 2.0 * float(use_ce) * mo.router_ce(probs, labels)
    obj.backward()

    moe = medmoe_b200.MoE(num_experts=K, hidden_dims=hidden, output_dim=D, router_input_dim=D)
    moe.load_state_dict(params)
    moe = moe.cuda()
    fg = [f.cuda().requires_grad_(True) for f in feats]
    sg = sw.cuda().requires_grad_(True)
    g2, l2, p2 = moe(fg, sg)
    ce = torch.nn.functional.cross_entropy(p2, labels.cuda())
    obj2 = (g2 * cg.cuda()).sum() + (l2 * cl.cuda()).sum() + 2.0 * float(use_ce) * ce
    obj2.backward()
    log(f"--- K={K} D={D} Ps={Ps} B={B} g={use_g} l={use_l} ce={use_ce} bf16in={bf16_oracle_inputs} top={idx[:, 0].tolist()}")
    log("  global_feat", rel(g2, gf), "local_feat", rel(l2, lf), "probs", (p2.cpu().double() - probs).abs().max().item())
    for s in range(4):
        e = rel(fg[s].grad, fr[s].grad)
        log(f"  d_feat{s} rel={e:.4g} |ref|={fr[s].grad.norm().item():.4g}")
        if s == 0 or e > 0.02:
            d = (fg[s].grad.double().cpu() - fr[s].grad)
            per_tok = d.norm(dim=-1) / fr[s].grad.norm(dim=-1).clamp_min(1e-30)       # [B, P_s]
            worst = per_tok.flatten().topk(min(8, per_tok.numel()))
            log(f"    per-token rel err: median={per_tok.median().item():.4g} mean={per_tok.mean().item():.4g} max={per_tok.max().item():.4g}"
                f" worst idx={[(int(i) // per_tok.shape[1], int(i) % per_tok.shape[1]) for i in worst.indices]}")
            per_ch = d.norm(dim=(0, 1)) / fr[s].grad.norm(dim=(0, 1)).clamp_min(1e-30)
            log(f"    per-channel rel err: median={per_ch.median().item():.4g} max={per_ch.max().item():.4g}")
    log("  d_swin", rel(sg.grad, sr.grad))
    for k, p in moe.named_parameters():
        ref = pr[k].grad if pr[k].grad is not None else torch.zeros_like(pr[k])
        if ref.norm() == 0:
            log(f"  {k}: ref zero, got max {p.grad.abs().max().item():.3g}")
        else:
            log(f"  {k}: rel={rel(p.grad, ref):.4g}")


if __name__ == "__main__":
    full = [96, 192, 384, 768]
    run(2, full, 768, [64, 16, 4, 1], 4, 0)
    run(2, full, 768, [64, 16, 4, 1], 4, 0, bf16_oracle_inputs=True)
    run(2, full, 768, [64, 16, 4, 1], 4, 0, use_l=False, use_ce=False)
    run(2, full, 768, [64, 16, 4, 1], 4, 0, use_g=False, use_ce=False)
    run(2, full, 768, [3136, 784, 196, 49], 2, 1, bf16_oracle_inputs=True)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/moe_diag.log", "w").write("\n".join(LOG) + "\n")
