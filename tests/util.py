"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def golden_params(g):
    """state_dict of a MoE golden case: stored, or re-created from its init seed (checksummed)."""
    params = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    if params:
        return params
    import medmoe_b200
    K = g["probs"].shape[1]
    hidden = [g[f"feat{s}"].shape[2] for s in range(4)]
    D = g["global_feat"].shape[1]
    torch.manual_seed(int(g["init_seed"]))
    moe = medmoe_b200.MoE(num_experts=K, hidden_dims=hidden, output_dim=D, router_input_dim=g["swin_feat"].shape[1])
    if "round_bf16" in g and bool(g["round_bf16"]):
        with torch.no_grad():
            for ex in moe.experts:
                for seq in ex.proj_convs:
                    seq[0].weight.copy_(seq[0].weight.to(torch.bfloat16).float())
                ex.attn_proj[0].weight.copy_(ex.attn_proj[0].weight.to(torch.bfloat16).float())
    params = {k: v.detach().clone() for k, v in moe.state_dict().items()}
    for k, v in params.items():
        assert abs(v.double().sum().item() - g["psum." + k].item()) < 1e-9, f"init stream drifted for {k}"
    return params


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_err_scaled(a, b):
    """max |a - b| / max |b|"""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()
