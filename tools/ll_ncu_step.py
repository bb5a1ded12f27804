"""One warm-up + one measured fwd/bwd of the word-patch attention loss at a reduced batch (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medmoe_b200  # noqa: E402

B, L, D, H = int(os.environ.get("B", "64")), int(os.environ.get("W", "25")), 768, 56
g = torch.Generator(device="cuda").manual_seed(1)
fused = (torch.randn(B, H * H, D, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
local = fused.transpose(1, 2).reshape(B, D, H, H)
words = torch.randn(B, D, L, device="cuda", generator=g) * 0.3
mod = medmoe_b200.GLORIALocalContrastiveLoss(return_att_maps=False)
for _ in range(2):
    x = local.detach().requires_grad_(True)
    w = words.detach().requires_grad_(True)
    o = mod(x, w, [L] * B)
    (o.loss0 + o.loss1).backward()
torch.cuda.synchronize()
print("ok", float(o.loss0.detach()))
