"""One warm-up + one measured MoE fwd/bwd step at the bench shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medmoe_b200  # noqa: E402

B = int(os.environ.get("B", "256"))
K, hidden, D, Ps = 4, [96, 192, 384, 768], 768, [3136, 784, 196, 49]
torch.manual_seed(0)
moe = medmoe_b200.MoE(num_experts=K).cuda()
g = torch.Generator(device="cuda").manual_seed(1)
feats = [torch.randn(B, p, d, device="cuda", generator=g).to(torch.bfloat16) for p, d in zip(Ps, hidden)]
sw = torch.randn(B, D, device="cuda", generator=g)
local = os.environ.get("LOCAL_GRAD", "0") == "1"
for it in range(2):
    moe.zero_grad(set_to_none=True)
    fg = [f.detach().requires_grad_(True) for f in feats]
    gf, lf, probs = moe(fg, sw)
    loss = gf.float().square().mean() + probs.square().mean()
    if local:
        loss = loss + lf.float().square().mean()
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
