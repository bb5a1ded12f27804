import os, sys, torch
sys.path.insert(0, "/root/repo")
import medmoe_b200
from torch.profiler import profile, ProfilerActivity
B, L, D, H = 256, 25, 768, 56
g = torch.Generator(device="cuda").manual_seed(1)
fused = (torch.randn(B, H * H, D, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
local = fused.transpose(1, 2).reshape(B, D, H, H)
words = torch.randn(B, D, L, device="cuda", generator=g) * 0.3
mod = medmoe_b200.GLORIALocalContrastiveLoss(return_att_maps=False)
def step():
    x = local.detach().requires_grad_(True); w = words.detach().requires_grad_(True)
    o = mod(x, w, [L] * B); (o.loss0 + o.loss1).backward()
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
