"""Word-patch attention loss (medmoe_b200.GLORIALocalContrastiveLoss) at training size: time per step and per kernel.

    python tools/local_loss_bench.py [--batch 256] [--words 25] [--img 224] [--steps 3] [--torch-batch 32]

Also times the reference's algorithm as stock PyTorch ops on the same GPU (losses.py:954-1026 restated in this file: a
Python loop over captions, fp32 bmm + softmax) at --torch-batch (its activations do not fit at 256) for a scale-free
comparison in pairs^2 per second.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medmoe_b200  # noqa: E402
from medmoe_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--words", type=int, default=25)
ap.add_argument("--img", type=int, default=224)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--torch-batch", type=int, default=32)
args = ap.parse_args()

B, L, D, H = args.batch, args.words, 768, args.img // 4
g = torch.Generator(device="cuda").manual_seed(12345)
fused = (torch.randn(B, H * H, D, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
local = fused.transpose(1, 2).reshape(B, D, H, H)           # the stride view medmoe_b200.MoE returns
words = torch.randn(B, D, L, device="cuda", generator=g) * 0.3
cap_lens = [L] * B
loss_mod = medmoe_b200.GLORIALocalContrastiveLoss(return_att_maps=False)


def step():
    x = local.detach().requires_grad_(True)
    w = words.detach().requires_grad_(True)
    out = loss_mod(x, w, cap_lens)
    (out.loss0 + out.loss1).backward()
    return out.loss0.detach() + out.loss1.detach()


step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30

prof = _lib.EventProfiler()
_lib.PROFILER = prof
step()
torch.cuda.synchronize()
_lib.PROFILER = None
kern = {k: {"calls": n, "ms": round(t, 3)} for k, (n, t) in sorted(prof.summary().items(), key=lambda kv: -kv[1][1])}
Wp = (L + 7) // 8 * 8
flops = 6 * 2.0 * B * ((H * H + 127) // 128 * 128) * D * ((B + 15) // 16 * 16) * Wp      # six GEMMs of the padded problem
useful = 6 * 2.0 * B * H * H * D * B * L

res = {"metric": "local loss fwd+bwd", "batch": B, "words": L, "tokens": H * H, "ms_per_step": round(ms, 2),
       "pairs_per_s": round(B / (ms * 1e-3), 1), "gemm_tflops_executed": round(flops / (ms * 1e-3) / 1e12, 1),
       "useful_tflops": round(useful / (ms * 1e-3) / 1e12, 1), "peak_mem_gb": round(peak_gb, 1), "loss": float(loss),
       "kernel_ms_sum": round(sum(v["ms"] for v in kern.values()), 2), "kernels": kern}

if args.torch_batch > 0:
    import torch.nn.functional as F

    def reference_algorithm(img, wrd, lens, temp1=4.0, temp2=5.0, temp3=10.0, eps=1e-8):
        """The reference's own schedule as stock torch ops (losses.py:954-1026, attention_fn :698-736): a Python loop over the
        captions, each repeated over the batch, two bmm + two softmax per caption.  Baseline leg of this tool only."""
        Bn, Dn = img.shape[:2]
        context = img.reshape(Bn, Dn, -1)
        cols = []
        for i in range(Bn):
            n = lens[i]
            word = wrd[i, :, :n].unsqueeze(0).repeat(Bn, 1, 1)
            attn = torch.softmax(torch.bmm(context.transpose(1, 2), word), dim=-1).transpose(1, 2)
            attn = torch.softmax(attn * temp1, dim=-1)
            wc = torch.bmm(context, attn.transpose(1, 2))
            cos = (word * wc).sum(1) / (word.norm(dim=1) * wc.norm(dim=1)).clamp(min=eps)
            cols.append(torch.log(torch.exp(temp2 * cos).sum(1, keepdim=True)))
        sim = torch.cat(cols, 1) * temp3
        labels = torch.arange(Bn, device=sim.device)
        return F.cross_entropy(sim, labels), F.cross_entropy(sim.t(), labels)

    tb = args.torch_batch
    xi = local[:tb].float().contiguous()
    wi = words[:tb].contiguous()

    def torch_step():
        x = xi.detach().requires_grad_(True)
        w = wi.detach().requires_grad_(True)
        l0, l1 = reference_algorithm(x, w, cap_lens[:tb])
        (l0 + l1).backward()

    torch_step()
    torch.cuda.synchronize()
    e0.record()
    torch_step()
    e1.record()
    torch.cuda.synchronize()
    tms = e0.elapsed_time(e1)
    res["torch_reference_algorithm"] = {"batch": tb, "ms_per_step": round(tms, 2),
                                        "pair_pairs_per_s": round(tb * tb / (tms * 1e-3)),
                                        "note": "fp32 (TF32 off) loop over captions on the same GPU"}
    res["pair_pairs_per_s"] = round(B * B / (ms * 1e-3))
print(json.dumps(res))
