"""Which part of the bench step breaks CUDA-graph capture?  python tools/capture_debug.py"""
import os, sys, traceback
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medmoe_b200

dev = torch.device("cuda", 0)
B, K, D = 64, 4, 768
Ps, HID = [3136, 784, 196, 49], [96, 192, 384, 768]
torch.manual_seed(0)
moe = medmoe_b200.MoE(num_experts=K).to(dev)
for ret_logits in (True, False):
    for keep_last in (False, True):
        loss_mod = medmoe_b200.FLAVAGlobalContrastiveLoss().to(dev)
        loss_mod.return_logits = ret_logits
        params = list(moe.parameters()) + list(loss_mod.parameters())
        d = {"feats": [torch.randn(B, p, w, device=dev).to(torch.bfloat16) for p, w in zip(Ps, HID)], "sw": torch.randn(B, D, device=dev),
             "txt": torch.randn(B, D, device=dev), "labels": torch.randint(0, K, (B,), device=dev)}
        holder = {}

        def step():
            for p in params:
                p.grad = None
            feats = [f.detach().requires_grad_(True) for f in d["feats"]]
            sw = d["sw"].detach().requires_grad_(True)
            gf, lf, probs = moe(feats, sw)
            loss = 0.5 * loss_mod(gf, d["txt"]).loss + 2.0 * F.cross_entropy(probs, d["labels"])
            loss.backward()
            if keep_last:
                holder["last"] = (gf.detach(), probs.detach())
            return loss
        for _ in range(3):
            l0 = step(); first = l0.detach().clone(); del l0
        pass
        torch.cuda.synchronize()
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                sl = step()
            gr.replay()
            torch.cuda.synchronize()
            print(f"ret_logits={ret_logits} keep_last={keep_last}: capture OK loss {float(sl):.5f} first {float(first):.5f}", flush=True)
        except Exception:
            print(f"ret_logits={ret_logits} keep_last={keep_last}: capture FAILED", flush=True)
            traceback.print_exc()
            break
