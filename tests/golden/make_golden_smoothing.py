"""Golden vectors of the reference's `contrastive_loss_with_temperature` WITH `cross_entropy_kwargs={"label_smoothing": 0.1}`
(src/losses.py:527-592, kwargs forwarded to F.cross_entropy at :579-583).  Build container only (needs /root/reference):

    python tests/golden/make_golden_smoothing.py      -> tests/golden/losses_smoothing.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_shim as rs  # noqa: E402


def main():
    L = rs.load_losses_module()
    torch.manual_seed(7)
    B, D = 40, 768
    a = torch.nn.functional.normalize(torch.randn(B, D), dim=-1).requires_grad_(True)
    b = torch.nn.functional.normalize(torch.randn(B, D), dim=-1).requires_grad_(True)
    scale = torch.tensor(2.3, requires_grad=True)
    mask = torch.rand(B) > 0.3
    out = {"a": a.detach(), "b": b.detach(), "scale": scale.detach(), "mask": mask}
    for tag, m in (("", None), ("mask.", mask)):
        for t in (a, b, scale):
            t.grad = None
        o = L.contrastive_loss_with_temperature(a, b, scale, mask=m, cross_entropy_kwargs={"label_smoothing": 0.1})
        o.loss.backward()
        out.update({tag + "loss": o.loss.detach(), tag + "loss_a": o.loss_a.detach(), tag + "loss_b": o.loss_b.detach(),
                    tag + "da": a.grad.clone(), tag + "db": b.grad.clone(), tag + "dscale": scale.grad.clone()})
    np.savez_compressed(os.path.join(HERE, "losses_smoothing.npz"),
                        **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("losses_smoothing.npz", float(out["loss"]), float(out["mask.loss"]))


if __name__ == "__main__":
    main()
