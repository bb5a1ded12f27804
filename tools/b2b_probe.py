"""Fused E1 -> E4 kernel (csrc/b2b.cuh) against the two grouped GEMMs it replaces, on the finest scale of cfg2
(256 images x 3136 rows, K1 = 96, 4 experts): ms per launch, HBM GB/s and executed TFLOP/s.

    python tools/b2b_probe.py [images] [K1]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medmoe_b200 import _lib, ops, plan as mmplan  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K1 = int(sys.argv[2]) if len(sys.argv) > 2 else 96
E, D, H, P = 4, 768, 384, 3136


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    g = torch.Generator().manual_seed(0)
    item_expert = torch.randint(0, E, (B,), generator=g, dtype=torch.int32).cuda()
    layout = mmplan.make_layout(B, 1, E, [P])
    plan = mmplan.build_plan(item_expert, layout)
    rows, tiles = layout.total_rows, layout.total_tiles
    gg = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn(rows, K1, device="cuda", generator=gg).to(torch.bfloat16)
    Wp = (torch.randn(E * D, K1, device="cuda", generator=gg) * K1 ** -0.5).to(torch.bfloat16)
    W1 = (torch.randn(E * H, D, device="cuda", generator=gg) * D ** -0.5).to(torch.bfloat16)
    bp = torch.randn(E, D, device="cuda", generator=gg)
    b1 = torch.randn(E, H, device="cuda", generator=gg)
    Y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    Z = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    Y2, Z2 = torch.empty_like(Y), torch.empty_like(Z)

    def fused():
        ops.expert_b2b_fwd(f, Wp, bp, W1, b1, Y, Z, plan=plan, tile_begin=0, tile_count=tiles)

    def e1():
        ops.gemm_rows(f, Wp, D, Y2, plan=plan, tile_begin=0, tile_count=tiles, bias=bp, flags=ops.EPI_RELU | ops.EPI_ZERO_PAD)

    def e4():
        ops.gemm_rows(Y2, W1, H, Z2, plan=plan, tile_begin=0, tile_count=tiles, bias=b1, flags=ops.EPI_ZERO_PAD)

    t_f, t_1, t_4 = timed(fused), timed(e1), timed(e4)
    same = torch.equal(Y.view(torch.int16), Y2.view(torch.int16)) and torch.equal(Z.view(torch.int16), Z2.view(torch.int16))
    n = B * P
    flops = 2.0 * n * D * (K1 + H)
    bytes_f = n * 2 * (K1 + D + H)
    bytes_2 = n * 2 * (K1 + D + D + H)
    print(f"rows {n} ({tiles} tiles, {tiles / 148:.1f} per SM), K1 = {K1}; bit-identical: {same}")
    print(f"fused  : {t_f:.3f} ms  {flops / t_f / 1e9:7.0f} TFLOP/s  {bytes_f / t_f / 1e6:6.0f} GB/s  ({t_f * 1e3 / (tiles / 148):.2f} us per tile)")
    print(f"E1 + E4: {t_1:.3f} + {t_4:.3f} = {t_1 + t_4:.3f} ms  {flops / (t_1 + t_4) / 1e9:7.0f} TFLOP/s  {bytes_2 / (t_1 + t_4) / 1e6:6.0f} GB/s")


if __name__ == "__main__":
    main()
