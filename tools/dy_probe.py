"""dY GEMM (rank-1 aux + gate epilogue, K = 384 -> N = 768) over the cfg2 row space: ms per launch for the CTA-pair kernel and the
single-CTA kernel.  MEDMOE_LIB=<variant .so> to time tuning variants (medmoe_b200/build.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medmoe_b200 import _lib, ops, plan as mmplan  # noqa: E402

B, E, D, H = 256, 4, 768, 384
Ps = [3136, 784, 196, 49]


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
    layout = mmplan.make_layout(B, 1, E, Ps)
    plan = mmplan.build_plan(item_expert, layout)
    rows = layout.total_rows
    gg = torch.Generator(device="cuda").manual_seed(1)
    dZ = torch.randn(rows, H, device="cuda", generator=gg).to(torch.bfloat16)
    W1T = (torch.randn(E * D, H, device="cuda", generator=gg) * H ** -0.5).to(torch.bfloat16)
    Y = torch.relu(torch.randn(rows, D, device="cuda", generator=gg)).to(torch.bfloat16)
    coef = torch.randn(rows, device="cuda", generator=gg)
    rimg = ((torch.arange(rows, device="cuda") // Ps[0]) % B).to(torch.int32)      # consecutive rows share a vector, as in the step
    vecs = torch.randn(B, D, device="cuda", generator=gg)
    out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)

    def run(flags):
        return lambda: ops.gemm_rows_rank1(dZ, W1T, D, out, plan=plan, tile_begin=0, tile_count=layout.total_tiles, row_coef=coef,
                                           row_vec=rimg, vecs=vecs, gate=Y, flags=flags)
    t_pair, t_single = timed(run(ops.EPI_PAIR_OK)), timed(run(0))
    nbytes = rows * (H + 2 * D) * 2
    print(f"dY pair  : {t_pair:.3f} ms  {nbytes / t_pair / 1e6:6.0f} GB/s")
    print(f"dY single: {t_single:.3f} ms  {nbytes / t_single / 1e6:6.0f} GB/s")


if __name__ == "__main__":
    main()
