"""Correctness + throughput probe of the CTA-pair row GEMM (gemm_pair.cuh), dense problems."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medmoe_b200 import _lib  # noqa: E402


def run(M, K, N, reps=0):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(1, N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)

    def call():
        _lib.call("mm_grouped_gemm_rows", _lib.ptr(A), M, K, K, _lib.ptr(W), 1, N, K, 0, 0, 0, M, _lib.ptr(bias), 0, 0, 0, 0,
                  _lib.ptr(out), N, 0, 0, 1.0, 1, _lib.stream_ptr())
    call()
    torch.cuda.synchronize()
    ref = torch.relu(A.float() @ W.float().t() + bias)
    err = (out.float() - ref).abs().max().item()
    bad = int((~torch.isfinite(out.float())).sum())
    msg = f"M={M} K={K} N={N}: max err {err:.4g} nonfinite {bad}"
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += f"  {ms:.3f} ms {2.0 * M * K * N / ms / 1e9:.0f} TFLOP/s"
    print(msg, flush=True)


for on in (0, 1):
    _lib.load().mm_debug_gemm_pair(on)
    print("pair mode", on, flush=True)
    for shape in [(256, 64, 192), (128, 128, 256), (1000, 768, 384), (4096, 768, 768), (300, 96, 768)]:
        run(*shape)
    run(1 << 20, 768, 384, reps=5)
    run(1 << 20, 768, 768, reps=5)
