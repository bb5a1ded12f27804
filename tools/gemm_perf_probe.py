"""Throughput of gemm_rows / gemm_wgrad over tile widths (dense problem, M = 1M rows): TFLOP/s per (K, N)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medmoe_b200 import _lib  # noqa: E402

M = int(os.environ.get("M", 1 << 20))


def run(K, N, reps=5):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    st = _lib.stream_ptr()

    def call():
        _lib.call("mm_grouped_gemm_rows", _lib.ptr(A), M, K, K, _lib.ptr(W), 1, N, K, 0, 0, 0, M, 0, 0, 0, 0, 0,
                  _lib.ptr(out), N, 0, 0, 1.0, 0, st)
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * M * K * N / (ms * 1e-3) / 1e12
    gbs = (M * K * 2 + M * N * 2) / (ms * 1e-3) / 1e9
    print(f"rows K={K:4d} N={N:4d}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s  {gbs:7.0f} GB/s (A + out)", flush=True)


for K, N in [(768, 128), (768, 256), (768, 384), (768, 512), (768, 768), (384, 768), (1536, 256), (1536, 384), (96, 768)]:
    run(K, N)
