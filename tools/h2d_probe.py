"""How fast does this box move a pinned host batch to the GPU? (one tensor vs the bench's seven, one vs two streams)"""
import torch
n = 290_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: d.copy_(h, non_blocking=True))
print(f"one copy: {ms:.3f} ms {n / ms / 1e6:.1f} GB/s")
s2 = torch.cuda.Stream()
half = n // 2
def two():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        d[half:].copy_(h[half:], non_blocking=True)
    d[:half].copy_(h[:half], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
ms = t(two)
print(f"two streams: {ms:.3f} ms {n / ms / 1e6:.1f} GB/s")
