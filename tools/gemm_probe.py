"""Diagnostic probe for the tcgen05 GEMM kernels (run on the GPU box, writes gpurun_out/gemm_probe.log).

For each case it reports max error and, when wrong, where the produced values sit in the
reference matrix (reveals descriptor / swizzle / layout mistakes as index permutations).
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medmoe_b200 import _lib, plan as mmplan  # noqa: E402

LOG = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.append(s)


def explain(got, ref, name, max_show=12):
    err = (got - ref).abs()
    bad = err > 1e-2 * max(1.0, ref.abs().max().item())
    nbad = int(bad.sum())
    log(f"  [{name}] shape={tuple(ref.shape)} max_err={err.max().item():.4g} bad={nbad}/{ref.numel()}"
        f" nan={int(torch.isnan(got).sum())}")
    if nbad == 0:
        return True
    rows_bad = bad.any(1).nonzero().flatten()
    cols_bad = bad.any(0).nonzero().flatten()
    log(f"    bad rows: n={rows_bad.numel()} first={rows_bad[:16].tolist()}  bad cols: n={cols_bad.numel()} first={cols_bad[:16].tolist()}")
    # where do the produced values live in the reference?
    flat_ref = ref.flatten()
    idx = bad.nonzero()[:max_show]
    for m, n in idx.tolist():
        v = got[m, n].item()
        d = (flat_ref - v).abs()
        j = int(d.argmin())
        log(f"    got[{m},{n}]={v:.5f} ref={ref[m, n].item():.5f}; nearest ref value at ({j // ref.shape[1]},{j % ref.shape[1]}) diff={d[j].item():.3g}")
    return False


def bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def rows_case(M, K, N, out_f32=True):
    A = bf16(M, K, seed=1)
    W = bf16(N, K, seed=2, scale=K ** -0.5)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    t0 = time.time()
    _lib.call("mm_grouped_gemm_rows", _lib.ptr(A), M, K, A.stride(0), _lib.ptr(W), 1, N, W.stride(0), 0, 0, 0, M,
              0, 0, 0, 0, 0, _lib.ptr(out), out.stride(0), int(out_f32), 0, 1.0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    log(f"rows M={M} K={K} N={N} f32={out_f32} ({(time.time() - t0) * 1e3:.1f} ms)")
    return explain(out.float(), ref, "rows")


def wgrad_case(rows_per, E, N1, N2):
    n_items = 8
    layout = mmplan.make_layout(n_items, 1, E, [rows_per], target_chunks=2)
    g = torch.Generator().manual_seed(5)
    item_expert = torch.randint(0, E, (n_items,), generator=g, dtype=torch.int32).cuda()
    plan = mmplan.build_plan(item_expert, layout)
    torch.cuda.synchronize()
    rows = layout.total_rows
    row_e = torch.full((rows,), -1, dtype=torch.long)
    for t, (e, v) in enumerate(plan.tile_info.cpu().tolist()):
        if e >= 0:
            row_e[t * 128:t * 128 + v] = e
    row_e = row_e.cuda()
    A = bf16(rows, N1, seed=3)
    A[row_e < 0] = 0
    B = bf16(rows, N2, seed=4)
    out = torch.zeros(E, N1, N2, device="cuda")
    _lib.call("mm_grouped_gemm_wgrad", _lib.ptr(A), rows, N1, A.stride(0), _lib.ptr(B), rows, N2, B.stride(0),
              _lib.ptr(plan.chunks), 0, layout.total_chunks, 0, _lib.ptr(out), _lib.stream_ptr())
    torch.cuda.synchronize()
    log(f"wgrad rows_per={rows_per} E={E} N1={N1} N2={N2} chunks={plan.chunks.cpu().tolist()[:6]}")
    ok = True
    for e in range(E):
        m = row_e == e
        ref = A[m].float().t() @ B[m].float()
        ok &= explain(out[e], ref, f"wgrad e={e}")
    return ok


def main():
    log("device", torch.cuda.get_device_name(0), "sms", _lib.call("mm_device_sm_count"))
    results = {}
    for case in [(128, 64, 32), (128, 64, 256), (128, 128, 64), (256, 768, 384), (300, 96, 768), (4096, 768, 768)]:
        try:
            results[("rows",) + case] = rows_case(*case)
        except Exception as ex:  # noqa: BLE001
            log("EXC rows", case, repr(ex))
            results[("rows",) + case] = False
    for case in [(100, 1, 128, 64), (300, 2, 128, 256), (700, 2, 384, 768), (500, 3, 768, 96)]:
        try:
            results[("wgrad",) + case] = wgrad_case(*case)
        except Exception as ex:  # noqa: BLE001
            log("EXC wgrad", case, repr(ex))
            results[("wgrad",) + case] = False
    log("SUMMARY", results)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_probe.log", "w") as f:
        f.write("\n".join(LOG) + "\n")


if __name__ == "__main__":
    main()
