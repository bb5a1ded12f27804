#!/usr/bin/env python
"""bench.py — train pairs/sec of the MoE block + global InfoNCE fwd/bwd on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1): BASELINE config 2 — batch 256 / GPU, 4 modality experts, multi-scale Swin-T stage
features of a 224^2 image (3136x96, 784x192, 196x384, 49x768 tokens) in bf16, synthetic.
One step = MoE forward -> global InfoNCE (FLAVA semantics: learnable temperature, embeddings
all-gathered over ranks) + router cross-entropy -> backward to every MoE parameter, the four
stage-feature tensors and swin_feat; for N > 1 the parameter gradients are all-reduced (DDP).
Weak scaling: every rank processes its own 256 pairs.

Printed JSON (one line, rank 0): see the contract in the task description; extra keys
`roofline`, `cpu_baseline`, `kernels` (per-kernel share of the step, CUDA events).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN = [96, 192, 384, 768]
D = 768


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="pairs per GPU")
    ap.add_argument("--experts", type=int, default=4)
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--topk", type=int, default=1, help="experts per image (1 = the reference; 2 = BASELINE config 4 extension)")
    ap.add_argument("--loss", default="flava", choices=["flava", "gloria"])
    ap.add_argument("--local-grad", action="store_true", help="also feed a dense synthetic cotangent into local_feat")
    ap.add_argument("--local-loss", action="store_true",
                    help="add the word-patch attention loss (GLORIALocalContrastiveLoss) on local_feat: the reference's full objective")
    ap.add_argument("--words", type=int, default=25, help="words per caption for --local-loss")
    ap.add_argument("--routing", default="natural", choices=["natural", "uniform", "skew"],
                    help="natural = the router as initialised; uniform = every expert gets B/K images; skew = all images to expert 0")
    ap.add_argument("--cpu-sample-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a captured CUDA graph")
    return ap.parse_args()


def token_counts(img):
    p0 = (img // 4) ** 2
    return [p0, p0 // 4, p0 // 16, p0 // 64]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU baseline: the oracle's restatement of the reference's dense schedule, all host threads
# --------------------------------------------------------------------------------------
def cpu_step_fn(batch, experts, img, loss_kind):
    from oracle import loss_oracle as lo
    from oracle import moe_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    Ps = token_counts(img)
    params = {k: v.requires_grad_(True) for k, v in mo.init_params(experts, HIDDEN, D, D, seed=0).items()}
    g = torch.Generator().manual_seed(12345)
    feats = [torch.randn(batch, p, d, generator=g).requires_grad_(True) for p, d in zip(Ps, HIDDEN)]
    sw = torch.randn(batch, D, generator=g).requires_grad_(True)
    txt = torch.randn(batch, D, generator=g)
    labels = torch.randint(0, experts, (batch,), generator=g)
    scale = torch.tensor(lo.DEFAULT_LOGIT_SCALE, requires_grad=True)

    def step():
        for t in list(params.values()) + feats + [sw, scale]:
            t.grad = None
        gf, lf, probs = mo.moe_forward_dense(params, feats, sw)          # reference schedule: all K experts, stack, gather
        if loss_kind == "flava":
            g_loss = lo.flava_global_loss(gf, txt, scale)[0]
        else:
            g_loss = lo.gloria_global_loss(gf, txt, 10.0)
        loss = 0.5 * g_loss + 2.0 * mo.router_ce(probs, labels)          # medmoe_module.py:308 weights (no local loss here)
        loss.backward()
        return float(loss.detach())
    return step


def run_cpu_baseline(args, steps, warmup):
    step = cpu_step_fn(args.cpu_sample_batch, args.experts, args.img, args.loss)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": args.cpu_sample_batch / dt, "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"batch {args.cpu_sample_batch} of the same workload (K={args.experts} dense experts as the reference "
                      f"runs them, {args.img}^2 tokens, fp32, {steps} step(s) after {warmup} warm-up), oracle/moe_oracle.py",
            "s_per_step": dt}


def config_dict(args, world):
    name = "cfg2" if (args.experts == 4 and args.topk == 1 and args.img == 224) else "cfg4-like" if args.topk > 1 else "custom"
    return {"workload": f"{name}: MoE block + global InfoNCE fwd/bwd, batch {args.batch}/GPU, {args.experts} experts top-{args.topk}, "
                        f"Swin-T stage features of a {args.img}^2 image ({'/'.join(map(str, token_counts(args.img)))} tokens), bf16",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "experts": args.experts, "topk": args.topk, "img": args.img,
            "loss": args.loss, "local_cotangent": bool(args.local_grad), "routing": args.routing,
            "local_loss_words": args.words if args.local_loss else 0,     # its word embeddings stay resident (not part of e2e H2D)
            "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": "inputs+intermediates per step (> 5 GB) exceed the 126 MB L2; no explicit flush"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cb = run_cpu_baseline(args, steps, warmup)
    out = {"impl": "reference", "metric": "train pairs/sec (MoE block + global InfoNCE fwd/bwd)", "value": cb["value"],
           "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["s_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": config_dict(args, world), "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "reference's dense CPU schedule (oracle port; the Python reference cannot travel to the GPU box), "
                   f"bounded sample of {args.cpu_sample_batch} pairs/step on {cb['cores']} host threads"}
    print(json.dumps(out))


def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)

    # stdout carries exactly one JSON line: everything libraries print (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    import torch.nn.functional as F

    import medmoe_b200
    from medmoe_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K = args.batch, args.experts
    Ps = token_counts(args.img)
    torch.manual_seed(0)
    moe = medmoe_b200.MoE(num_experts=K, topk=args.topk).to(dev)
    loss_mod = (medmoe_b200.FLAVAGlobalContrastiveLoss() if args.loss == "flava" else medmoe_b200.GLORIAGlobalContrastiveLoss()).to(dev)
    params = [p for p in list(moe.parameters()) + list(loss_mod.parameters())]

    # synthetic batch (seed 12345 + rank, the reference's seed, pretraining_medmoe.yaml:18), pinned on the host
    g = torch.Generator().manual_seed(12345 + rank)
    host = {
        "feats": [torch.randn(B, p, d, generator=g).to(torch.bfloat16).pin_memory() for p, d in zip(Ps, HIDDEN)],
        "sw": torch.randn(B, D, generator=g).pin_memory(),
        "txt": torch.randn(B, D, generator=g).pin_memory(),
        "labels": torch.randint(0, K, (B,), generator=g).pin_memory(),
    }
    local_mod, txt_local, cap_lens = None, None, None
    if args.local_loss:     # synthetic word embeddings [B, 768, W] (the text tower is outside the path), full-length captions
        local_mod = medmoe_b200.GLORIALocalContrastiveLoss(return_att_maps=False)
        txt_local = (0.3 * torch.randn(B, D, args.words, generator=g)).to(dev)
        cap_lens = [args.words] * B
    if args.routing != "natural":
        # routing balance variants: the router stays the real kernel, its weights / input are chosen so that the
        # arg-max is forced.  uniform: swin_feat carries a one-hot of (image index mod K) that the router passes through.
        with torch.no_grad():
            for lin in (moe.router[0], moe.router[2]):
                lin.weight.zero_()
                lin.bias.zero_()
            if args.routing == "skew":
                moe.router[2].bias[0] = 20.0
            else:
                idx = torch.arange(K, device=dev)
                moe.router[0].weight[idx, idx] = 1.0
                moe.router[2].weight[idx, idx] = 1.0
                host["sw"].zero_()
                host["sw"][torch.arange(B), torch.arange(B) % K] = 20.0
    h2d_bytes = sum(t.numel() * t.element_size() for t in host["feats"]) + sum(
        host[k].numel() * host[k].element_size() for k in ("sw", "txt", "labels"))
    cot_local = None
    if args.local_grad:
        cot_local = (torch.randn(B, D, int(Ps[0] ** 0.5), int(Ps[0] ** 0.5), device=dev) / Ps[0]).to(torch.bfloat16)
        cot_local = cot_local.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)   # same strides as local_feat

    def to_device(batch, non_blocking=True):
        return {"feats": [f.to(dev, non_blocking=non_blocking) for f in batch["feats"]],
                "sw": batch["sw"].to(dev, non_blocking=non_blocking), "txt": batch["txt"].to(dev, non_blocking=non_blocking),
                "labels": batch["labels"].to(dev, non_blocking=non_blocking)}

    def step(d):
        for p in params:
            p.grad = None
        feats = [f.detach().requires_grad_(True) for f in d["feats"]]
        sw = d["sw"].detach().requires_grad_(True)
        gf, lf, probs = moe(feats, sw)
        if args.loss == "flava":
            g_loss = loss_mod(gf.float(), d["txt"]).loss
        else:
            g_loss = loss_mod(gf.float(), d["txt"], temp3=10.0)
        loss = 0.5 * g_loss + 2.0 * F.cross_entropy(probs, d["labels"])
        if local_mod is not None:     # medmoe_module.py:220-233, 308: local_loss_weight * (loss0 + loss1)
            l_out = local_mod(lf, txt_local, cap_lens)
            loss = loss + 0.5 * (l_out.loss0 + l_out.loss1)
        if cot_local is not None:     # a dense synthetic cotangent for local_feat, handed straight to autograd (no glue kernels)
            torch.autograd.backward([loss, lf], [None, cot_local])
        else:
            loss.backward()
        if world > 1:   # DDP gradient averaging: one flat bucket, one NCCL all-reduce (AVG) over NVLink, one multi-tensor copy back
            grads = [p.grad for p in params]
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stage(msg):
        if os.environ.get("MEDMOE_BENCH_VERBOSE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    resident = to_device(host, non_blocking=False)
    stage("eager warm-up")
    for _ in range(args.warmup):
        step(resident)
    barrier()
    stage("eager warm-up done")

    # ---------------- whole-step CUDA graph (routing is resolved on the device, so nothing syncs) ----------------
    graph, graph_loss, graph_err = None, None, None
    launches0 = _lib.call("mm_launch_count")
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(resident)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            launches0 = _lib.call("mm_launch_count")
            stage("capturing")
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_loss = step(resident)
            launches = _lib.call("mm_launch_count") - launches0
            stage("captured")
        except Exception as ex:  # noqa: BLE001  (capture is an optimisation; the eager path is always valid)
            graph, graph_err = None, repr(ex)
            torch.cuda.synchronize()

    def run_resident():
        if graph is not None:
            graph.replay()
            return graph_loss
        return step(resident)

    # clocks are sampled from here to the end of the timed region (100 ms period); the warm-up replays keep the
    # GPU under the same load so that short timed regions still see several samples
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 50)):    # same count on every rank: each replay contains collectives
        run_resident()
    barrier()
    stage("graph warm-up done")

    # ---------------- timed region 1: inputs resident in HBM ----------------
    if graph is None:
        launches0 = _lib.call("mm_launch_count")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = run_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    stage(f"timed region 1 done: {ms:.3f} ms/step")
    if graph is None:
        launches = (_lib.call("mm_launch_count") - launches0) // args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- per-kernel CUDA events: the same K steps, eager, one event pair per C-ABI call ----------------
    prof = _lib.EventProfiler()
    _lib.PROFILER = prof
    _lib.load().mm_trace_enable(1)        # composite entry points (combine fwd/bwd) also time their own kernels
    for _ in range(args.steps):
        step(resident)
    barrier()
    _lib.PROFILER = None
    kern = prof.summary()
    sub = _lib.trace_collect()
    _lib.load().mm_trace_enable(0)
    if any(k.startswith("combine_fwd.") for k in sub):
        kern.pop("mm_interp_softmax_combine_fwd", None)
    if any(k.startswith("combine_bwd.") for k in sub):
        kern.pop("mm_interp_softmax_combine_bwd", None)
        kern.pop("mm_interp_softmax_combine_bwd_global", None)
        kern.pop("mm_interp_softmax_combine_bwd_tc", None)
    kern.update(sub)

    # ---------------- timed region 2: end to end from pinned host buffers ----------------
    # H2D of step i+1 (copy stream, into a staging set) overlaps the compute of step i; the graph reads a fixed
    # input set, so each step starts with a device-to-device move staging -> inputs (0.3 GB, ~0.1 ms).
    copy_stream = torch.cuda.Stream()
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    staging = to_device(host, non_blocking=False)
    stage_free = torch.cuda.Event()
    stage_free.record()

    def flat(d):
        return d["feats"] + [d["sw"], d["txt"], d["labels"]]

    def prefetch():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(stage_free)
            for dst, src in zip(flat(staging), flat(host)):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    pending = {"ev": None}

    def e2e_loop(n):
        # steady-state pipeline: every step issues one full H2D (the NEXT batch) and consumes the batch copied during the
        # previous step (the very first one during the warm-up); the loop ends only after its last H2D has landed, so n
        # steps contain exactly n complete host->device copies and n complete fwd+bwd passes.
        if pending["ev"] is None:
            pending["ev"] = prefetch()
        cur = torch.cuda.current_stream()
        for i in range(n):
            cur.wait_event(pending["ev"])
            for dst, src in zip(flat(resident), flat(staging)):
                dst.copy_(src, non_blocking=True)
            stage_free.record(cur)
            pending["ev"] = prefetch()     # next batch's H2D overlaps this step's compute
            loss = run_resident()
            host_loss.copy_(loss.detach(), non_blocking=True)
        cur.wait_event(pending["ev"])
        cur.synchronize()
        return float(host_loss)

    stage("profiled pass done")
    e2e_loop(max(2, args.warmup))
    barrier()
    stage("e2e warm-up done")
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = e2e_loop(args.steps)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / args.steps

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        Bi = B * args.topk                # (image, expert choice) items: the row space holds one set of rows per item
        R = Bi * sum(Ps)
        H = D // 2
        P0 = Ps[0]
        peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        peak_bw = peaks["hbm_gbs"]
        dloc = P0 * D * 2 if args.local_grad else 0
        # algorithmic work per launch: FLOPs executed by the tensor pipe and bytes that must cross HBM
        flops, nbytes = {}, {}
        for s, (p, d_s) in enumerate(zip(Ps, HIDDEN)):
            for tag in ("E1", "dX", "dWp"):
                flops[f"{tag}.s{s}"] = 2.0 * Bi * p * d_s * D
                nbytes[f"{tag}.s{s}"] = Bi * p * (d_s + D) * 2
            flops[f"dY.s{s}"] = 2.0 * Bi * p * H * D
            nbytes[f"dY.s{s}"] = Bi * p * (H + 2 * D) * 2 + (Bi * p * D * 2 if args.local_grad else Bi * p * 8)
        flops["E4"] = flops["dW1"] = flops["dY"] = 2.0 * R * D * H
        nbytes["E4"] = nbytes["dW1"] = R * (D + H) * 2
        nbytes["dY"] = R * (H + 2 * D) * 2 + (R * D * 2 if args.local_grad else R * 8)
        rows_c = sum(Ps) - P0
        for k, v in {
            "combine_fwd.logits": sum(Ps) * H * 2 + P0 * 16,
            "combine_fwd.out": sum(Ps) * D * 2 + P0 * D * 2 / args.topk + P0 * 16,      # out is per image, Y rows per item
            "combine_bwd.dbeta": sum(Ps) * D * 2 + P0 * (16 + 32) + dloc,
            "combine_bwd.dUT": sum(Ps) * D * 2 + P0 * 16 + dloc,
            "combine_bwd.dZ": sum(Ps) * H * 2 * 2 + P0 * (16 + 32),
            "combine_bwd.rowdot": sum(Ps) * D * 2 + sum(Ps) * 8,
            "combine_bwd.dZ.rows": rows_c * H * 2 * 2,
            "combine_bwd.dZ.ident": P0 * H * 2 * 2 + P0 * 16,
            "mm_dispatch_rows": 2 * sum(p * d for p, d in zip(Ps, HIDDEN)) * 2,
            "mm_undispatch_rows": 2 * sum(p * d for p, d in zip(Ps, HIDDEN)) * 2,
        }.items():
            nbytes[k] = v * Bi
        kernels = {}
        total_kernel_ms = sum(t for _, t in kern.values())
        for label, (n, t) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
            tag = label.split(":")[0]
            sec = t / n * 1e-3
            ent = {"calls_per_step": n / args.steps, "ms_per_step": t / args.steps, "share": t / total_kernel_ms}
            t_tensor = flops[tag] / (peak_tf * 1e12) if tag in flops else 0.0
            t_hbm = nbytes[tag] / (peak_bw * 1e9) if tag in nbytes else 0.0
            if tag in flops:
                ent["tflops"] = flops[tag] / sec / 1e12
            if tag in nbytes:
                ent["gbs"] = nbytes[tag] / sec / 1e9
            if t_tensor > 0 or t_hbm > 0:      # the roofline that binds this kernel = the slower of its two floors
                ent["bound"] = "tensor" if t_tensor >= t_hbm else "hbm"
                ent["roofline_frac"] = max(t_tensor, t_hbm) / sec
            kernels[label] = ent
        # dominant kernel -> roofline object
        top_label = next(iter(kernels))
        top = kernels[top_label]
        n_top, t_top = kern[top_label]
        tag = top_label.split(":")[0]
        # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full captures (profiles/), cfg2 only
        known_traffic = {"combine_fwd.out": 2.964e9, "combine_fwd.logits": 0.836e9, "combine_bwd.dZ.rows": 0.402e9,
                         "dY": 4.05e9, "E4": 2.441e9, "E1.s0": 1.337e9, "combine_bwd.dZ.ident": 1.214e9,
                         "combine_bwd.rowdot": 1.645e9, "dW1": 2.650e9}
        if B == 256 and args.img == 224 and args.local_grad:
            known_traffic = {"combine_bwd.dbeta": 2.898e9, "combine_bwd.dUT": 2.838e9}
        traffic = known_traffic.get(tag) if (B == 256 and args.img == 224 and args.topk == 1 and args.experts == 4) else None
        if top.get("bound") == "tensor":
            roof = {"kernel": top_label, "bound": "tensor", "achieved": top["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": top["tflops"] / peak_tf, "traffic": traffic,
                    "peak_source": f"{peaks_kind} (sustained bf16, kernel timed inside a long step)",
                    "flops_per_launch": flops[tag], "avg_launch_ms": t_top / n_top}
        elif top.get("bound") == "hbm":
            roof = {"kernel": top_label, "bound": "hbm", "achieved": top["gbs"], "peak": peak_bw, "unit": "GB/s",
                    "frac": top["gbs"] / peak_bw, "traffic": traffic, "peak_source": peaks_kind,
                    "bytes_per_launch": nbytes[tag], "avg_launch_ms": t_top / n_top,
                    "traffic_source": "profiles/ (ncu --set full dram__bytes_read+write of this kernel), see DESIGN.md"}
            if tag in flops:
                roof["tflops"] = top["tflops"]
        else:
            roof = {"kernel": top_label, "bound": "hbm", "achieved": None, "peak": peak_bw, "unit": "GB/s",
                    "frac": None, "traffic": None}
        gemm_ms = sum(v["ms_per_step"] for k, v in kernels.items() if k.split(":")[0] in flops)
        gemm_flops = sum(flops[k.split(":")[0]] * v["calls_per_step"] for k, v in kernels.items() if k.split(":")[0] in flops)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = run_cpu_baseline(args, 1, 1)
        out = {
            "metric": "train pairs/sec (MoE block + global InfoNCE fwd/bwd)", "value": B * world / (ms * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args, world),
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "last_loss": last,
                    "how": "pinned host batch -> H2D on a copy stream (double-buffered) -> MoE/loss fwd+bwd through "
                           "medmoe_b200.MoE / FLAVAGlobalContrastiveLoss (captured once as a CUDA graph) -> loss D2H, every step; "
                           "steady-state pipeline: step i computes on the batch copied during step i-1 while its own H2D "
                           "(batch i+1) runs, and the timed region ends after its last H2D landed (n steps = n full copies)"},
            "gpu_launches": int(launches), "cuda_graph": graph is not None, "cuda_graph_error": graph_err,
            "clocks": clocks, "roofline": roof,
            "gemm_summary": {"ms_per_step": gemm_ms, "executed_tflop_per_step": gemm_flops / 1e12,
                             "tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                             "frac_of_sustained_peak": (gemm_flops / (gemm_ms * 1e-3) / 1e12) /
                             peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) if gemm_ms else None,
                             "note": "executed FLOPs (attention Linear evaluated at native resolution: 4165 rows/img, not 12544)"},
            "kernels": kernels, "kernel_ms_per_step": total_kernel_ms / args.steps, "loss": float(loss.detach()),
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    sys.stderr.flush()
    if world > 1:
        # captured NCCL work keeps the communicator busy at teardown: drop the graph first and never hang on exit
        graph = None
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
