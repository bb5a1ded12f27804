"""2 real GPUs, NCCL: the embedding exchange of the global contrastive loss exactly as the product runs it — ONE packed
all-gather of cat([a, b], 1) forward, ONE reduce-scatter(SUM) backward (medmoe_b200/losses.py, medmoe_b200/distributed.py;
reference src/utils/distributed.py:28-58 + src/losses.py:503-524) — feeding the fused InfoNCE kernels, against the
single-process oracle on concatenated embeddings (oracle.loss_oracle.flava_multi_rank).  Also the overlapped gradient
all-reduce (OverlappedGradSync) against plain averaging.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

WORLD = 2
B, D = 48, 768


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        import medmoe_b200
        from oracle import loss_oracle as lo
        g = torch.Generator().manual_seed(100)
        img_all = torch.randn(WORLD * B, D, generator=g)
        txt_all = torch.randn(WORLD * B, D, generator=g)
        mask_all = torch.rand(WORLD * B, generator=g) > 0.3

        # ---- oracle: single process, concatenated problem; DDP averages the per-rank losses ----
        ir = img_all.clone().requires_grad_(True); tr = txt_all.clone().requires_grad_(True)
        sr = torch.tensor(lo.DEFAULT_LOGIT_SCALE, requires_grad=True)
        ia, tb = torch.nn.functional.normalize(ir, dim=-1), torch.nn.functional.normalize(tr, dim=-1)
        losses, mean = lo.flava_multi_rank(list(ia.chunk(WORLD)), list(tb.chunk(WORLD)), sr)
        mean.backward()

        # ---- product path on this rank's GPU over NCCL ----
        mod = medmoe_b200.FLAVAGlobalContrastiveLoss().to(dev)
        i_loc = img_all[rank * B:(rank + 1) * B].to(dev).requires_grad_(True)
        t_loc = txt_all[rank * B:(rank + 1) * B].to(dev).requires_grad_(True)
        out = mod(i_loc, t_loc)
        assert abs(out.loss.item() - losses[rank].item()) < 1e-5 * abs(losses[rank].item()), (out.loss.item(), losses[rank].item())
        ref_logits = (ia[rank * B:(rank + 1) * B] @ tb.t() * torch.exp(sr)).detach()
        assert (out.image_logits.cpu() - ref_logits).abs().max().item() < 1e-4
        (out.loss / WORLD).backward()            # DDP's gradient averaging of the loss
        dist.all_reduce(mod.logit_scale.grad, op=dist.ReduceOp.SUM)

        def rel(a, b):
            return ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
        assert rel(i_loc.grad, ir.grad[rank * B:(rank + 1) * B]) < 1e-4
        assert rel(t_loc.grad, tr.grad[rank * B:(rank + 1) * B]) < 1e-4
        assert abs(mod.logit_scale.grad.item() - sr.grad.item()) < 1e-4 * max(1.0, abs(sr.grad.item()))

        # ---- the exchange ran over NVLink peer memory (csrc/p2p.cu), and that path equals the NCCL collectives bit for bit
        #      (two ranks: the sum of two floats has one order), eagerly and under CUDA-graph replay ----
        from medmoe_b200 import distributed as mmd
        assert mmd.PeerExchange.last_backend == "p2p", mmd.PeerExchange._disabled_reason
        xs = torch.zeros(B, 2 * D, device=dev, requires_grad=True)
        gs = torch.zeros(WORLD * B, 2 * D, device=dev)

        def exchange():
            y = mmd.all_gather_cat(xs)
            (dx,) = torch.autograd.grad(y, xs, gs)
            return y, dx

        def fill(it):
            gx = torch.Generator().manual_seed(1000 + 10 * it + rank)
            xs.data.copy_(torch.randn(B, 2 * D, generator=gx))
            gs.copy_(torch.randn(WORLD * B, 2 * D, generator=gx))

        def check(y, dx, what):
            ref_y = torch.empty_like(gs)
            dist.all_gather_into_tensor(ref_y, xs.detach())
            ref_dx = torch.empty(B, 2 * D, device=dev)
            dist.reduce_scatter_tensor(ref_dx, gs.clone(), op=dist.ReduceOp.SUM)
            assert torch.equal(y, ref_y), what
            assert torch.equal(dx, ref_dx), what
        for it in range(3):
            fill(it)
            check(*exchange(), f"eager {it}")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            exchange()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y_s, dx_s = exchange()
        for it in range(5):                        # the device-side step counters keep the parity right across replays
            fill(100 + it)
            graph.replay()
            check(y_s, dx_s, f"replay {it}")
        os.environ["MEDMOE_P2P_EXCHANGE"] = "0"   # same call through NCCL
        fill(7)
        check(*exchange(), "nccl")
        assert mmd.PeerExchange.last_backend == "nccl"
        os.environ["MEDMOE_P2P_EXCHANGE"] = "1"

        # ---- mask path over NCCL: mean over the selected local rows (losses.py:574-577) ----
        m_loc = mask_all[rank * B:(rank + 1) * B]
        ref_m = lo.contrastive_loss_with_temperature(ia[rank * B:(rank + 1) * B].detach(), tb[rank * B:(rank + 1) * B].detach(),
                                                     sr.detach(), ia.detach(), tb.detach(), rank=rank, mask=m_loc)[0]
        out_m = medmoe_b200.FLAVAGlobalContrastiveLoss().to(dev)(i_loc.detach(), t_loc.detach(), mask=m_loc.to(dev))
        assert abs(out_m.loss.item() - ref_m.item()) < 1e-5 * abs(ref_m.item())

        # ---- overlapped gradient all-reduce of a small MoE == plain averaging of the per-rank gradients ----
        torch.manual_seed(5)
        moe = medmoe_b200.MoE(num_experts=3).to(dev)
        Ps = [64, 16, 4, 1]
        gl = torch.Generator().manual_seed(200 + rank)
        feats = [torch.randn(5, p, d, generator=gl).to(dev) for p, d in zip(Ps, [96, 192, 384, 768])]
        sw = torch.randn(5, D, generator=gl).to(dev)

        def fwd_bwd():
            moe.zero_grad(set_to_none=True)
            gf, lf, probs = moe([f.clone().requires_grad_(True) for f in feats], sw)
            (gf.square().sum() + lf.float().square().sum() * 1e-3 + probs[:, 0].sum()).backward()
        fwd_bwd()                                   # no hook: this rank's own gradients
        want = {}
        for k, p_ in moe.named_parameters():
            w = p_.grad.clone()
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
            want[k] = w / WORLD
        sync = medmoe_b200.OverlappedGradSync(moe)
        fwd_bwd()                                   # hook: the experts' flat bucket is all-reduced from inside the backward
        flat = moe.last_flat_grad
        lo_, hi_ = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        n_view = sum(lo_ <= p_.grad.data_ptr() < hi_ for p_ in moe.experts.parameters())
        sync.finish()
        torch.cuda.synchronize()
        for k, p_ in moe.named_parameters():
            if want[k].abs().max() == 0:
                assert p_.grad.abs().max() == 0, k
            else:
                assert rel(p_.grad, want[k].cpu()) < 1e-5, (k, rel(p_.grad, want[k].cpu()))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"views={n_view}")
    finally:
        dist.destroy_process_group()


def test_flava_loss_and_grad_sync_over_nccl_world2(tmp_path):
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(WORLD))
