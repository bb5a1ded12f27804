"""CPU, world_size 2 over gloo: the embedding exchange of the global contrastive loss
(medmoe_b200.distributed) — forward all-gather, backward reduce-scatter(SUM), LOCAL / NONE
modes, rank-offset labels — against the single-process computation on concatenated embeddings
(the preferred multi-rank oracle).  Only host logic and torch.distributed run here;
the CUDA kernels are exercised by the -m gpu tests with the same rank/offset arithmetic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from medmoe_b200.distributed import BackpropType, all_gather_cat, concat_gather_all_gpu, gather_tensor, get_rank
        from oracle import loss_oracle as lo
        assert get_rank() == rank
        g = torch.Generator().manual_seed(100)
        all_a = torch.nn.functional.normalize(torch.randn(WORLD * 4, 16, generator=g), dim=-1)
        all_b = torch.nn.functional.normalize(torch.randn(WORLD * 4, 16, generator=g), dim=-1)
        a = all_a[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
        b = all_b[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
        scale = torch.tensor(2.0, requires_grad=True)

        # GLOBAL: gathered tensors equal the concatenation, gradient = sum over ranks of d loss_r / d a (reduce-scatter)
        ga, gb = all_gather_cat(a, BackpropType.GLOBAL), all_gather_cat(b, BackpropType.GLOBAL)
        assert torch.equal(ga.detach(), all_a) and torch.equal(gb.detach(), all_b)
        loss = lo.contrastive_loss_with_temperature(a, b, scale, ga, gb, rank=rank)[0]
        loss.backward()
        # single-process reference: sum of every rank's loss w.r.t. this rank's rows
        ra = all_a.clone().requires_grad_(True); rb = all_b.clone().requires_grad_(True); rs = torch.tensor(2.0, requires_grad=True)
        losses, mean = lo.flava_multi_rank(list(ra.chunk(WORLD)), list(rb.chunk(WORLD)), rs)
        torch.stack(losses).sum().backward()
        assert abs(loss.item() - losses[rank].item()) < 1e-6
        assert torch.allclose(a.grad, ra.grad[rank * 4:(rank + 1) * 4], atol=1e-6)
        assert torch.allclose(b.grad, rb.grad[rank * 4:(rank + 1) * 4], atol=1e-6)

        # list form and concat helper mirror the reference API
        parts = gather_tensor(a.detach(), BackpropType.NONE)
        assert len(parts) == WORLD and torch.equal(torch.cat(parts), all_a)
        assert torch.equal(concat_gather_all_gpu(a.detach(), BackpropType.NONE), all_a)

        # LOCAL: gradient flows only into this worker's own block
        a2 = all_a[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
        gl = all_gather_cat(a2, BackpropType.LOCAL)
        assert torch.equal(gl.detach(), all_a)
        w = torch.arange(1, WORLD * 4 + 1, dtype=torch.float32).unsqueeze(1)
        (gl * w).sum().backward()
        assert torch.allclose(a2.grad, w[rank * 4:(rank + 1) * 4].expand(4, 16))

        # NONE: no gradient at all
        a3 = all_a[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
        assert not all_gather_cat(a3, BackpropType.NONE).requires_grad
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_forward_backward_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(WORLD))


def test_single_process_passthrough():
    from medmoe_b200.distributed import all_gather_cat, gather_tensor, get_rank
    x = torch.randn(3, 5)
    assert get_rank() == 0
    assert all_gather_cat(x) is x and len(gather_tensor(x)) == 1
