"""GPU unit tests of the tcgen05 grouped GEMM kernels through the C-ABI.

Reference here is a plain PyTorch fp32 matmul of the same bf16-rounded operands (a
floating-point kernel: tolerance = bf16 output rounding, 2^-8 relative, plus fp32
accumulation-order noise).
"""
import pytest
import torch

from medmoe_b200 import _lib, plan as mmplan

pytestmark = pytest.mark.gpu

EPI_RELU, EPI_ZERO_PAD = 1, 2


def _bf16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def gemm_rows(A, W, N, *, tile_info=None, tile_begin=0, tile_count=0, M=0, bias=None, aux=None, gate=None,
              out=None, out_f32=False, colsum=None, flags=0, out_scale=1.0):
    E = W.shape[0] // N
    K = A.shape[1]
    _lib.call("mm_grouped_gemm_rows", _lib.ptr(A), A.shape[0], K, A.stride(0), _lib.ptr(W), E, N, W.stride(0),
              _lib.ptr(tile_info), tile_begin, tile_count, M, _lib.ptr(bias),
              _lib.ptr(aux), aux.stride(0) if aux is not None else 0,
              _lib.ptr(gate), gate.stride(0) if gate is not None else 0,
              _lib.ptr(out), out.stride(0), int(out_f32), _lib.ptr(colsum), float(out_scale), flags,
              _lib.stream_ptr())
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,K,N", [(128, 64, 256), (300, 96, 768), (517, 192, 768), (1000, 768, 384),
                                   (256, 384, 96), (130, 768, 192), (64, 768, 128), (2048, 768, 2048)])
@pytest.mark.parametrize("out_f32", [False, True])
def test_dense_rows_gemm(M, K, N, out_f32):
    A = _bf16(M, K, seed=1)
    W = _bf16(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    gemm_rows(A, W, N, M=M, bias=bias, out=out, out_f32=out_f32, flags=EPI_RELU)
    ref = torch.relu(A.float() @ W.float().t() + bias)
    got = out.float()
    assert torch.isfinite(got).all()
    tol = 2e-5 if out_f32 else 2 ** -7
    err = (got - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def _random_plan(n_items, K, P, seed=0):
    g = torch.Generator().manual_seed(seed)
    item_expert = torch.randint(0, K, (n_items,), generator=g, dtype=torch.int32)
    layout = mmplan.make_layout(n_items, 1, K, P, target_chunks=4)
    plan = mmplan.build_plan(item_expert.cuda(), layout)
    torch.cuda.synchronize()
    return item_expert, layout, plan


def test_dispatch_build_matches_host_logic():
    for seed, (n, K, P) in enumerate([(7, 4, [49, 13]), (64, 6, [196, 49, 16]), (1500, 8, [3, 1]), (5, 3, [300])]):
        item_expert, layout, plan = _random_plan(n, K, P, seed)
        ref = mmplan.reference_plan(item_expert.tolist(), layout)
        assert plan.counts.tolist() == ref["counts"]
        assert plan.offsets.tolist() == ref["offsets"]
        assert plan.perm.tolist() == ref["perm"]
        assert plan.inv_perm.tolist() == ref["inv_perm"]
        assert plan.slot_expert.tolist() == ref["slot_expert"]
        assert plan.seg_start.tolist() == ref["seg_start"]
        assert plan.slot_row.tolist() == ref["slot_row"]
        assert [tuple(t) for t in plan.tile_info.tolist()] == ref["tile_info"]
        assert [tuple(t) for t in plan.chunks.tolist()] == ref["chunks"]


def _row_expert(layout, plan):
    """expert id per global row (-1 = padding / unused), from the device-built tables."""
    row_e = torch.full((layout.total_rows,), -1, dtype=torch.long)
    ti = plan.tile_info.cpu()
    for t in range(layout.total_tiles):
        e, v = ti[t].tolist()
        if e >= 0:
            row_e[t * 128: t * 128 + v] = e
    return row_e.cuda()


@pytest.mark.parametrize("K,N", [(96, 768), (768, 384), (384, 768), (768, 192)])
def test_grouped_rows_gemm_full_epilogue(K, N):
    n_items, E, P = 37, 4, [49, 20]
    _, layout, plan = _random_plan(n_items, E, P, seed=3)
    rows = layout.total_rows
    A = _bf16(rows, K, seed=4)
    W = _bf16(E * N, K, scale=K ** -0.5, seed=5)
    bias = torch.randn(E, N, device="cuda")
    aux = _bf16(rows, N, seed=6)
    gate = torch.relu(_bf16(rows, N, seed=7))
    out = torch.full((rows, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    colsum = torch.zeros(E, N, device="cuda")
    gemm_rows(A, W, N, tile_info=plan.tile_info, tile_begin=0, tile_count=layout.total_tiles, bias=bias, aux=aux,
              gate=gate, out=out, colsum=colsum, flags=EPI_ZERO_PAD)
    row_e = _row_expert(layout, plan)
    valid = row_e >= 0
    Wf = W.float().view(E, N, K)
    ref = torch.zeros(rows, N, device="cuda")
    for e in range(E):
        m = row_e == e
        ref[m] = (A[m].float() @ Wf[e].t() + bias[e] + aux[m].float()) * (gate[m] > 0)
    got = out.float()
    # tiles that belong to an expert are fully written (valid rows: result, padding: zeros)
    ti = plan.tile_info.cpu()
    owned = torch.zeros(rows, dtype=torch.bool)
    for t in range(layout.total_tiles):
        if ti[t, 0] >= 0:
            owned[t * 128:(t + 1) * 128] = True
    owned = owned.cuda()
    assert torch.isfinite(got[owned]).all()
    assert (got[owned & ~valid] == 0).all()
    err = (got[valid] - ref[valid]).abs().max().item()
    assert err <= 2 ** -7 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    # with aux/gate the column sums are those of the values actually written (gated, bf16-rounded)
    got_cs = torch.stack([got[row_e == e].sum(0) for e in range(E)])
    assert torch.allclose(colsum, got_cs, rtol=1e-4, atol=1e-3)
    ref_cs = torch.stack([ref[row_e == e].sum(0) for e in range(E)])
    assert rel_err_t(colsum, ref_cs) < 5e-3


def rel_err_t(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("N1,N2", [(768, 96), (768, 192), (768, 384), (768, 768), (384, 64)])
def test_grouped_wgrad_with_column_sums(N1, N2):
    """mm_grouped_gemm_wgrad_colsum: dW plus the per-expert column sums of A (bias gradients) from the same MMAs."""
    n_items, E, P = 29, 3, [196, 49]
    _, layout, plan = _random_plan(n_items, E, P, seed=18)
    rows = layout.total_rows
    row_e = _row_expert(layout, plan)
    A = _bf16(rows, N1, seed=19)
    A[row_e < 0] = 0
    Bm = _bf16(rows, N2, seed=20)
    out = torch.zeros(E, N1, N2, device="cuda")
    cs = torch.zeros(E, N1, device="cuda")
    _lib.call("mm_grouped_gemm_wgrad_colsum", _lib.ptr(A), rows, N1, A.stride(0), _lib.ptr(Bm), rows, N2, Bm.stride(0),
              _lib.ptr(plan.chunks), 0, layout.total_chunks, 0, _lib.ptr(out), _lib.ptr(cs), _lib.stream_ptr())
    torch.cuda.synchronize()
    for e in range(E):
        m = row_e == e
        ref = A[m].float().t() @ Bm[m].float()
        err = (out[e] - ref).abs().max().item()
        assert err <= 1e-3 * max(1.0, ref.abs().max().item()), f"expert {e}: max abs err {err}"
        ref_cs = A[m].float().sum(0)
        err = (cs[e] - ref_cs).abs().max().item()
        assert err <= 1e-3 * max(1.0, ref_cs.abs().max().item()), f"expert {e}: column sums, max abs err {err}"


@pytest.mark.parametrize("N1,N2", [(384, 768), (768, 96), (768, 192), (768, 384), (768, 768)])
def test_grouped_wgrad(N1, N2):
    n_items, E, P = 29, 3, [196, 49]
    _, layout, plan = _random_plan(n_items, E, P, seed=8)
    rows = layout.total_rows
    row_e = _row_expert(layout, plan)
    A = _bf16(rows, N1, seed=9)
    A[row_e < 0] = 0          # contract: the A-side operand is zero in padding rows
    Bm = _bf16(rows, N2, seed=10)
    out = torch.zeros(E, N1, N2, device="cuda")
    _lib.call("mm_grouped_gemm_wgrad", _lib.ptr(A), rows, N1, A.stride(0), _lib.ptr(Bm), rows, N2, Bm.stride(0),
              _lib.ptr(plan.chunks), 0, layout.total_chunks, 0, _lib.ptr(out), _lib.stream_ptr())
    torch.cuda.synchronize()
    for e in range(E):
        m = row_e == e
        ref = A[m].float().t() @ Bm[m].float()
        err = (out[e] - ref).abs().max().item()
        assert err <= 1e-3 * max(1.0, ref.abs().max().item()), f"expert {e}: max abs err {err}"


def test_wgrad_region_subrange():
    """Per-scale launch: chunk sub-range + tile_base offset + pointer offset into the row space."""
    n_items, E, P = 21, 3, [100, 36]
    _, layout, plan = _random_plan(n_items, E, P, seed=11)
    s = 1
    r0, r1 = layout.region_base[s], layout.region_base[s] + layout.region_rows[s]
    row_e = _row_expert(layout, plan)[r0:r1]
    A = _bf16(r1 - r0, 768, seed=12)
    A[row_e < 0] = 0
    Bm = _bf16(r1 - r0, 192, seed=13)
    out = torch.zeros(E, 768, 192, device="cuda")
    _lib.call("mm_grouped_gemm_wgrad", _lib.ptr(A), r1 - r0, 768, A.stride(0), _lib.ptr(Bm), r1 - r0, 192, Bm.stride(0),
              _lib.ptr(plan.chunks), layout.chunk_base[s], layout.chunk_cap[s], layout.tile_base[s], _lib.ptr(out),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    for e in range(E):
        m = row_e == e
        ref = A[m].float().t() @ Bm[m].float()
        err = (out[e] - ref).abs().max().item()
        assert err <= 1e-3 * max(1.0, ref.abs().max().item()), f"expert {e}: max abs err {err}"


@pytest.mark.parametrize("K,N,with_aux", [(384, 768, False), (768, 384, False), (128, 256, False), (384, 768, True)])
def test_grouped_rows_gemm_rank1_aux(K, N, with_aux):
    """mm_grouped_gemm_rows_rank1: out = (A W_e^T + row_coef[row] * vecs[row_vec[row]]) * [gate > 0]; tiles no expert owns
    come back zero-filled (the combine kernels stage whole row ranges and rely on finite contents)."""
    n_items, E, P = 37, 4, [49, 20]
    _, layout, plan = _random_plan(n_items, E, P, seed=23)
    rows = layout.total_rows
    A = _bf16(rows, K, seed=24)
    W = _bf16(E * N, K, scale=K ** -0.5, seed=25)
    gate = torch.relu(_bf16(rows, N, seed=27))
    g = torch.Generator(device="cuda").manual_seed(28)
    n_vec = 11
    vecs = torch.randn(n_vec, N, device="cuda", generator=g)
    row_coef = torch.randn(rows, device="cuda", generator=g)
    row_vec = torch.randint(0, n_vec, (rows,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.full((rows, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    aux = _bf16(rows, N, seed=29) if with_aux else None
    _lib.call("mm_grouped_gemm_rows_rank1", _lib.ptr(A), rows, K, A.stride(0), _lib.ptr(W), E, N, W.stride(0),
              _lib.ptr(plan.tile_info), 0, layout.total_tiles, _lib.ptr(row_coef), _lib.ptr(row_vec), _lib.ptr(vecs),
              vecs.stride(0), _lib.ptr(aux), N if with_aux else 0, _lib.ptr(gate), gate.stride(0), _lib.ptr(out),
              out.stride(0), 0, 0, _lib.stream_ptr())
    if not with_aux and N % 256 == 0:
        # the same launch on CTA pairs (cta_group::2, flag MM_EPI_PAIR_OK: the plan's segments are 256-row aligned): bit-identical
        out2 = torch.full((rows, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        _lib.call("mm_grouped_gemm_rows_rank1", _lib.ptr(A), rows, K, A.stride(0), _lib.ptr(W), E, N, W.stride(0),
                  _lib.ptr(plan.tile_info), 0, layout.total_tiles, _lib.ptr(row_coef), _lib.ptr(row_vec), _lib.ptr(vecs),
                  vecs.stride(0), 0, 0, _lib.ptr(gate), gate.stride(0), _lib.ptr(out2), out2.stride(0), 0, 8, _lib.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(out.view(torch.int16), out2.view(torch.int16))
    torch.cuda.synchronize()
    row_e = _row_expert(layout, plan)
    valid = row_e >= 0
    Wf = W.float().view(E, N, K)
    ref = torch.zeros(rows, N, device="cuda")
    for e in range(E):
        m = row_e == e
        ref[m] = (A[m].float() @ Wf[e].t() + row_coef[m, None] * vecs[row_vec[m].long()]
                  + (aux[m].float() if with_aux else 0.0)) * (gate[m] > 0)
    got = out.float()
    assert torch.isfinite(got).all()                     # every tile is written: owned (result / zero padding) or zero-filled
    assert (got[~valid] == 0).all()
    err = (got[valid] - ref[valid]).abs().max().item()
    assert err <= 2 ** -7 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.fixture
def pair_mode():
    """CTA-pair kernels on (the default) for launches that pass MM_EPI_PAIR_OK; tests switch them off to get the single-CTA result."""
    lib = _lib.load()
    lib.mm_debug_gemm_pair(3)
    yield
    lib.mm_debug_gemm_pair(3)


@pytest.mark.parametrize("M,K,N", [(256, 64, 192), (128, 128, 256), (300, 96, 768), (1000, 768, 384), (4096 + 130, 384, 768)])
def test_pair_rows_gemm_dense_bit_identical_to_single(pair_mode, M, K, N):
    A = _bf16(M, K, seed=1)
    W = _bf16(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm_rows(A, W, N, M=M, bias=bias, out=out, flags=EPI_RELU | 8)
    _lib.load().mm_debug_gemm_pair(0)
    single = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm_rows(A, W, N, M=M, bias=bias, out=single, flags=EPI_RELU | 8)
    # same k order, same fp32 accumulation: the two tile shapes must agree bit for bit
    assert torch.equal(out, single)
    ref = torch.relu(A.float() @ W.float().t() + bias)
    assert (out.float() - ref).abs().max().item() <= 2 ** -7 * max(1.0, ref.abs().max().item())


def test_pair_rows_gemm_grouped_256_row_segments(pair_mode):
    # hand-made tile table: every expert segment starts on an even tile (256-row alignment), ragged ends, holes
    E, K, N = 3, 768, 384
    tiles = [(0, 128), (0, 77), (1, 128), (-1, 0), (-1, 0), (-1, 0), (2, 5), (-1, 0), (1, 128), (1, 128)]
    tile_info = torch.tensor(tiles, dtype=torch.int32, device="cuda")
    rows = len(tiles) * 128
    A = _bf16(rows, K, seed=4)
    W = _bf16(E * N, K, scale=K ** -0.5, seed=5)
    bias = torch.randn(E, N, device="cuda")
    out = torch.full((rows, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm_rows(A, W, N, tile_info=tile_info, tile_begin=0, tile_count=len(tiles), bias=bias, out=out, flags=EPI_RELU | 8)
    Wf = W.float().view(E, N, K)
    for t, (e, v) in enumerate(tiles):
        blk = out[t * 128:(t + 1) * 128].float()
        assert torch.isfinite(blk).all(), f"tile {t} not written"
        if e < 0:
            assert (blk == 0).all()
            continue
        ref = torch.relu(A[t * 128: t * 128 + v].float() @ Wf[e].t() + bias[e])
        assert (blk[v:] == 0).all()
        assert (blk[:v] - ref).abs().max().item() <= 2 ** -7 * max(1.0, ref.abs().max().item())


# ---- back-to-back expert GEMMs (csrc/b2b.cuh): E1 -> E4 in one kernel -------------------------------------------------
def _b2b_case(n_items, E, P, K1, seed, idle_expert=False):
    from medmoe_b200 import ops
    D, H = 768, 384
    g = torch.Generator().manual_seed(seed)
    item_expert = torch.randint(1 if idle_expert else 0, E, (n_items,), generator=g, dtype=torch.int32)
    layout = mmplan.make_layout(n_items, 1, E, P, target_chunks=4)
    plan = mmplan.build_plan(item_expert.cuda(), layout)
    rows = layout.total_rows
    row_e = _row_expert(layout, plan)
    f = _bf16(rows, K1, seed=seed + 1)
    f[row_e < 0] = 0
    Wp = _bf16(E * D, K1, scale=K1 ** -0.5, seed=seed + 2)
    W1 = _bf16(E * H, D, scale=D ** -0.5, seed=seed + 3)
    bp = torch.randn(E, D, device="cuda")
    b1 = torch.randn(E, H, device="cuda")
    Y = torch.full((rows, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    Z = torch.full((rows, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    assert _lib.call("mm_expert_b2b_fwd_supported", K1, D, H) >= 1      # 2: only on CTA pairs (K1 in (128, 192])
    ops.expert_b2b_fwd(f, Wp, bp, W1, b1, Y, Z, plan=plan, tile_begin=0, tile_count=layout.total_tiles)
    torch.cuda.synchronize()
    return layout, plan, row_e, f, Wp, W1, bp, b1, Y, Z


@pytest.mark.parametrize("K1", [96, 64, 128, 32, 192, 144])
@pytest.mark.parametrize("idle", [False, True])
def test_b2b_forward_bit_identical_to_the_two_gemms(K1, idle):
    """The fused kernel runs the same MMAs in the same k order on the same bf16-rounded Y: Y and Z must be bit-identical
    to mm_grouped_gemm_rows(E1) followed by mm_grouped_gemm_rows(E4), including zeroed padding rows and unowned tiles."""
    D, H = 768, 384
    layout, plan, row_e, f, Wp, W1, bp, b1, Y, Z = _b2b_case(41, 4, [196, 49], K1, seed=30 + K1, idle_expert=idle)
    rows = layout.total_rows
    Y2 = torch.full((rows, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    Z2 = torch.full((rows, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm_rows(f, Wp, D, tile_info=plan.tile_info, tile_begin=0, tile_count=layout.total_tiles, bias=bp, out=Y2,
              flags=EPI_RELU | EPI_ZERO_PAD)
    gemm_rows(Y2, W1, H, tile_info=plan.tile_info, tile_begin=0, tile_count=layout.total_tiles, bias=b1, out=Z2,
              flags=EPI_ZERO_PAD)
    assert torch.isfinite(Y.float()).all() and torch.isfinite(Z.float()).all()      # every tile written, owned or not
    assert torch.equal(Y.view(torch.int16), Y2.view(torch.int16))
    assert torch.equal(Z.view(torch.int16), Z2.view(torch.int16))
    assert (Y[row_e < 0] == 0).all() and (Z[row_e < 0] == 0).all()


def test_b2b_forward_vs_fp32_matmul():
    D, H, E, K1 = 768, 384, 3, 96
    layout, plan, row_e, f, Wp, W1, bp, b1, Y, Z = _b2b_case(23, E, [784, 196], K1, seed=77)
    Wpf, W1f = Wp.float().view(E, D, K1), W1.float().view(E, H, D)
    for e in range(E):
        m = row_e == e
        y_ref = torch.relu(f[m].float() @ Wpf[e].t() + bp[e])
        assert (Y[m].float() - y_ref).abs().max().item() <= 2 ** -7 * max(1.0, y_ref.abs().max().item())
        z_ref = Y[m].float() @ W1f[e].t() + b1[e]            # GEMM 2 consumes the bf16-rounded Y, like the unfused path
        assert (Z[m].float() - z_ref).abs().max().item() <= 2 ** -7 * max(1.0, z_ref.abs().max().item())


def test_b2b_forward_many_tiles_per_cta():
    """More tiles than SMs: every CTA walks several tiles, so the barrier phases wrap many times."""
    layout, plan, row_e, f, Wp, W1, bp, b1, Y, Z = _b2b_case(24, 4, [3136], 96, seed=5)
    assert layout.total_tiles >= 4 * 148
    E, D, H, K1 = 4, 768, 384, 96
    Wpf, W1f = Wp.float().view(E, D, K1), W1.float().view(E, H, D)
    for e in range(E):
        m = row_e == e
        y_ref = torch.relu(f[m].float() @ Wpf[e].t() + bp[e])
        assert (Y[m].float() - y_ref).abs().max().item() <= 2 ** -7 * max(1.0, y_ref.abs().max().item())
        z_ref = Y[m].float() @ W1f[e].t() + b1[e]
        assert (Z[m].float() - z_ref).abs().max().item() <= 2 ** -7 * max(1.0, z_ref.abs().max().item())
