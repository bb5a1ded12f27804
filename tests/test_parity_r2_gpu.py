"""GPU parity tests added in round 2 (VERDICT r1 "close the parity holes"):
  (a) BASELINE cfg2 at its REAL size (B = 256, K = 4, 224^2): 8 random images of the batch against the oracle
      (images are independent, so their outputs / feature gradients do not depend on the other 248);
  (b) BASELINE cfg4 geometry proper: K = 8, top-2, 9216/2304/576/144 tokens;
  (c) NaN / Inf in swin_feat: no fault, expert 0 like torch.argmax, NaN probabilities propagate (reference swin.py:99-100);
  (d) router near-tie report (|p1 - p2| < 1e-6);
  (e) reference-keyed checkpoint -> CUDA module -> forward vs the reference's golden output (f4), and
      `aggregate_tokens` on CUDA against the golden of the reference method (f3).
Tolerances as in tests/test_moe_gpu.py (TIGHT mode: oracle evaluated at bf16-representable operands)."""
import numpy as np
import pytest
import torch

import medmoe_b200
from medmoe_b200 import checkpoint, ops, text as mmtext
from oracle import moe_oracle as mo
from tests.test_moe_gpu import ACT_TOL, TIGHT, _grad_ok, _module_from
from tests.util import GOLDEN, golden_params, load_golden, rel_err

pytestmark = pytest.mark.gpu

HID, D = [96, 192, 384, 768], 768


def _bf16_params(K, seed):
    params = mo.init_params(K, HID, D, D, seed=seed)
    return {k: (v.to(torch.bfloat16).float() if (".proj_convs." in k or ".attn_proj.0." in k) and k.endswith("weight") else v)
            for k, v in params.items()}


def test_cfg2_full_batch_subset_vs_oracle():
    K, Ps, B, NSUB = 4, [3136, 784, 196, 49], 256, 8
    params = _bf16_params(K, seed=101)
    moe = _module_from(params, K, HID, D)
    g = torch.Generator().manual_seed(102)
    feats = [torch.randn(B, p, d, generator=g).to(torch.bfloat16) for p, d in zip(Ps, HID)]
    sw = torch.randn(B, D, generator=g)
    sub = torch.randperm(B, generator=g)[:NSUB]
    cg = torch.randn(B, D, generator=g)
    cl_sub = torch.randn(NSUB, D, 56, 56, generator=g) / 3136        # local cotangent: the 8 checked images only, zero elsewhere
    fg = [f.cuda().requires_grad_(True) for f in feats]
    sg = sw.cuda().requires_grad_(True)
    gf, lf, probs = moe(fg, sg)
    cl = torch.zeros(B, D, 56, 56, device="cuda", dtype=torch.bfloat16)
    cl[sub.cuda()] = cl_sub.cuda().to(torch.bfloat16)
    ((gf * cg.cuda()).sum() + (lf.float() * cl.float()).sum()).backward()

    pr = {k: v.clone() for k, v in params.items()}
    fr = [f[sub].float().requires_grad_(True) for f in feats]
    (gf_r, lf_r, probs_r), idx = mo.moe_forward_sparse(pr, fr, sw[sub])
    ((gf_r * cg[sub]).sum() + (lf_r * cl_sub.to(torch.bfloat16).float()).sum()).backward()
    assert torch.equal(moe.last_top_expert[sub.cuda(), 0].long().cpu(), idx[:, 0])
    assert (probs[sub.cuda()].cpu() - probs_r).abs().max().item() < 1e-5
    assert rel_err(gf[sub.cuda()].float().cpu(), gf_r) < ACT_TOL
    assert rel_err(lf[sub.cuda()].float().cpu(), lf_r) < 1.5e-2      # + bf16 rounding of the returned activations
    for s in range(4):
        _grad_ok(fg[s].grad[sub.cuda()].float().cpu(), fr[s].grad, TIGHT, f"d_feat{s}")
    assert moe.near_tie_count() == 0


def test_cfg4_geometry_top2_k8_vs_generalised_oracle():
    """BASELINE config 4 as it is named: 8 experts, top-2, 384^2 input = 9216/2304/576/144 tokens (extension: parity is against
    the generalised oracle)."""
    K, Ps, B = 8, [9216, 2304, 576, 144], 3
    params = _bf16_params(K, seed=111)
    torch.manual_seed(112)
    feats = [torch.randn(B, p, d).to(torch.bfloat16).float() for p, d in zip(Ps, HID)]
    sw, cg = torch.randn(B, D), torch.randn(B, D)
    cl = torch.randn(B, D, 96, 96) / 9216
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    fr = [f.clone().requires_grad_(True) for f in feats]
    sr = sw.clone().requires_grad_(True)
    (gf, lf, probs), idx = mo.moe_forward_sparse(pr, fr, sr, topk=2)
    ((gf * cg).sum() + (lf * cl).sum()).backward()

    moe = _module_from(params, K, HID, D, topk=2)
    fg = [f.cuda().requires_grad_(True) for f in feats]
    sg = sw.cuda().requires_grad_(True)
    gf2, lf2, probs2 = moe(fg, sg)
    assert lf2.shape == (B, D, 96, 96)
    assert torch.equal(moe.last_top_expert.long().cpu(), idx)
    assert rel_err(gf2.cpu(), gf) < ACT_TOL and rel_err(lf2.cpu(), lf) < ACT_TOL
    ((gf2 * cg.cuda()).sum() + (lf2 * cl.cuda()).sum()).backward()
    for s in range(4):
        _grad_ok(fg[s].grad.cpu(), fr[s].grad, TIGHT, f"d_feat{s}")
    used = set(idx.flatten().tolist())
    for k, p in moe.named_parameters():
        if k.startswith("experts.") and int(k.split(".")[1]) in used and not k.endswith("attn_proj.2.bias"):
            _grad_ok(p.grad.cpu(), pr[k].grad, TIGHT, k, key="attn0" if ".attn_proj.0." in k else "grad")
        elif k.startswith("experts.") and int(k.split(".")[1]) not in used:
            assert float(p.grad.abs().max()) == 0.0
    _grad_ok(sg.grad.cpu(), sr.grad, dict(grad=2e-2, cos=0.995), "d_swin_feat")


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_non_finite_router_input_goes_to_expert_zero_without_fault(bad):
    K, Ps, B = 4, [64, 16, 4, 1], 5
    moe = _module_from(mo.init_params(K, HID, D, D, seed=3), K, HID, D)
    torch.manual_seed(0)
    feats = [torch.randn(B, p, d, device="cuda") for p, d in zip(Ps, HID)]
    sw = torch.randn(B, D, device="cuda")
    sw[2, 17] = bad
    sw[4, :] = bad
    fg = [f.clone().requires_grad_(True) for f in feats]
    gf, lf, probs = moe(fg, sw)
    torch.cuda.synchronize()
    ref_idx = torch.argmax(torch.softmax(moe.router(sw), -1), -1)        # torch: NaN is the maximum, first one wins
    assert torch.equal(moe.last_top_expert[:, 0].long(), ref_idx)
    assert int(moe.last_top_expert[2, 0]) == 0 and int(moe.last_top_expert[4, 0]) == 0
    assert torch.isnan(probs[2]).all() and torch.isnan(probs[4]).all() and torch.isfinite(probs[[0, 1, 3]]).all()
    assert torch.isfinite(gf).all() and torch.isfinite(lf).all()         # expert 0 ran on finite stage features
    gf.sum().backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(f.grad).all() for f in fg)


def test_dispatch_build_clamps_out_of_range_expert_ids():
    from medmoe_b200 import plan as mmplan
    B, K, Ps = 9, 3, [16, 4, 1, 1]
    item_expert = torch.tensor([0, 2, -1, 1, 7, 2, 0, 1 << 20, 1], dtype=torch.int32, device="cuda")
    layout = mmplan.make_layout(B, 1, K, Ps)
    plan = mmplan.build_plan(item_expert, layout)
    torch.cuda.synchronize()
    clean = [e if 0 <= e < K else 0 for e in item_expert.cpu().tolist()]
    ref = mmplan.reference_plan(clean, layout)
    assert plan.counts.cpu().tolist() == ref["counts"] and plan.perm.cpu().tolist() == ref["perm"]
    assert plan.slot_row.cpu().tolist() == ref["slot_row"]


def test_router_near_tie_report():
    K, Ps, B = 4, [64, 16, 4, 1], 6
    params = mo.init_params(K, HID, D, D, seed=5)
    params["router.2.weight"][1] = params["router.2.weight"][3]          # experts 1 and 3 always get the same logit
    params["router.2.bias"][1] = params["router.2.bias"][3]
    moe = _module_from(params, K, HID, D)
    torch.manual_seed(1)
    feats = [torch.randn(B, p, d, device="cuda") for p, d in zip(Ps, HID)]
    sw = torch.randn(B, D, device="cuda")
    with torch.no_grad():
        _, _, probs = moe(feats, sw)
    top2 = probs.topk(2, dim=-1).values
    expect = (top2[:, 0] - top2[:, 1] < ops.NEAR_TIE_TOL).int()
    assert torch.equal(moe.last_near_tie, expect)
    tied_top = torch.isin(torch.argmax(probs, -1), torch.tensor([1, 3], device="cuda"))
    assert torch.equal(expect.bool(), tied_top) and moe.near_tie_count() == int(tied_top.sum())
    # ties break to the lowest index, like torch.argmax
    assert (moe.last_top_expert[tied_top, 0] == 1).all()


def test_reference_checkpoint_into_cuda_module_matches_reference_output(tmp_path):
    """f4: a Lightning-style checkpoint with the reference's key prefix -> medmoe_b200.MoE on the GPU -> forward equals the
    reference's own output (golden moe_k6, generated by the unmodified reference MoE)."""
    g = load_golden("moe_k6")
    params = golden_params(g)
    sd = {"model.image_encoder.model.moe." + k: v for k, v in params.items()}
    sd["model.text_encoder.model.embeddings.word_embeddings.weight"] = torch.zeros(3, 3)
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": sd, "epoch": 1, "global_step": 10}, path)
    K = g["probs"].shape[1]
    hidden = [g[f"feat{s}"].shape[2] for s in range(4)]
    torch.manual_seed(99)
    moe = medmoe_b200.MoE(num_experts=K, hidden_dims=hidden, output_dim=g["global_feat"].shape[1],
                          router_input_dim=g["swin_feat"].shape[1]).cuda()
    res = checkpoint.load_reference_checkpoint(moe, str(path))
    assert not res.missing_keys and not res.unexpected_keys
    with torch.no_grad():
        gf, lf, probs = moe([g[f"feat{s}"].cuda() for s in range(4)], g["swin_feat"].cuda())
    assert torch.equal(torch.argmax(probs, -1).cpu(), g["top_expert"])
    assert (probs.cpu() - g["probs"]).abs().max().item() < 1e-5
    assert rel_err(gf.cpu(), g["global_feat"]) < ACT_TOL and rel_err(lf.cpu(), g["local_feat"]) < ACT_TOL


def test_aggregate_tokens_on_cuda_matches_reference_golden():
    """f3: word-piece aggregation as a segmented sum on the GPU (no per-token host round trips) against the output of the
    reference's `BertEncoder.aggregate_tokens` (tests/golden/make_golden_text.py)."""
    z = np.load(f"{GOLDEN}/text_aggregate.npz")
    vocab = [str(w) for w in z["vocab"]]
    table = mmtext.continuation_table(dict(enumerate(vocab))).cuda()
    emb = torch.from_numpy(z["embeddings"]).cuda()
    ids = torch.from_numpy(z["caption_ids"]).cuda()
    got, n_words = mmtext.aggregate_tokens(emb, ids, table, int(z["sep_id"]))
    assert got.is_cuda and torch.allclose(got.cpu(), torch.from_numpy(z["aggregated"]), atol=1e-6)
    assert n_words.cpu().tolist() == z["n_words"].tolist()
    # the aggregated word embeddings feed the local loss as [B, D, L]: sum of the last layers, like text_encoder.py:119-127
    words = got.sum(1).permute(0, 2, 1).contiguous()
    assert words.shape == (emb.shape[0], emb.shape[3], emb.shape[2])


def test_attn_proj0_gradient_against_storage_aware_oracle():
    """the 1e-1 bound on the attn_proj.0 (W1, b1) gradients.  The whole gap to the fp32 oracle is the ReLU
    gate of the attention hidden layer being decided on bf16-STORED Z (and bf16-stored Y feeding it): against the oracle that
    rounds at the same two storage points (`expert_forward_storage_aware`; rounding is straight-through for autograd) the same
    gradients agree ten times tighter.  Both errors are printed (`pytest -s`) and recorded in DESIGN.md §2."""
    K, Ps, B = 4, [3136, 784, 196, 49], 6
    params = _bf16_params(K, seed=131)
    torch.manual_seed(132)
    feats = [torch.randn(B, p, d).to(torch.bfloat16).float() for p, d in zip(Ps, HID)]
    sw, cg = torch.randn(B, D), torch.randn(B, D)
    cl = torch.randn(B, D, 56, 56) / 3136
    ref = {}
    for name, storage in (("fp32", None), ("stored", torch.bfloat16)):
        pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        (gf, lf, _), idx = mo.moe_forward_sparse(pr, [f.clone() for f in feats], sw.clone(), storage=storage)
        ((gf * cg).sum() + (lf * cl).sum()).backward()
        ref[name] = {k: v.grad for k, v in pr.items()}
    moe = _module_from(params, K, HID, D)
    gf2, lf2, _ = moe([f.cuda() for f in feats], sw.cuda())
    ((gf2 * cg.cuda()).sum() + (lf2 * cl.cuda()).sum()).backward()
    used = set(idx.flatten().tolist())
    worst = {"fp32": 0.0, "stored": 0.0}
    for k, p in moe.named_parameters():
        if ".attn_proj.0." not in k or int(k.split(".")[1]) not in used:
            continue
        for name in worst:
            worst[name] = max(worst[name], rel_err(p.grad.cpu(), ref[name][k]))
    print(f"\nattn_proj.0 gradient, worst norm-wise error over experts: vs fp32 oracle {worst['fp32']:.4f}, "
          f"vs storage-aware oracle {worst['stored']:.4f}")
    assert worst["fp32"] < TIGHT["attn0"]
    assert worst["stored"] < 1e-2      # measured 1.5e-3 (vs 1.9e-2 against the fp32 oracle)


def test_finest_scale_addressed_in_image_order_is_bit_identical_to_the_sorted_copy():
    """Round 2: with top-1 routing and bf16 features the finest stage feature is not permuted — the back-to-back kernel, the conv
    weight-gradient GEMM and the input-gradient GEMM address the image-order tensors through a 64-row group map
    (mm_dispatch_group_map, *_gather / *_scatter entry points).  Same operands, same order of accumulation: outputs and every
    gradient must equal the sorted-copy path bit for bit."""
    K, Ps, B = 4, [3136, 784, 196, 49], 7
    params = _bf16_params(K, seed=141)
    torch.manual_seed(142)
    feats = [torch.randn(B, p, d).to(torch.bfloat16) for p, d in zip(Ps, HID)]
    sw, cg = torch.randn(B, D), torch.randn(B, D)
    cl = (torch.randn(B, D, 56, 56) / 3136).to(torch.bfloat16)
    results = []
    for direct in (True, False):
        ops.USE_DIRECT_FINEST = direct
        try:
            moe = _module_from(params, K, HID, D)
            fg = [f.cuda().requires_grad_(True) for f in feats]
            gf, lf, probs = moe(fg, sw.cuda())
            assert moe.last_direct_finest == direct
            ((gf * cg.cuda()).sum() + (lf.float() * cl.cuda().float()).sum()).backward()
            results.append((gf.detach().clone(), lf.detach().clone(), [f.grad.clone() for f in fg],
                            {k: p.grad.clone() for k, p in moe.named_parameters() if p.grad is not None}))
        finally:
            ops.USE_DIRECT_FINEST = True
    (gf_a, lf_a, df_a, dp_a), (gf_b, lf_b, df_b, dp_b) = results
    assert torch.equal(gf_a, gf_b) and torch.equal(lf_a, lf_b)
    for s in range(4):
        assert torch.equal(df_a[s], df_b[s]), f"d_feat{s}"
    for k in dp_a:
        # weight gradients are fp32 red.add reductions over row chunks: the order of the atomic adds is not fixed
        assert rel_err(dp_a[k].cpu(), dp_b[k].cpu()) < 1e-5 or float(dp_b[k].abs().max()) == 0.0, k
