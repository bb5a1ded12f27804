"""CPU: host-side dispatch logic (medmoe_b200.plan) — static row layout and the pure-Python
restatement of mm_dispatch_build that the GPU test compares the kernel against bit for bit."""
import random

from medmoe_b200 import plan as mmplan


def _check(n_items, K, P, seed, topk=1):
    rnd = random.Random(seed)
    experts = [rnd.randrange(K) for _ in range(n_items * topk)]
    lay = mmplan.make_layout(n_items, topk, K, P)
    ref = mmplan.reference_plan(experts, lay)
    n = lay.n_items
    assert sorted(ref["perm"]) == list(range(n)) and all(ref["inv_perm"][ref["perm"][s]] == s for s in range(n))
    # stable counting sort: slots of one expert keep item order
    for e in range(K):
        items = [ref["perm"][s] for s in range(ref["offsets"][e], ref["offsets"][e + 1])]
        assert items == sorted(items) and all(experts[i] == e for i in items)
    for s in range(lay.S):
        assert lay.region_base[s] % 128 == 0 and lay.region_rows[s] % 128 == 0
        end_prev = lay.region_base[s]
        for e in range(K):
            st = ref["seg_start"][s][e]
            assert st % 128 == 0 and st >= end_prev
            end_prev = st + ref["counts"][e] * P[s]
        assert end_prev <= lay.region_base[s] + lay.region_rows[s]
        # every row of every slot is covered by exactly one owned tile entry with the right expert
        for slot in range(n):
            r0 = ref["slot_row"][s][slot]
            for r in (r0, r0 + P[s] - 1):
                e, valid = ref["tile_info"][r // 128]
                assert e == ref["slot_expert"][slot] and r % 128 < valid
    # CTA pairs (tcgen05 cta_group::2 kernels): segments start on 256-row boundaries and regions hold an even number of tiles,
    # so the tiles (2j, 2j + 1) of a region never belong to two experts
    for s in range(lay.S):
        assert lay.tile_base[s] % 2 == 0 and lay.region_tiles[s] % 2 == 0
        assert all(ref["seg_start"][s][e] % mmplan.SEG_ALIGN == 0 for e in range(K))
        for t in range(lay.tile_base[s], lay.tile_base[s] + lay.region_tiles[s], 2):
            e0, e1 = ref["tile_info"][t][0], ref["tile_info"][t + 1][0]
            assert e0 == e1 or e0 < 0 or e1 < 0, (s, t, e0, e1)
    # wgrad chunks: disjoint, cover every owned tile once, never mix experts
    covered = {}
    for (e, first, cnt, s) in ref["chunks"]:
        for t in range(first, first + cnt):
            assert t not in covered
            covered[t] = e
    owned = {t: e for t, (e, v) in enumerate(ref["tile_info"]) if e >= 0}
    assert covered == owned


def test_reference_plan_invariants():
    _check(7, 4, [49, 13, 5, 1], 0)
    _check(64, 6, [196, 49, 16, 4], 1)
    _check(300, 8, [3, 2, 1, 1], 2)
    _check(5, 3, [300, 75, 19, 5], 3)
    _check(16, 8, [64, 16, 4, 1], 4, topk=2)
    _check(1, 1, [3136, 784, 196, 49], 5)


def test_layout_capacity_is_routing_independent():
    lay = mmplan.make_layout(256, 1, 4, [3136, 784, 196, 49])
    assert lay.total_rows == sum(lay.region_rows) and lay.total_tiles == lay.total_rows // 128
    # worst-case skew (everything to one expert) and perfectly balanced routing both fit
    for experts in ([0] * 256, [i % 4 for i in range(256)]):
        ref = mmplan.reference_plan(experts, lay)
        assert len(ref["tile_info"]) == lay.total_tiles and len(ref["chunks"]) == lay.total_chunks


def test_tensor_core_path_geometry_predicates():
    """Which token geometries take the tensor-core / rank-1 backward (pure host logic of the C-ABI, no GPU needed)."""
    from medmoe_b200 import _lib

    def sup(name, Ps, D=768):
        return bool(_lib.call(name, Ps[0], _lib.host_i32(Ps), D))

    swin224, swin384 = [3136, 784, 196, 49], [9216, 2304, 576, 144]
    for Ps in (swin224, swin384, [64, 16, 4, 1]):
        assert sup("mm_combine_bwd_global_supported", Ps) and sup("mm_combine_bwd_tc_supported", Ps)
    assert not sup("mm_combine_bwd_global_supported", [100, 30, 7, 1])        # non-integer ratios: generic kernels
    assert not sup("mm_combine_bwd_tc_supported", [96, 24, 6, 3])             # ratio 32 at the coarsest scale but P % 64 != 0
    assert sup("mm_combine_bwd_global_supported", [96, 24, 6, 3])             # the rank-1 path only needs even integer ratios
    assert not sup("mm_combine_bwd_tc_supported", swin224, D=640)             # output_dim must be one of 256/512/768/1024


def test_local_loss_tables_cover_every_tile_once():
    """Host tables of the word-patch attention loss: images as 'experts' of the grouped GEMMs (medmoe_b200/local_loss.py)."""
    import torch
    from medmoe_b200 import local_loss as ll
    for B, tpi in [(1, 1), (5, 25), (3, 72), (37, 2)]:
        img_tiles, img_chunks, n_img, all_chunks, n_all = ll._build_tables(B, tpi, torch.device("cpu"))
        ti = img_tiles.tile_info
        assert ti.shape == (B * tpi, 2) and ti.dtype == torch.int32
        assert ti[:, 0].tolist() == [t // tpi for t in range(B * tpi)] and (ti[:, 1] == 128).all()
        for chunks, n, per_image in ((img_chunks.chunks, n_img, True), (all_chunks.chunks, n_all, False)):
            assert chunks.shape == (n, 4)
            seen = []
            for e, first, cnt, _ in chunks.tolist():
                assert 0 < cnt <= 64
                tiles = list(range(first, first + cnt))
                if per_image:
                    assert all(t // tpi == e for t in tiles)          # a chunk never straddles two images
                else:
                    assert e == 0
                seen += tiles
            assert sorted(seen) == list(range(B * tpi))
    lens = ll._cap_len_tensor([3, 9, 1], 16, torch.device("cpu"))
    assert lens.tolist() == [3, 9, 1] + [0] * 13 and lens.dtype == torch.int32
