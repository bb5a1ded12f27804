"""Static row layout + per-step dispatch plan of the expert-sorted activation buffers.

Layout: every activation that is indexed by "token row" lives in one row
space made of S scale regions.  Region s holds, for every expert e in order, the
`count[e] * P_s` native rows of the images routed to e, padded to a multiple of 256 rows so
that a 128-row GEMM tile never straddles two experts and neither does a CTA pair's two tiles.  Region capacities are static
(`n_items * P_s` rounded up + 256 rows per expert), so buffers and launch grids do not
depend on the routing and no host<->device sync is needed; only the small int tables
written by `mm_dispatch_build` change per step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import torch

from . import _lib

TILE_M = 128
SEG_ALIGN = 256      # csrc/mm_common.cuh: expert segments start on multiples of 256 rows (two tiles), see there


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class RowLayout:
    n_images: int
    topk: int
    num_experts: int
    P: List[int]              # native rows per image per scale
    region_rows: List[int]    # capacity of each region (multiple of 128)
    region_base: List[int]    # first global row of each region
    region_tiles: List[int]
    tile_base: List[int]      # first global tile of each region
    chunk_tiles: List[int]    # tiles per wgrad chunk
    chunk_cap: List[int]
    chunk_base: List[int]
    total_rows: int
    total_tiles: int
    total_chunks: int

    @property
    def n_items(self) -> int:
        return self.n_images * self.topk

    @property
    def S(self) -> int:
        return len(self.P)


def make_layout(n_images: int, topk: int, num_experts: int, P: Sequence[int], target_chunks: int = 96) -> RowLayout:
    n_items = n_images * topk
    region_rows, region_base, region_tiles, tile_base = [], [], [], []
    chunk_tiles, chunk_cap, chunk_base = [], [], []
    row = tile = chunk = 0
    for p in P:
        rows = _round_up(n_items * p, SEG_ALIGN) + num_experts * SEG_ALIGN
        tiles = rows // TILE_M
        # wgrad chunks: enough of them to fill the SMs, few enough that the fp32 red.add traffic of their partial
        # results stays small (small regions get proportionally fewer chunks)
        want = target_chunks if tiles >= 1024 else max(8, target_chunks // 4)
        g = 1
        while g * 2 <= max(1, tiles // want) and g < 64:
            g *= 2
        cap = (tiles + g - 1) // g + num_experts
        region_rows.append(rows); region_base.append(row); region_tiles.append(tiles); tile_base.append(tile)
        chunk_tiles.append(g); chunk_cap.append(cap); chunk_base.append(chunk)
        row += rows; tile += tiles; chunk += cap
    return RowLayout(n_images, topk, num_experts, list(P), region_rows, region_base, region_tiles, tile_base,
                     chunk_tiles, chunk_cap, chunk_base, row, tile, chunk)


@dataclass
class DispatchPlan:
    layout: RowLayout
    counts: torch.Tensor       # [K] int32
    offsets: torch.Tensor      # [K+1]
    perm: torch.Tensor         # [n_items] slot -> item
    inv_perm: torch.Tensor     # [n_items] item -> slot
    slot_expert: torch.Tensor  # [n_items]
    seg_start: torch.Tensor    # [S, K]
    slot_row: torch.Tensor     # [S, n_items]
    tile_info: torch.Tensor    # [total_tiles, 2]
    chunks: torch.Tensor       # [total_chunks, 4]


def build_plan(item_expert: torch.Tensor, layout: RowLayout) -> DispatchPlan:
    """item_expert: int32 [n_items] on the GPU (router top-k indices, row-major [B, topk])."""
    assert item_expert.dtype == torch.int32 and item_expert.is_cuda and item_expert.is_contiguous()
    dev = item_expert.device
    n, K, S = layout.n_items, layout.num_experts, layout.S
    assert item_expert.numel() == n
    i32 = dict(dtype=torch.int32, device=dev)
    plan = DispatchPlan(
        layout,
        torch.empty(K, **i32), torch.empty(K + 1, **i32), torch.empty(n, **i32), torch.empty(n, **i32),
        torch.empty(n, **i32), torch.empty(S, K, **i32), torch.empty(S, n, **i32),
        torch.empty(layout.total_tiles, 2, **i32), torch.empty(layout.total_chunks, 4, **i32))
    _lib.call("mm_dispatch_build", _lib.ptr(item_expert), n, K, S,
              _lib.host_i32(layout.P), _lib.host_i32(layout.region_base), _lib.host_i32(layout.region_tiles),
              _lib.host_i32(layout.chunk_base), _lib.host_i32(layout.chunk_cap), _lib.host_i32(layout.chunk_tiles),
              _lib.ptr(plan.counts), _lib.ptr(plan.offsets), _lib.ptr(plan.perm), _lib.ptr(plan.inv_perm),
              _lib.ptr(plan.slot_expert), _lib.ptr(plan.seg_start), _lib.ptr(plan.slot_row),
              _lib.ptr(plan.tile_info), _lib.ptr(plan.chunks), _lib.stream_ptr())
    return plan


def reference_plan(item_expert: Sequence[int], layout: RowLayout) -> dict:
    """Pure-Python restatement of mm_dispatch_build (host logic; used by the CPU tests)."""
    K, S, n = layout.num_experts, layout.S, layout.n_items
    counts = [0] * K
    for e in item_expert:
        counts[e] += 1
    offsets = [0]
    for e in range(K):
        offsets.append(offsets[-1] + counts[e])
    seg_start = [[0] * K for _ in range(S)]
    for s in range(S):
        r = layout.region_base[s]
        for e in range(K):
            seg_start[s][e] = r
            r += _round_up(counts[e] * layout.P[s], SEG_ALIGN)
    perm, inv_perm, slot_expert = [0] * n, [0] * n, [0] * n
    slot_row = [[0] * n for _ in range(S)]
    seen = [0] * K
    for item, e in enumerate(item_expert):
        rank = seen[e]; seen[e] += 1
        slot = offsets[e] + rank
        perm[slot] = item; inv_perm[item] = slot; slot_expert[slot] = e
        for s in range(S):
            slot_row[s][slot] = seg_start[s][e] + rank * layout.P[s]
    tile_info = []
    for s in range(S):
        for t in range(layout.region_tiles[s]):
            row = layout.region_base[s] + t * TILE_M
            info = (-1, 0)
            for e in range(K):
                rows_e = counts[e] * layout.P[s]
                if rows_e > 0 and seg_start[s][e] <= row < seg_start[s][e] + rows_e:
                    info = (e, min(TILE_M, seg_start[s][e] + rows_e - row))
            tile_info.append(info)
    chunks = []
    for s in range(S):
        g = layout.chunk_tiles[s]
        region = []
        for e in range(K):
            nt = _round_up(counts[e] * layout.P[s], TILE_M) // TILE_M
            first = seg_start[s][e] // TILE_M
            for c in range(0, nt, g):
                region.append((e, first + c, min(g, nt - c), s))
        assert len(region) <= layout.chunk_cap[s]
        region += [(0, 0, 0, s)] * (layout.chunk_cap[s] - len(region))
        chunks += region
    return dict(counts=counts, offsets=offsets, perm=perm, inv_perm=inv_perm, slot_expert=slot_expert,
                seg_start=seg_start, slot_row=slot_row, tile_info=tile_info, chunks=chunks)
