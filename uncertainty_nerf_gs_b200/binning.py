"""Tile binning of projected 2-D Gaussians (setup step, NOT on the timed hot path).

BASELINE.json's splat configuration composites over *pre-binned* per-tile lists, so binning happens
once outside the timed region.  This module builds those lists with plain torch ops (any device),
following the published gsplat 0.1.x scheme the reference relies on through
``gsplat.rasterize_gaussians`` (reference activesplatfacto_model.py:260-273; gsplat itself is not
vendored): tile rectangle from centre +- radius, one (tile << 32 | depth bits) int64 key per
(Gaussian, tile) intersection, stable sort, per-tile [start, end) ranges.  A CUDA version reusing the
radix sort is the "next" row (f2) of SURVEY.md section 8.
"""
from __future__ import annotations

from typing import Tuple

import torch

Tensor = torch.Tensor
TILE = 16


def tile_grid(height: int, width: int) -> Tuple[int, int]:
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE  # tiles_x, tiles_y


def bin_gaussians(xys: Tensor, depths: Tensor, radii: Tensor, height: int, width: int) -> Tuple[Tensor, Tensor]:
    """Returns ``(gaussian_ids [I] int32 sorted by (tile, depth), tile_bins [tiles, 2] int32)``."""
    tiles_x, tiles_y = tile_grid(height, width)
    dev = xys.device
    r = radii.to(torch.float32)
    cx, cy = xys[:, 0] / TILE, xys[:, 1] / TILE
    tr = r / TILE
    # (int) casts truncate toward zero, as the C code does
    x0 = torch.clamp(torch.trunc(cx - tr).long(), 0, tiles_x)
    x1 = torch.clamp(torch.trunc(cx + tr + 1).long(), 0, tiles_x)
    y0 = torch.clamp(torch.trunc(cy - tr).long(), 0, tiles_y)
    y1 = torch.clamp(torch.trunc(cy + tr + 1).long(), 0, tiles_y)
    nx, ny = (x1 - x0).clamp(min=0), (y1 - y0).clamp(min=0)
    hits = torch.where(radii > 0, nx * ny, torch.zeros_like(nx))
    total = int(hits.sum().item())
    ids = torch.repeat_interleave(torch.arange(xys.shape[0], device=dev), hits)
    first = torch.cumsum(hits, 0) - hits
    local = torch.arange(total, device=dev) - first[ids]
    w = nx[ids].clamp(min=1)
    ty = y0[ids] + local // w
    tx = x0[ids] + local % w
    tile = ty * tiles_x + tx
    depth_bits = depths.to(torch.float32).contiguous().view(torch.int32).long()[ids] & 0xFFFFFFFF
    keys = (tile << 32) | depth_bits
    order = torch.sort(keys, stable=True).indices
    sorted_tiles = tile[order]
    gaussian_ids = ids[order].to(torch.int32)
    bounds = torch.searchsorted(sorted_tiles, torch.arange(tiles_x * tiles_y + 1, device=dev))
    tile_bins = torch.stack([bounds[:-1], bounds[1:]], dim=1).to(torch.int32)
    return gaussian_ids.contiguous(), tile_bins.contiguous()
