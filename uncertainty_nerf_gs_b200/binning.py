"""Tile binning of projected 2-D Gaussians on the device (setup step, NOT on the timed hot path).

BASELINE.json's splat configuration composites over *pre-binned* per-tile lists, so binning happens once
outside the timed region.  ``bin_gaussians`` builds those lists with the CUDA kernels of ``csrc/binning.cu``
(count -> scan -> expand -> two stable radix sorts (depth, then tile) -> ranges), following the published
gsplat 0.1.x scheme the reference relies on through ``gsplat.rasterize_gaussians`` (reference
activesplatfacto_model.py:260-273; gsplat itself is not vendored).  The torch restatement used to check it
lives in ``oracle/splat.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from . import _lib, ops

Tensor = torch.Tensor
TILE = 16


def tile_grid(height: int, width: int) -> Tuple[int, int]:
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE  # tiles_x, tiles_y


def bin_gaussians(xys: Tensor, depths: Tensor, radii: Tensor, height: int, width: int) -> Tuple[Tensor, Tensor]:
    """Returns ``(gaussian_ids [I] int32 sorted by (tile, depth), tile_bins [tiles, 2] int32)`` (CUDA tensors)."""
    lib = _lib.load()
    xys = ops._dev_f32(xys, "xys")
    depths = ops._dev_f32(depths.reshape(-1), "depths")
    if radii.dtype != torch.int32 or not radii.is_cuda:
        raise TypeError("radii must be a CUDA int32 tensor")
    radii = radii.contiguous()
    g = xys.shape[0]
    dev = xys.device
    tiles_x, tiles_y = tile_grid(height, width)
    offsets = torch.empty(g + 1, dtype=torch.int64, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    ws = ops._workspace(lib.ub_bin_count_workspace_bytes(g), dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ub_bin_count(xys.data_ptr(), radii.data_ptr(), g, height, width, offsets.data_ptr(),
                                    total.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream()))
        n = int(total.item())          # the one host synchronisation of binning: the output size
        ids = torch.empty(n, dtype=torch.int32, device=dev)
        bins = torch.empty(tiles_x * tiles_y, 2, dtype=torch.int32, device=dev)
        ws2 = ops._workspace(lib.ub_bin_gaussians_workspace_bytes(n), dev)
        _lib.check(lib.ub_bin_gaussians(xys.data_ptr(), depths.data_ptr(), radii.data_ptr(), g, height, width,
                                        offsets.data_ptr(), n, ids.data_ptr(), bins.data_ptr(), ws2.data_ptr(),
                                        ws2.numel(), ops._stream()))
    ops._count(4 + 4 + 24)
    return ids, bins
