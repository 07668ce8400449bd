"""Projection, view-dependent colours and tile binning of Gaussians on the device (the producers of the tile
compositor; setup steps, NOT on the timed hot path).

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


def project_gaussians(means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor, viewmat: Tensor, fx: float,
                      fy: float, cx: float, cy: float, img_height: int, img_width: int, block_width: int = TILE,
                      clip_thresh: float = 0.01) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Drop-in for ``gsplat.project_gaussians`` as called at activesplatfacto_model.py:221-234 (forward only):
    ``(xys, depths, radii, conics, compensation, num_tiles_hit, cov3d)``.  ``viewmat``: CUDA ``[3, 4]`` (or
    ``[4, 4]``) world-to-camera matrix -- it stays on the device, no host synchronisation."""
    if block_width != TILE:
        raise ValueError(f"block_width must be {TILE}")
    lib = _lib.load()
    means3d, scales, quats = ops._dev_f32(means3d, "means3d"), ops._dev_f32(scales, "scales"), ops._dev_f32(quats, "quats")
    vm = ops._dev_f32(viewmat.reshape(-1, 4)[:3].contiguous(), "viewmat")
    g = means3d.shape[0]
    if means3d.shape != (g, 3) or scales.shape != (g, 3) or quats.shape != (g, 4):
        raise ValueError("means3d [G,3], scales [G,3], quats [G,4] expected")
    dev = means3d.device
    xys, depths = torch.empty(g, 2, device=dev), torch.empty(g, device=dev)
    radii = torch.empty(g, dtype=torch.int32, device=dev)
    conics, comp = torch.empty(g, 3, device=dev), torch.empty(g, device=dev)
    tiles = torch.empty(g, dtype=torch.int32, device=dev)
    cov3d = torch.empty(g, 6, device=dev)
    with ops._guard(dev):
        _lib.check(lib.ub_project_gaussians(means3d.data_ptr(), scales.data_ptr(), float(glob_scale), quats.data_ptr(),
                                            vm.data_ptr(), float(fx), float(fy), float(cx), float(cy), img_height,
                                            img_width, float(clip_thresh), g, xys.data_ptr(), depths.data_ptr(),
                                            radii.data_ptr(), conics.data_ptr(), comp.data_ptr(), tiles.data_ptr(),
                                            cov3d.data_ptr(), ops._stream()))
    ops._count(1)
    return xys, depths, radii, conics, comp, tiles, cov3d


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    """Drop-in for ``gsplat.spherical_harmonics`` (activesplatfacto_model.py:245, forward only): ``coeffs [G, K, 3]``
    with ``K = (degree + 1)^2``, unnormalised ``viewdirs [G, 3]`` -> colours ``[G, 3]``."""
    lib = _lib.load()
    coeffs, viewdirs = ops._dev_f32(coeffs, "coeffs"), ops._dev_f32(viewdirs, "viewdirs")
    g, k = coeffs.shape[0], coeffs.shape[1]
    degree = {1: 0, 4: 1, 9: 2, 16: 3}.get(k)
    if degree is None or coeffs.shape[2] != 3:
        raise ValueError("coeffs must be [G, (degree+1)^2, 3] with degree <= 3")
    out = torch.empty(g, 3, device=coeffs.device)
    with ops._guard(coeffs.device):
        _lib.check(lib.ub_spherical_harmonics(degree, int(degrees_to_use), viewdirs.data_ptr(), coeffs.data_ptr(), g,
                                              out.data_ptr(), ops._stream()))
    ops._count(1)
    return out


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
    with ops._guard(dev):
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
