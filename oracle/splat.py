"""Oracle: tile-based front-to-back alpha compositing of 2-D Gaussians and the four
active-splatfacto rasterisation passes (torch, CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the per-pixel loop restates
the published algorithm of gsplat 0.1.11 ``rasterize_forward`` (the version
``/root/reference/README.md:30`` pins; not vendored, not installable here); the pass structure and the
post-processing follow the reference's call sites,
``nerfuncertainty/models/activesplatfacto/activesplatfacto_model.py:260-367``.
One deliberate difference to gsplat's CUDA kernel: ``exp`` is the accurate float32 exponential, not
the fast ``__expf`` intrinsic.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor
TILE = 16


# ---- tile binning (published gsplat 0.1.x scheme; the reference reaches it through rasterize_gaussians,
# activesplatfacto_model.py:260-273): tile rectangle from centre +- radius, one (tile << 32 | depth bits) int64
# key per (Gaussian, tile) intersection, stable sort, per-tile [start, end) ranges ----
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



def rasterize(xys: Tensor, conics: Tensor, opacities: Tensor, colors: Tensor, gaussian_ids: Tensor,
              tile_bins: Tensor, height: int, width: int, background: Tensor) -> Tuple[Tensor, Tensor]:
    """``out [H, W, C] = sum_i c_i alpha_i T_i + T_final * background``, ``alpha [H, W] = 1 - T_final``.
    Vectorised over the pixels of a tile, sequential over the tile's depth-sorted Gaussian list."""
    ch = colors.shape[1]
    out = torch.zeros(height, width, ch)
    final_t = torch.ones(height, width)
    tiles_x = (width + TILE - 1) // TILE
    tiles_y = (height + TILE - 1) // TILE
    opac = opacities.reshape(-1)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            lo, hi = (int(v) for v in tile_bins[ty * tiles_x + tx])
            i0, j0 = ty * TILE, tx * TILE
            i1, j1 = min(i0 + TILE, height), min(j0 + TILE, width)
            ii, jj = torch.meshgrid(torch.arange(i0, i1), torch.arange(j0, j1), indexing="ij")
            px = jj.float() + 0.5
            py = ii.float() + 0.5
            T = torch.ones_like(px)
            acc = torch.zeros(*px.shape, ch)
            done = torch.zeros_like(px, dtype=torch.bool)
            for idx in range(lo, hi):
                if bool(done.all()):
                    break
                g = int(gaussian_ids[idx])
                dx = xys[g, 0] - px
                dy = xys[g, 1] - py
                sigma = 0.5 * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
                alpha = torch.clamp(opac[g] * torch.exp(-sigma), max=0.999)
                skip = (sigma < 0) | (alpha < 1.0 / 255.0)
                next_t = T * (1.0 - alpha)
                stop = (~done) & (~skip) & (next_t <= 1e-4)
                done = done | stop
                take = (~done) & (~skip)
                vis = alpha * T
                acc = torch.where(take[..., None], acc + colors[g] * vis[..., None], acc)
                T = torch.where(take, next_t, T)
            out[i0:i1, j0:j1] = acc + T[..., None] * background
            final_t[i0:i1, j0:j1] = T
    return out, 1.0 - final_t


def active_splatfacto_outputs(xys: Tensor, depths: Tensor, conics: Tensor, opacities: Tensor, rgbs: Tensor,
                              betas: Tensor, gaussian_ids: Tensor, tile_bins: Tensor, height: int, width: int,
                              background: Tensor) -> Dict[str, Tensor]:
    """The rasterisation block of ``ActiveSplatfactoModel.get_outputs``
    (``activesplatfacto_model.py:260-367``) as four separate 3-channel passes, like the reference."""
    args = (gaussian_ids, tile_bins, height, width)
    zeros3 = torch.zeros(3)
    rgb, alpha = rasterize(xys, conics, opacities, rgbs, *args, background)
    alpha = alpha[..., None]
    rgb = torch.clamp(rgb, max=1.0)
    unc_im = rasterize(xys, conics, opacities, betas.reshape(-1, 1).repeat(1, 3), *args, zeros3)[0][..., 0:1]
    depth_im = rasterize(xys, conics, opacities, depths[:, None].repeat(1, 3), *args, zeros3)[0][..., 0:1]
    depth_im = torch.where(alpha > 0, depth_im / alpha, depth_im.detach().max())
    xy_to_pix = torch.floor(xys).long()
    valid = (xy_to_pix[:, 0] > 0) & (xy_to_pix[:, 0] < width) & (xy_to_pix[:, 1] > 0) & (xy_to_pix[:, 1] < height)
    pv = xy_to_pix[valid]
    fetched = depth_im[pv[:, 1], pv[:, 0], 0]
    resid = depths.clone()
    resid[valid] -= fetched
    dvar_im = rasterize(xys, conics, opacities, (resid[:, None] ** 2).repeat(1, 3), *args, zeros3)[0][..., 0:1]
    dvar_im = torch.where(alpha > 0, dvar_im / alpha, dvar_im.detach().max())
    return {
        "rgb": rgb, "depth": depth_im, "accumulation": alpha, "background": background,
        "uncertainty": unc_im, "rgb_var": unc_im ** 2, "rgb_std": unc_im,
        "depth_var": dvar_im, "depth_std": dvar_im.sqrt(),
    }
