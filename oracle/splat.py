"""Oracle: tile-based front-to-back alpha compositing of 2-D Gaussians and the four
active-splatfacto rasterisation passes (torch, CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Two layers.  The pass structure and the
post-processing are the reference's (``nerfuncertainty/models/activesplatfacto/activesplatfacto_model.py:260-367``)
and are PINNED: ``tests/test_oracle_pinned.py::test_live_active_splatfacto_get_outputs`` executes that method
unmodified with this module standing in for gsplat and demands bit equality of every output key
(``tests/golden/ref_splat.npz``).  The rasteriser underneath restates the published algorithm of gsplat 0.1.11
``rasterize_forward`` / ``project_gaussians`` / ``spherical_harmonics`` (the version ``/root/reference/README.md:30``
pins; a third-party CUDA dependency, not vendored, not installable here): that layer has no reference output to
be pinned to.
``exp``: gsplat's CUDA kernel (and ours) uses the fast ``__expf`` intrinsic, which a CPU cannot reproduce;
``rasterize`` therefore takes the per-(splat, pixel) ``sigma`` / ``alpha`` values from the device through its ``probe``
argument when exact agreement of the threshold decisions is wanted (tests/test_gpu_splat_exact.py), and falls back to
torch's accurate float32 ``exp`` otherwise (decisions can then differ where a value sits within an ulp of a threshold).
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



def project_gaussians(means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor, viewmat: Tensor, fx: float,
                      fy: float, cx: float, cy: float, height: int, width: int, clip_thresh: float = 0.01
                      ) -> Dict[str, Tensor]:
    """gsplat 0.1.11 ``project_gaussians`` (published algorithm; call site activesplatfacto_model.py:221-234):
    ``xys, depths, radii, conics, compensation, num_tiles_hit, cov3d`` plus ``radius_real`` (the value before
    ``ceil``, for tests that must tolerate a one-ulp difference at an integer boundary).  ``viewmat [3,4]``."""
    f32 = torch.float32
    means3d, scales, quats, viewmat = (t.to(f32) for t in (means3d, scales, quats, viewmat))
    g = means3d.shape[0]
    Wm, tvec = viewmat[:3, :3], viewmat[:3, 3]
    p_view = means3d @ Wm.T + tvec
    keep = ~(p_view[:, 2] <= clip_thresh)
    q = quats / quats.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)
    M = R * (glob_scale * scales)[:, None, :]
    V = M @ M.transpose(1, 2)
    cov3d = torch.stack([V[:, 0, 0], V[:, 0, 1], V[:, 0, 2], V[:, 1, 1], V[:, 1, 2], V[:, 2, 2]], -1)
    lim_x, lim_y = 1.3 * (0.5 * width / fx), 1.3 * (0.5 * height / fy)
    tz = p_view[:, 2]
    tx = tz * torch.clamp(p_view[:, 0] / tz, -lim_x, lim_x)
    ty = tz * torch.clamp(p_view[:, 1] / tz, -lim_y, lim_y)
    rz = 1.0 / tz
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx * rz, zero, -fx * tx * rz * rz], -1),
                     torch.stack([zero, fy * rz, -fy * ty * rz * rz], -1)], -2)          # [G, 2, 3]
    T = J @ Wm
    cov = T @ V @ T.transpose(1, 2)
    c00, c01, c11 = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]
    det_orig = c00 * c11 - c01 * c01
    a, b, c = c00 + 0.3, c01, c11 + 0.3
    det = a * c - b * b
    comp = torch.sqrt(torch.clamp(det_orig / det, min=0.0))
    has_conic = keep & (det != 0)
    inv_det = 1.0 / det
    conics = torch.stack([c * inv_det, -b * inv_det, a * inv_det], -1)
    mid = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius_real = 3.0 * torch.sqrt(torch.maximum(mid + disc, mid - disc))
    radius = torch.ceil(radius_real)
    rw = 1.0 / (p_view[:, 2] + 1e-6)
    xys = torch.stack([p_view[:, 0] * rw * fx + cx, p_view[:, 1] * rw * fy + cy], -1)
    tiles_x, tiles_y = tile_grid(height, width)
    tc, tr = xys / TILE, radius / TILE
    x0 = torch.clamp(torch.trunc(tc[:, 0] - tr).long(), 0, tiles_x)
    x1 = torch.clamp(torch.trunc(tc[:, 0] + tr + 1).long(), 0, tiles_x)
    y0 = torch.clamp(torch.trunc(tc[:, 1] - tr).long(), 0, tiles_y)
    y1 = torch.clamp(torch.trunc(tc[:, 1] + tr + 1).long(), 0, tiles_y)
    area = torch.nan_to_num((x1 - x0) * (y1 - y0))
    vis = has_conic & (area > 0)
    z1 = lambda t, m: torch.where(m if t.dim() == 1 else m[:, None], t, torch.zeros_like(t))
    return {"xys": z1(xys, vis), "depths": z1(p_view[:, 2], vis), "radii": z1(radius, vis).to(torch.int32),
            "conics": z1(conics, has_conic), "compensation": z1(comp, vis),
            "num_tiles_hit": z1(area.to(f32), vis).to(torch.int32), "cov3d": z1(cov3d, keep),
            "radius_real": z1(radius_real, vis)}


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    """gsplat 0.1.11 ``spherical_harmonics`` (3DGS real SH basis; call site activesplatfacto_model.py:245):
    ``coeffs [G, K, 3]`` -> colours ``[G, 3]`` from the first ``(degrees_to_use + 1)^2`` bases; the direction is
    normalised inside."""
    cf = coeffs.to(torch.float32)
    col = _SH_C0 * cf[:, 0]
    if degrees_to_use < 1:
        return col
    d = viewdirs.to(torch.float32)
    d = d / torch.sqrt((d * d).sum(-1, keepdim=True))
    x, y, z = (d[:, i:i + 1] for i in range(3))
    col = col + _SH_C1 * (-y * cf[:, 1] + z * cf[:, 2] - x * cf[:, 3])
    if degrees_to_use < 2:
        return col
    xx, xy, xz, yy, yz, zz = x * x, x * y, x * z, y * y, y * z, z * z
    col = col + (_SH_C2[0] * xy * cf[:, 4] + _SH_C2[1] * yz * cf[:, 5] + _SH_C2[2] * (2 * zz - xx - yy) * cf[:, 6]
                 + _SH_C2[3] * xz * cf[:, 7] + _SH_C2[4] * (xx - yy) * cf[:, 8])
    if degrees_to_use < 3:
        return col
    return col + (_SH_C3[0] * y * (3 * xx - yy) * cf[:, 9] + _SH_C3[1] * xy * z * cf[:, 10]
                  + _SH_C3[2] * y * (4 * zz - xx - yy) * cf[:, 11] + _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * cf[:, 12]
                  + _SH_C3[4] * x * (4 * zz - xx - yy) * cf[:, 13] + _SH_C3[5] * z * (xx - yy) * cf[:, 14]
                  + _SH_C3[6] * x * (xx - 3 * yy) * cf[:, 15])


def rasterize(xys: Tensor, conics: Tensor, opacities: Tensor, colors: Tensor, gaussian_ids: Tensor,
              tile_bins: Tensor, height: int, width: int, background: Tensor, tiles=None, probe=None,
              want_counts: bool = False):
    """``out [H, W, C] = sum_i c_i alpha_i T_i + T_final * background``, ``alpha [H, W] = 1 - T_final``.
    Vectorised over the pixels of a tile, sequential over the tile's depth-sorted Gaussian list.

    ``tiles``: iterable of tile indices to rasterise (default: all; other pixels keep background / alpha 0).
    ``probe(tile_index, lo, hi) -> (sigma, alpha) [hi - lo, 16, 16]``: take the per-(splat, pixel) ``sigma`` and
    ``alpha = min(0.999, opacity * exp(-sigma))`` from there instead of computing them with torch's ``exp`` -- the
    CUDA kernels use the fast ``__expf`` intrinsic like gsplat's own kernel, which no CPU can reproduce; with the
    device's values every threshold decision below is the same float32 comparison the kernel makes, so the set of
    contributing splats of every pixel (and with it the alpha image) is reproduced exactly.
    ``want_counts``: also return the number of contributing splats per pixel ``[H, W]`` (int64)."""
    ch = colors.shape[1]
    out = torch.zeros(height, width, ch)
    final_t = torch.ones(height, width)
    counts = torch.zeros(height, width, dtype=torch.int64)
    tiles_x = (width + TILE - 1) // TILE
    tiles_y = (height + TILE - 1) // TILE
    opac = opacities.reshape(-1)
    chosen = None if tiles is None else set(int(t) for t in tiles)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            if chosen is not None and ty * tiles_x + tx not in chosen:
                out[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE] = background
                continue
            lo, hi = (int(v) for v in tile_bins[ty * tiles_x + tx])
            probed = probe(ty * tiles_x + tx, lo, hi) if probe is not None and hi > lo else None
            i0, j0 = ty * TILE, tx * TILE
            i1, j1 = min(i0 + TILE, height), min(j0 + TILE, width)
            ii, jj = torch.meshgrid(torch.arange(i0, i1), torch.arange(j0, j1), indexing="ij")
            px = jj.float() + 0.5
            py = ii.float() + 0.5
            T = torch.ones_like(px)
            acc = torch.zeros(*px.shape, ch)
            done = torch.zeros_like(px, dtype=torch.bool)
            cnt = torch.zeros_like(px, dtype=torch.int64)
            for idx in range(lo, hi):
                if bool(done.all()):
                    break
                g = int(gaussian_ids[idx])
                if probed is not None:
                    sigma = probed[0][idx - lo, :i1 - i0, :j1 - j0]
                    alpha = probed[1][idx - lo, :i1 - i0, :j1 - j0]
                else:
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
                cnt = cnt + take.long()
            out[i0:i1, j0:j1] = acc + T[..., None] * background
            final_t[i0:i1, j0:j1] = T
            counts[i0:i1, j0:j1] = cnt
    if want_counts:
        return out, 1.0 - final_t, counts
    return out, 1.0 - final_t


def active_splatfacto_outputs(xys: Tensor, depths: Tensor, conics: Tensor, opacities: Tensor, rgbs: Tensor,
                              betas: Tensor, gaussian_ids: Tensor, tile_bins: Tensor, height: int, width: int,
                              background: Tensor, probe=None, depth_image_override: Tensor = None) -> Dict[str, Tensor]:
    """The rasterisation block of ``ActiveSplatfactoModel.get_outputs``
    (``activesplatfacto_model.py:260-367``) as four separate 3-channel passes, like the reference.
    ``probe``: see ``rasterize``.  ``depth_image_override``: use this depth image for the per-Gaussian residuals of
    the depth-variance pass (tests isolate that pass from the rounding of the depth pass with it)."""
    args = (gaussian_ids, tile_bins, height, width)
    zeros3 = torch.zeros(3)
    kw = {"probe": probe}
    rgb, alpha = rasterize(xys, conics, opacities, rgbs, *args, background, **kw)
    alpha = alpha[..., None]
    rgb = torch.clamp(rgb, max=1.0)
    unc_im = rasterize(xys, conics, opacities, betas.reshape(-1, 1).repeat(1, 3), *args, zeros3, **kw)[0][..., 0:1]
    depth_im = rasterize(xys, conics, opacities, depths[:, None].repeat(1, 3), *args, zeros3, **kw)[0][..., 0:1]
    depth_im = torch.where(alpha > 0, depth_im / alpha, depth_im.detach().max())
    xy_to_pix = torch.floor(xys).long()
    valid = (xy_to_pix[:, 0] > 0) & (xy_to_pix[:, 0] < width) & (xy_to_pix[:, 1] > 0) & (xy_to_pix[:, 1] < height)
    pv = xy_to_pix[valid]
    fetched = (depth_im if depth_image_override is None else depth_image_override)[pv[:, 1], pv[:, 0], 0]
    resid = depths.clone()
    resid[valid] -= fetched
    dvar_im = rasterize(xys, conics, opacities, (resid[:, None] ** 2).repeat(1, 3), *args, zeros3, **kw)[0][..., 0:1]
    dvar_im = torch.where(alpha > 0, dvar_im / alpha, dvar_im.detach().max())
    return {
        "rgb": rgb, "depth": depth_im, "accumulation": alpha, "background": background,
        "uncertainty": unc_im, "rgb_var": unc_im ** 2, "rgb_std": unc_im,
        "depth_var": dvar_im, "depth_std": dvar_im.sqrt(),
    }
