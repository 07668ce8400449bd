"""GPU: the tile compositor against the oracle with EXACT threshold decisions.

Contract (`composite_tiles_kernel`): gsplat 0.1.11's per-pixel loop in gsplat's order --
``sigma < 0 or alpha < 1/255 -> skip``, ``T (1 - alpha) <= 1e-4 -> stop before this splat`` -- with
``alpha = min(0.999, opacity * __expf(-sigma))``, the fast intrinsic gsplat's own kernel uses.  A CPU cannot reproduce
``__expf``, so the oracle takes the per-(splat, pixel) ``(sigma, alpha)`` values from the device
(``ub_tile_alpha_probe``: the same device function, rounding pinned) and replays every decision, the transmittance
products and the accumulation in torch float32.  Consequences demanded here, on EVERY pixel, no exempt fraction:

* the alpha image ``1 - T_final`` is bit-identical (=> the set of contributing splats of every pixel is identical:
  T is a product of the same float32 factors in the same order);
* every composited channel agrees to 1e-5 relative (the kernel accumulates with FMA, torch with mul + add).
"""
import numpy as np
import pytest
import torch

from oracle import splat as osp
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu


def _probe_for(c, ids_gpu, tiles_x):
    from uncertainty_nerf_gs_b200 import ops

    def probe(tile, lo, hi):
        s, a = ops.tile_alpha_probe(c["xys"], c["conics"], c["opacities"], ids_gpu, lo, hi - lo, tile % tiles_x,
                                    tile // tiles_x)
        return s.cpu(), a.cpu()
    return probe


def _assert_channels(got, want, name, rtol=1e-5, atol=2e-6):
    torch.testing.assert_close(got, want, rtol=rtol, atol=atol, equal_nan=True, msg=lambda m: f"{name}: {m}")


@pytest.mark.parametrize("hw,n", [((40, 56), 400), ((33, 47), 1500), ((64, 80), 6000)])
def test_fused_pass_has_the_oracles_contributing_sets(built_library, hw, n):
    from uncertainty_nerf_gs_b200 import ops

    h, w = hw
    sc = synthetic.splat_scene(n, h, w, seed=n, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    c = {k: v.cuda() for k, v in sc.items()}
    ids_g, bins_g = ids.cuda(), bins.cuda()
    tiles_x = (w + 15) // 16
    planes = [c["rgbs"], c["betas"], c["depths"][:, None].contiguous()]
    bg = [0.1, 0.2, 0.3, 0.0, 0.0]
    (rgb, beta, depth), alpha, _ = ops.composite_tiles_planes(c["xys"], c["conics"], c["opacities"], planes, ids_g, bins_g,
                                                              h, w, bg)
    colors = torch.cat([sc["rgbs"], sc["betas"], sc["depths"][:, None]], dim=1)
    want, want_alpha, counts = osp.rasterize(sc["xys"], sc["conics"], sc["opacities"], colors, ids, bins, h, w,
                                             torch.tensor(bg), probe=_probe_for(c, ids_g, tiles_x), want_counts=True)
    assert torch.equal(alpha[..., 0].cpu(), want_alpha)                    # bit-exact: identical contributing sets
    assert int(counts.max()) > 3 and int((counts == 0).sum()) >= 0
    _assert_channels(rgb.cpu(), want[..., :3], "rgb")
    _assert_channels(beta.cpu(), want[..., 3:4], "beta")
    _assert_channels(depth.cpu(), want[..., 4:5], "depth", atol=2e-5)      # depths up to 10: atol scales with the value


@pytest.mark.parametrize("hw,n", [((40, 56), 400), ((33, 47), 1500)])
def test_active_splatfacto_outputs_every_pixel(built_library, hw, n):
    """All nine outputs of the reference's rasterisation block (activesplatfacto_model.py:260-367), every pixel."""
    from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs

    h, w = hw
    sc = synthetic.splat_scene(n, h, w, seed=n, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    c = {k: v.cuda() for k, v in sc.items()}
    bgt = torch.tensor([0.1, 0.2, 0.3])
    out = active_splatfacto_outputs(c["xys"], c["depths"], c["conics"], c["opacities"], c["rgbs"], c["betas"], ids.cuda(),
                                    bins.cuda(), h, w, bgt.cuda())
    # the depth-variance pass is checked as a function of ITS inputs: the per-Gaussian residuals are taken against the
    # device's depth image (a residual d - D amplifies the 1e-7 rounding of D by D / |d - D|, which no tolerance on
    # depth_var itself could bound)
    ref = osp.active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"], sc["betas"],
                                        ids, bins, h, w, bgt, probe=_probe_for(c, ids.cuda(), (w + 15) // 16),
                                        depth_image_override=out["depth"].cpu())
    assert list(out.keys()) == list(ref.keys())
    assert torch.equal(out["accumulation"].cpu(), ref["accumulation"])
    for k in ("rgb", "uncertainty", "rgb_var", "rgb_std"):
        _assert_channels(out[k].cpu(), ref[k], k)
    for k in ("depth", "depth_var", "depth_std"):
        _assert_channels(out[k].cpu(), ref[k], k, atol=2e-5)
    assert torch.equal(out["background"].cpu(), bgt)


def test_depth_residuals_are_the_reference_torch_lines(built_library):
    """``ub_splat_depth_residual`` vs activesplatfacto_model.py:325-341, bit for bit (gather at the floor of the centre,
    strict ``0 < x < W`` mask)."""
    from uncertainty_nerf_gs_b200 import ops

    h, w, n = 40, 56, 3000
    sc = synthetic.splat_scene(n, h, w, seed=3, mean_scale_px=4.0)
    sc["xys"][:5] = torch.tensor([[0.5, 10.0], [1.0, 1.0], [w - 0.5, h - 0.5], [float(w), 3.0], [5.0, 0.99]])
    depth_im = torch.rand(h, w, 1, generator=torch.Generator().manual_seed(1)) * 9 + 0.1
    xy_to_pix = torch.floor(sc["xys"]).long()
    valid = (xy_to_pix[:, 0] > 0) & (xy_to_pix[:, 0] < w) & (xy_to_pix[:, 1] > 0) & (xy_to_pix[:, 1] < h)
    pv = xy_to_pix[valid]
    resid = sc["depths"].clone()
    resid[valid] -= depth_im[pv[:, 1], pv[:, 0], 0]
    got = ops.splat_depth_residual(sc["xys"].cuda(), sc["depths"].cuda(), depth_im.cuda())
    assert torch.equal(got.cpu(), resid[:, None] ** 2)


def test_one_million_gaussians_on_random_tiles(built_library):
    """configs[3] at its real size: 64 random tiles of the 1 M-Gaussian 1297 x 840 view (~2400 splats per tile) through
    the oracle; alpha bit-exact, rgb / beta / depth to 1e-5 on every pixel of those tiles."""
    from uncertainty_nerf_gs_b200 import binning, ops

    h, w, g = 840, 1297, 1_000_000
    sc = synthetic.splat_scene(g, h, w, seed=0, device="cuda")
    ids, bins = binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    tiles_x, tiles_y = (w + 15) // 16, (h + 15) // 16
    rng = np.random.default_rng(0)
    tiles = sorted(set(rng.choice(tiles_x * tiles_y, size=62, replace=False).tolist()) | {0, tiles_x * tiles_y - 1})
    planes = [sc["rgbs"], sc["betas"], sc["depths"][:, None].contiguous()]
    bg = [0.2, 0.4, 0.6, 0.0, 0.0]
    (rgb, beta, depth), alpha, _ = ops.composite_tiles_planes(sc["xys"], sc["conics"], sc["opacities"], planes, ids, bins,
                                                              h, w, bg)
    # the oracle only touches the chosen tiles' list entries: move those to the CPU, remapped to a compact id space
    bins_c = bins.cpu()
    spans = [(int(bins_c[t, 0]), int(bins_c[t, 1])) for t in tiles]
    local_ids = torch.cat([ids[lo:hi] for lo, hi in spans]).long()
    uniq, inv = torch.unique(local_ids, return_inverse=True)
    small = {k: sc[k][uniq].cpu() for k in ("xys", "conics", "opacities", "rgbs", "betas", "depths")}
    new_bins = torch.zeros_like(bins_c)
    pos = 0
    for t, (lo, hi) in zip(tiles, spans):
        new_bins[t] = torch.tensor([pos, pos + hi - lo])
        pos += hi - lo
    new_ids = inv.to(torch.int32).cpu()
    start_of = {t: lo for t, (lo, hi) in zip(tiles, spans)}

    def probe(tile, lo, hi):       # the oracle's compact positions -> the device's list positions
        first = start_of[tile]
        s, a = ops.tile_alpha_probe(sc["xys"], sc["conics"], sc["opacities"], ids, first, hi - lo, tile % tiles_x,
                                    tile // tiles_x)
        return s.cpu(), a.cpu()

    colors = torch.cat([small["rgbs"], small["betas"], small["depths"][:, None]], dim=1)
    want, want_alpha, counts = osp.rasterize(small["xys"], small["conics"], small["opacities"], colors, new_ids, new_bins,
                                             h, w, torch.tensor(bg), tiles=tiles, probe=probe, want_counts=True)
    mask = torch.zeros(h, w, dtype=torch.bool)
    for t in tiles:
        ty, tx = divmod(t, tiles_x)
        mask[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] = True
    assert int(mask.sum()) >= 60 * 256
    assert torch.equal(alpha[..., 0].cpu()[mask], want_alpha[mask])
    assert float(counts[mask].float().mean()) > 5
    _assert_channels(rgb.cpu()[mask], want[..., :3][mask], "rgb")
    _assert_channels(beta.cpu()[mask], want[..., 3:4][mask], "beta")
    _assert_channels(depth.cpu()[mask], want[..., 4:5][mask], "depth", atol=2e-5)
