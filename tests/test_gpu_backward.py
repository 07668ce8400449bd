"""GPU: gradients of the fused compositor (ub_composite_rays_backward through the autograd Function) against
torch autograd on the CPU oracle in training mode (row f1 of SURVEY.md section 8)."""
import pytest
import torch

from oracle import compositing as oc
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu


def _oracle_train_outputs(density, deltas, starts, ends, rgb, beta, background):
    w = oc.get_weights(density, deltas)
    out_rgb = oc.render_rgb(rgb, w, background, training=True)
    with torch.no_grad():
        depth = oc.render_depth_median(w, starts, ends)
    expected = oc.render_depth_expected(w, starts, ends)
    acc = oc.render_accumulation(w)
    var = oc.render_uncertainty(beta, w ** 2)
    dvar = oc.depth_variance(w, starts, ends, depth)
    return {"rgb": out_rgb, "accumulation": acc, "expected_depth": expected, "rgb_var": var, "rgb_std": var.sqrt(),
            "depth_var": dvar, "depth_std": dvar.sqrt(), "weights": w}


@pytest.mark.parametrize("num_samples,background", [(48, "last_sample"), (32, (0.2, 0.5, 0.9)), (64, "random"),
                                                    (96, "last_sample")])
def test_backward_matches_autograd_on_the_oracle(built_library, num_samples, background):
    from uncertainty_nerf_gs_b200.autograd import composite_rays_train

    R = 333
    inp = synthetic.ray_samples(R, num_samples, seed=num_samples, edge_cases=False)
    inp["starts"][:40] += 2.0   # some empty-ish rays get clipped to the chunk bounds
    inp["ends"][:40] += 2.0
    inp["density"][:5] = 0.0    # empty rays: expected depth 0 is clipped to the chunk minimum (zero gradient);
    #                             rgb_std = sqrt(0) makes torch's sqrt backward produce NaN there -- must match too
    g = torch.Generator().manual_seed(1)
    coef = {k: torch.randn(s, generator=g) for k, s in
            (("rgb", (R, 3)), ("accumulation", (R, 1)), ("expected_depth", (R, 1)), ("rgb_var", (R, 1)),
             ("rgb_std", (R, 1)), ("depth_var", (R, 1)), ("depth_std", (R, 1)), ("weights", (R, num_samples, 1)))}

    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("density", "rgb", "beta")}
    ref = _oracle_train_outputs(leaves["density"], inp["deltas"], inp["starts"], inp["ends"], leaves["rgb"],
                                leaves["beta"], background)
    sum((ref[k] * coef[k]).sum() for k in coef).backward()

    cu = {k: inp[k].cuda() for k in inp}
    for k in ("density", "rgb", "beta"):
        cu[k].requires_grad_(True)
    out = composite_rays_train(cu["density"], cu["deltas"], cu["starts"], cu["ends"], cu["rgb"], cu["beta"], background)
    for k in coef:
        torch.testing.assert_close(out[k].detach().cpu(), ref[k].detach(), rtol=1e-5, atol=2e-6, equal_nan=True,
                                   msg=lambda m: f"fwd {k}: {m}")
    sum((out[k] * coef[k].cuda()).sum() for k in coef).backward()
    for k in ("density", "rgb", "beta"):
        a, b = cu[k].grad.cpu(), leaves[k].grad
        scale = float(torch.nan_to_num(b).abs().max())
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-6 * max(scale, 1.0), equal_nan=True,
                                   msg=lambda m: f"grad {k}: {m}")
        assert bool(torch.isfinite(b[5:]).all())
    assert not out["depth"].requires_grad


def test_backward_partial_gradients_and_unsupported_shape(built_library):
    from uncertainty_nerf_gs_b200 import ops
    from uncertainty_nerf_gs_b200.autograd import composite_rays_train

    inp = {k: v.cuda() for k, v in synthetic.ray_samples(64, 48, seed=2, edge_cases=False).items()}
    inp["density"].requires_grad_(True)
    out = composite_rays_train(**inp)
    out["rgb"].sum().backward()                       # only one output feeds the loss; rgb / beta need no grad
    assert inp["density"].grad is not None and bool(torch.isfinite(inp["density"].grad).all())
    bad = {k: v.cuda() for k, v in synthetic.ray_samples(8, 50, seed=2, edge_cases=False).items()}
    bad["density"].requires_grad_(True)
    o = composite_rays_train(**bad)
    with pytest.raises(ops.UBError):
        o["rgb"].sum().backward()


# ---- tile compositor (splat) backward: ub_composite_tiles_planes_backward vs autograd through the oracle rasteriser ----
def _splat_leaves(sc):
    return {k: sc[k].clone().requires_grad_(True) for k in ("xys", "conics", "opacities", "rgbs", "betas")}


@pytest.mark.parametrize("n,hw,bg", [(400, (40, 56), (0.1, 0.2, 0.3)), (300, (17, 33), (0.0, 0.0, 0.0)),
                                      (1500, (48, 48), (0.9, 0.5, 0.2))])
def test_tile_backward_matches_autograd_on_the_oracle(built_library, n, hw, bg):
    from oracle import splat as osp
    from uncertainty_nerf_gs_b200.autograd import composite_tiles_train

    h, w = hw
    sc = synthetic.splat_scene(n, h, w, seed=n + 1, mean_scale_px=4.0)
    sc["opacities"][::9] = 0.9995            # some splats hit the alpha clamp (zero geometry / opacity gradient)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    g = torch.Generator().manual_seed(3)
    c_rgb, c_beta, c_alpha = (torch.randn(h, w, 3, generator=g), torch.randn(h, w, 1, generator=g),
                              torch.randn(h, w, generator=g))

    # oracle: two separate passes sharing the geometry, like the reference's two rasterize_gaussians calls
    lv = _splat_leaves(sc)
    rgb, alpha = osp.rasterize(lv["xys"], lv["conics"], lv["opacities"], lv["rgbs"], ids, bins, h, w, torch.tensor(bg))
    beta = osp.rasterize(lv["xys"], lv["conics"], lv["opacities"], lv["betas"].reshape(-1, 1), ids, bins, h, w,
                         torch.zeros(1))[0]
    ((rgb * c_rgb).sum() + (beta * c_beta).sum() + (alpha * c_alpha).sum()).backward()

    cu = {k: sc[k].cuda().requires_grad_(True) for k in ("xys", "conics", "opacities", "rgbs", "betas")}
    outs, a = composite_tiles_train(cu["xys"], cu["conics"], cu["opacities"], [cu["rgbs"], cu["betas"].reshape(-1, 1)],
                                    ids.cuda(), bins.cuda(), h, w, list(bg) + [0.0])
    torch.testing.assert_close(outs[0].cpu().detach(), rgb.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(outs[1].cpu().detach(), beta.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(a[..., 0].cpu().detach(), alpha.detach(), rtol=1e-4, atol=1e-5)
    ((outs[0] * c_rgb.cuda()).sum() + (outs[1] * c_beta.cuda()).sum() + (a[..., 0] * c_alpha.cuda()).sum()).backward()
    for k in ("xys", "conics", "opacities", "rgbs", "betas"):
        got, want = cu[k].grad.cpu(), lv[k].grad
        scale = float(want.abs().max())
        torch.testing.assert_close(got, want, rtol=2e-3, atol=2e-5 * max(scale, 1.0), msg=lambda m: f"grad {k}: {m}")
    assert float(lv["xys"].grad.abs().max()) > 0 and float(lv["opacities"].grad.abs().max()) > 0


def test_tile_backward_zero_upstream_and_partial_planes(built_library):
    """A plane whose upstream gradient is None contributes nothing; planes that do not require grad get none."""
    from oracle import splat as osp
    from uncertainty_nerf_gs_b200.autograd import composite_tiles_train

    h, w = 33, 40
    sc = synthetic.splat_scene(200, h, w, seed=9, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    xys = sc["xys"].cuda().requires_grad_(True)
    rgbs = sc["rgbs"].cuda().requires_grad_(True)
    depth_plane = sc["depths"].reshape(-1, 1).cuda()                      # no grad wanted
    outs, a = composite_tiles_train(xys, sc["conics"].cuda(), sc["opacities"].cuda(), [rgbs, depth_plane],
                                    ids.cuda(), bins.cuda(), h, w)
    outs[0].sum().backward()                                              # depth plane and alpha unused
    lv_x = sc["xys"].clone().requires_grad_(True)
    lv_c = sc["rgbs"].clone().requires_grad_(True)
    ref, _ = osp.rasterize(lv_x, sc["conics"], sc["opacities"], lv_c, ids, bins, h, w, torch.zeros(3))
    ref.sum().backward()
    torch.testing.assert_close(rgbs.grad.cpu(), lv_c.grad, rtol=2e-3, atol=1e-5)
    torch.testing.assert_close(xys.grad.cpu(), lv_x.grad, rtol=2e-3, atol=2e-5 * max(1.0, float(lv_x.grad.abs().max())))
