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
