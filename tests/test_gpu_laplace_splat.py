"""GPU parity: last-layer Laplace MC moments and tile alpha-compositing (C ABI) vs the oracle.

Laplace: ``sigma2 = E[y^2] - E[y]^2`` in float32 cancels catastrophically (SURVEY hard part 6), so parity is
pinned on the two moments at 1e-5 relative and on sigma2 with an absolute floor of 1e-5 * E[y^2].
Splat: the pass structure is pinned to the reference (tests/test_oracle_pinned.py); the rasteriser under it restates gsplat 0.1.11 (not vendored).
The alpha < 1/255 and T <= 1e-4 decisions are thresholds, so a pixel may take one Gaussian more or less
when a 1-ulp exp difference crosses them: bounded count of outliers, tight tolerance elsewhere.
"""
import pytest
import torch

from oracle import laplace as ol, splat as osp
from uncertainty_nerf_gs_b200 import binning, synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("out_dim,act,n_points", [(3, "sigmoid", 1000), (1, "exp", 777), (3, "identity", 1)])
def test_laplace_moments(built_library, out_dim, act, n_points):
    from uncertainty_nerf_gs_b200 import ops

    lap = synthetic.laplace_head(n_points, 64, out_dim, 100, seed=out_dim)
    if act == "exp":
        lap["mu_q"] = lap["mu_q"] * 0.2          # keep trunc_exp in range
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
    fn = {"sigmoid": torch.sigmoid, "exp": torch.exp, "identity": lambda v: v}[act]
    mu, mu2, s2 = ol.sample_laplace(lap["x"], theta, out_dim, fn)
    out = ops.laplace_ll_moments(lap["x"].cuda(), theta.cuda(), out_dim, act, want_mean2=True)
    torch.testing.assert_close(out["mean"].cpu(), mu, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["mean2"].cpu(), mu2, rtol=1e-5, atol=1e-6)
    floor = 1e-5 * mu2.abs() + 1e-7
    assert bool(((out["sigma2"].cpu() - s2).abs() <= floor + 1e-5 * s2.abs()).all())


def test_laplace_chunked_samples(built_library):
    """More parameter draws than fit one shared-memory fill (700 x 196 floats > 200 KB)."""
    from uncertainty_nerf_gs_b200 import ops

    lap = synthetic.laplace_head(300, 64, 3, 700, seed=9)
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
    mu, mu2, _ = ol.sample_laplace(lap["x"], theta, 3, torch.sigmoid)
    out = ops.laplace_ll_moments(lap["x"].cuda(), theta.cuda(), 3, "sigmoid", want_mean2=True)
    torch.testing.assert_close(out["mean"].cpu(), mu, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(out["mean2"].cpu(), mu2, rtol=2e-5, atol=1e-6)


def _scene(n, h, w, seed):
    sc = synthetic.splat_scene(n, h, w, seed=seed, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    return sc, ids, bins


@pytest.mark.parametrize("n,hw", [(400, (40, 56)), (5000, (100, 130)), (1, (16, 16)), (300, (17, 33))])
def test_cuda_binning_is_bit_exact(built_library, n, hw):
    """ub_bin_count / ub_bin_gaussians against the torch restatement of gsplat's scheme: identical lists."""
    h, w = hw
    sc = synthetic.splat_scene(n, h, w, seed=n, mean_scale_px=4.0)
    sc["radii"][::7] = 0                                    # culled Gaussians
    sc["depths"][1::5] = sc["depths"][0]                    # depth ties: intersection order must break them
    ids_ref, bins_ref = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    ids, bins = binning.bin_gaussians(sc["xys"].cuda(), sc["depths"].cuda(), sc["radii"].cuda(), h, w)
    assert torch.equal(ids.cpu(), ids_ref) and torch.equal(bins.cpu(), bins_ref)


def test_cuda_binning_no_intersections(built_library):
    sc = synthetic.splat_scene(50, 32, 32, seed=1)
    sc["radii"][:] = 0
    ids, bins = binning.bin_gaussians(sc["xys"].cuda(), sc["depths"].cuda(), sc["radii"].cuda(), 32, 32)
    assert ids.numel() == 0 and bins.shape == (4, 2) and int(bins.abs().max()) == 0


def _close_fraction(a, b, rtol, atol):
    return float(torch.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True).float().mean())


@pytest.mark.parametrize("hw,n", [((40, 56), 400), ((33, 47), 1500)])
def test_active_splatfacto_outputs_without_the_probe(built_library, hw, n):
    """Against the oracle's own ``exp`` (torch's accurate float32 exponential instead of the device's ``__expf``):
    a pixel may take one splat more or less where an alpha sits within an ulp of a threshold, so this variant only
    bounds the outliers; the exact, every-pixel comparison is tests/test_gpu_splat_exact.py."""
    from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs

    h, w = hw
    sc, ids, bins = _scene(n, h, w, seed=n)
    bg = torch.tensor([0.1, 0.2, 0.3])
    ref = osp.active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"],
                                        sc["betas"], ids, bins, h, w, bg)
    out = active_splatfacto_outputs(sc["xys"].cuda(), sc["depths"].cuda(), sc["conics"].cuda(),
                                    sc["opacities"].cuda(), sc["rgbs"].cuda(), sc["betas"].cuda(), ids.cuda(),
                                    bins.cuda(), h, w, bg.cuda())
    assert list(out.keys()) == list(ref.keys())
    for k in ("rgb", "accumulation", "uncertainty", "rgb_var", "rgb_std", "depth"):
        frac = _close_fraction(out[k].cpu(), ref[k], 1e-4, 2e-5)
        assert frac >= 0.995, f"{k}: only {frac:.5f} of pixels within tolerance"


def test_fused_channels_equal_separate_passes(built_library):
    """One 5-channel pass == the reference's separate 3-channel launches, bit for bit (same loop)."""
    from uncertainty_nerf_gs_b200 import ops

    h, w = 64, 80
    sc, ids, bins = _scene(3000, h, w, seed=5)
    c = {k: v.cuda() for k, v in sc.items()}
    ids, bins = ids.cuda(), bins.cuda()
    colors = torch.cat([c["rgbs"], c["betas"], c["depths"][:, None]], dim=1)
    fused, alpha = ops.composite_tiles(c["xys"], c["conics"], c["opacities"], colors, ids, bins, h, w,
                                       [0.1, 0.2, 0.3, 0.0, 0.0])
    rgb, alpha2 = ops.composite_tiles(c["xys"], c["conics"], c["opacities"], c["rgbs"], ids, bins, h, w,
                                      [0.1, 0.2, 0.3])
    beta3, _ = ops.composite_tiles(c["xys"], c["conics"], c["opacities"], c["betas"].repeat(1, 3), ids, bins, h, w)
    assert torch.equal(fused[..., :3], rgb) and torch.equal(alpha, alpha2)
    assert torch.equal(fused[..., 3], beta3[..., 0])


def test_empty_tiles_and_background(built_library):
    from uncertainty_nerf_gs_b200 import ops

    h, w = 20, 30
    tiles = 2 * 2
    z = lambda *s: torch.zeros(*s, device="cuda")
    out, alpha = ops.composite_tiles(z(1, 2), z(1, 3), z(1), z(1, 3), torch.zeros(0, dtype=torch.int32, device="cuda"),
                                     torch.zeros(tiles, 2, dtype=torch.int32, device="cuda"), h, w, [0.5, 0.25, 1.0])
    assert float(alpha.abs().max()) == 0.0
    assert torch.equal(out, torch.tensor([0.5, 0.25, 1.0], device="cuda").expand(h, w, 3))


@pytest.mark.parametrize("n_samples,n_points", [(100, 4099), (1, 130), (33, 257), (101, 64)])
def test_laplace_tensor_core_path_vs_fma_and_oracle(built_library, n_samples, n_points):
    """rgb head on tcgen05 (3xTF32 split) against the fp32-FMA kernel and the oracle: partial last chunk
    (3 * n_samples not a multiple of 16), partial last tile, single draw."""
    from uncertainty_nerf_gs_b200 import ops

    lap = synthetic.laplace_head(n_points, 64, 3, n_samples, seed=n_samples)
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
    mu, mu2, _ = ol.sample_laplace(lap["x"], theta, 3, torch.sigmoid)
    tc = ops.laplace_ll_moments(lap["x"].cuda(), theta.cuda(), 3, "sigmoid", want_mean2=True, tensor_cores=True)
    fma = ops.laplace_ll_moments(lap["x"].cuda(), theta.cuda(), 3, "sigmoid", want_mean2=True, tensor_cores=False)
    for out in (tc, fma):
        torch.testing.assert_close(out["mean"].cpu(), mu, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out["mean2"].cpu(), mu2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tc["mean"], fma["mean"], rtol=1e-5, atol=1e-6)


def test_laplace_tensor_core_path_large_magnitudes(built_library):
    """Pre-activations of magnitude ~30 (saturating sigmoids) and features spanning 6 orders of magnitude:
    the hi/lo split must keep float32-level accuracy."""
    from uncertainty_nerf_gs_b200 import ops

    lap = synthetic.laplace_head(2000, 64, 3, 100, seed=77)
    x = lap["x"] * torch.logspace(-3, 3, 64)[None, :] * 0.05
    theta = ol.posterior_samples(lap["mu_q"] * 3.0, lap["ggn"], lap["eps_draws"])
    mu, mu2, _ = ol.sample_laplace(x, theta, 3, torch.sigmoid)
    out = ops.laplace_ll_moments(x.cuda(), theta.cuda(), 3, "sigmoid", want_mean2=True)
    torch.testing.assert_close(out["mean"].cpu(), mu, rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(out["mean2"].cpu(), mu2, rtol=2e-5, atol=2e-6)


# ---- producers of the splat path: projection + spherical harmonics (row f2) ----
@pytest.mark.parametrize("n,hw", [(5000, (120, 160)), (257, (33, 47))])
def test_project_gaussians_matches_oracle(built_library, n, hw):
    h, w = hw
    sc = synthetic.gaussians_3d(n, h, w, seed=n)
    ref = osp.project_gaussians(sc["means"], sc["scales"], 1.0, sc["quats"], sc["viewmat"], sc["fx"], sc["fy"],
                                sc["cx"], sc["cy"], h, w)
    xys, depths, radii, conics, comp, tiles, cov3d = binning.project_gaussians(
        sc["means"].cuda(), sc["scales"].cuda(), 1, sc["quats"].cuda(), sc["viewmat"].cuda(), sc["fx"], sc["fy"],
        sc["cx"], sc["cy"], h, w, 16)
    # integer outputs: identical unless 3 sqrt(lambda_max) sits within float32 rounding of an integer
    frac = ref["radius_real"] - torch.floor(ref["radius_real"])
    sure = (frac > 1e-3) & (frac < 1 - 1e-3) | (ref["radii"] == 0)
    visible = ref["radii"] > 0
    assert int(visible.sum()) > n // 10 and int((~visible).sum()) > 0            # both populations are exercised
    assert torch.equal(radii.cpu()[sure], ref["radii"][sure])
    assert torch.equal(tiles.cpu()[sure], ref["num_tiles_hit"][sure])
    same = radii.cpu() == ref["radii"]
    for name, got, want in (("xys", xys, ref["xys"]), ("depths", depths, ref["depths"]),
                            ("compensation", comp, ref["compensation"])):
        torch.testing.assert_close(got.cpu()[same], want[same], rtol=2e-4, atol=2e-4, msg=lambda m: f"{name}: {m}")
    torch.testing.assert_close(conics.cpu(), ref["conics"], rtol=5e-4, atol=1e-6)
    # off-diagonal covariance entries are sums with cancellation: absolute tolerance relative to the matrix scale
    torch.testing.assert_close(cov3d.cpu(), ref["cov3d"], rtol=1e-4, atol=1e-6 * float(ref["cov3d"].abs().max()))


@pytest.mark.parametrize("use", [0, 1, 2, 3])
def test_spherical_harmonics_matches_oracle(built_library, use):
    sc = synthetic.gaussians_3d(3001, 64, 64, seed=use)
    dirs = sc["means"] - sc["camera_position"]
    ref = osp.spherical_harmonics(use, dirs, sc["sh_coeffs"])
    got = binning.spherical_harmonics(use, dirs.cuda(), sc["sh_coeffs"].cuda())
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-5, atol=2e-6)
    with pytest.raises(Exception):
        binning.spherical_harmonics(4, dirs.cuda(), sc["sh_coeffs"].cuda())      # more degrees than coefficients


def test_projection_to_image_chain(built_library):
    """3-D scene -> project -> SH colours -> bin -> fused active-splatfacto outputs, against the oracle chain."""
    from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs

    h, w, n = 48, 64, 600
    sc = synthetic.gaussians_3d(n, h, w, seed=11)
    cu = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in sc.items()}
    xys, depths, radii, conics, comp, tiles, _ = binning.project_gaussians(
        cu["means"], cu["scales"], 1, cu["quats"], cu["viewmat"], sc["fx"], sc["fy"], sc["cx"], sc["cy"], h, w, 16)
    rgbs = torch.clamp(binning.spherical_harmonics(3, cu["means"] - cu["camera_position"], cu["sh_coeffs"]) + 0.5, min=0.0)
    ids, bins = binning.bin_gaussians(xys, depths, radii, h, w)
    assert int(tiles.sum()) == ids.numel()                                       # num_tiles_hit is the binning's count
    bg = torch.tensor([0.2, 0.3, 0.4])
    out = active_splatfacto_outputs(xys, depths, conics, cu["opacities"], rgbs, cu["betas"], ids, bins, h, w, bg.cuda())
    # oracle chain on the GPU's projected quantities (the projection itself is compared above)
    ref = osp.active_splatfacto_outputs(xys.cpu(), depths.cpu(), conics.cpu(), sc["opacities"], rgbs.cpu(), sc["betas"],
                                        ids.cpu(), bins.cpu(), h, w, bg)
    for k in ("rgb", "depth", "accumulation", "uncertainty", "depth_var"):
        torch.testing.assert_close(out[k].cpu(), ref[k], rtol=2e-4, atol=2e-5, equal_nan=True, msg=lambda m: f"{k}: {m}")


_CULL_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r})
from oracle import splat as osp
from uncertainty_nerf_gs_b200 import ops, synthetic
h, w = 70, 100
sc = synthetic.splat_scene(3000, h, w, seed=21, mean_scale_px=5.0)
sc["opacities"][::11] = 0.003          # below 1/255: can never contribute
sc["opacities"][5::13] = 0.9995        # clamped alpha
sc["opacities"][7::301] = float("nan")
sc["conics"][3::97, 1] = 5.0           # indefinite conic (sigma < 0 somewhere): no cull allowed
sc["conics"][9::211] = float("nan")
ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
cu = {{k: v.cuda() for k, v in sc.items()}}
outs, alpha, _ = ops.composite_tiles_planes(cu["xys"], cu["conics"], cu["opacities"], [cu["rgbs"], cu["betas"], cu["depths"][:, None].contiguous()],
                                            ids.cuda(), bins.cuda(), h, w, [0.1, 0.2, 0.3, 0.0, 0.0])
torch.save([o.cpu() for o in outs] + [alpha.cpu()], {out!r})
"""


def test_warp_row_cull_changes_nothing(built_library, tmp_path):
    """The warp-level row cull of the tile compositor must be invisible: same bits with UB_TILES_NO_CULL=1."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    results = []
    for flag in ("0", "1"):
        out = str(tmp_path / f"tiles_{flag}.pt")
        env = dict(os.environ, UB_TILES_NO_CULL=flag)
        subprocess.run([sys.executable, "-c", _CULL_SCRIPT.format(root=root, out=out)], check=True, env=env, timeout=300)
        results.append(torch.load(out))
    for a, b in zip(*results):
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
    assert float(torch.nan_to_num(results[0][-1]).max()) > 0.5


def test_splat_producers_edge_cases(built_library):
    """Empty scenes, everything behind the camera, degree-0 colours, and a backward pass with no intersections."""
    from uncertainty_nerf_gs_b200 import ops
    from uncertainty_nerf_gs_b200.autograd import composite_tiles_train

    dev = "cuda"
    vm = torch.eye(4, device=dev)[:3]
    empty = binning.project_gaussians(torch.zeros(0, 3, device=dev), torch.zeros(0, 3, device=dev), 1,
                                      torch.zeros(0, 4, device=dev), vm, 50.0, 50.0, 16.0, 16.0, 32, 32, 16)
    assert all(t.shape[0] == 0 for t in empty)
    sc = synthetic.gaussians_3d(64, 32, 32, seed=3)
    behind = sc["means"].clone()
    behind[:, 2] = -5.0                                                   # camera looks down +z with this view matrix
    xys, depths, radii, conics, comp, tiles, cov3d = binning.project_gaussians(
        behind.cuda(), sc["scales"].cuda(), 1, sc["quats"].cuda(), vm, 50.0, 50.0, 16.0, 16.0, 32, 32, 16)
    assert int(radii.abs().sum()) == 0 and int(tiles.sum()) == 0 and float(cov3d.abs().sum()) == 0.0
    ids, bins = binning.bin_gaussians(xys, depths, radii, 32, 32)
    assert ids.numel() == 0
    col0 = binning.spherical_harmonics(0, sc["means"].cuda(), sc["sh_coeffs"][:, :1].contiguous().cuda())
    torch.testing.assert_close(col0.cpu(), 0.28209479177387814 * sc["sh_coeffs"][:, 0], rtol=1e-6, atol=1e-7)
    # nothing intersects: the image is the background, alpha 0, and every gradient is exactly zero
    rgbs = torch.rand(64, 3, device=dev, requires_grad=True)
    x = xys.clone().requires_grad_(True)
    (img,), a = composite_tiles_train(x, conics + 1.0, torch.full((64,), 0.5, device=dev), [rgbs], ids, bins, 32, 32,
                                      [0.25, 0.5, 0.75])
    assert float(a.detach().abs().max()) == 0.0 and torch.equal(img.detach()[0, 0].cpu(), torch.tensor([0.25, 0.5, 0.75]))
    (img.sum() + a.sum()).backward()
    assert float(rgbs.grad.abs().max()) == 0.0 and float(x.grad.abs().max()) == 0.0


def test_laplace_density_head_on_tensor_cores_opt_in(built_library):
    """UB_LAPLACE_TC_DENSITY=1 (read once per process, so a child process): the density head as a [P,64] x [64,100]
    tcgen05 GEMM.  Faster, but exp amplifies the tensor core's truncating accumulation: 5e-5 relative, not 1e-5 --
    the reason it is not the default."""
    import os
    import subprocess
    import sys

    code = r'''
import torch, sys
sys.path.insert(0, ".")
from oracle import laplace as ol
from uncertainty_nerf_gs_b200 import ops, synthetic
lap = synthetic.laplace_head(1500, 64, 1, 100, seed=1)
lap["mu_q"] = lap["mu_q"] * 0.2
theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
mu, mu2, s2 = ol.sample_laplace(lap["x"], theta, 1, torch.exp)
out = ops.laplace_ll_moments(lap["x"].cuda(), theta.cuda(), 1, "exp", want_mean2=True)
torch.testing.assert_close(out["mean"].cpu(), mu, rtol=5e-5, atol=1e-6)
torch.testing.assert_close(out["mean2"].cpu(), mu2, rtol=1e-4, atol=1e-6)
print("ok")
'''
    env = dict(os.environ, UB_LAPLACE_TC_DENSITY="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr
