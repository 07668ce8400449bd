"""GPU parity: fused ray compositing (C ABI) vs the torch-CPU oracle on identical seeded inputs.

Tolerances (north star: float outputs within 1e-5 relative in fp32):
* per-ray sums: rtol 1e-5 plus a small absolute floor for values that are sums of O(1) terms --
  float32 summation order and 1-ulp differences between CUDA ``expf`` and torch's vectorised CPU
  ``exp`` are not controllable (an alpha = 1 - exp(-dd) with tiny dd differs by one ulp of 1.0).
* median depth is an *index* decision (searchsorted(cumsum(w), 0.5)): it must match exactly except on
  rays whose oracle cumulative weight at the deciding sample is within a few ulp of 0.5, where the
  1-ulp exp difference may move the index by one sample; those rays are counted and bounded.
"""
import numpy as np
import pytest
import torch

from oracle import compositing as oc
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _borderline_rays(inp, tol=2e-6):
    """rays whose oracle cumulative weight comes within `tol` of 0.5 at some sample"""
    w = oc.get_weights(inp["density"], inp["deltas"])
    cw = torch.cumsum(w[..., 0], dim=-1)
    return (cw - 0.5).abs().min(dim=-1).values < tol


def _check(out, ref, inp, keys, atol):
    border = _borderline_rays(inp)
    depth_ok = torch.isclose(out["depth"].cpu(), ref["depth"], rtol=RTOL, atol=0.0)[:, 0]
    bad = ~depth_ok & ~border
    assert not bad.any(), f"{int(bad.sum())} non-borderline rays with a different median sample"
    assert int((~depth_ok).sum()) <= max(2, int(1e-4 * depth_ok.numel())), "too many borderline flips"
    for k in keys:
        a, b = out[k].cpu(), ref[k]
        if k in ("depth_var", "depth_std", "depth"):
            a, b = a[depth_ok], b[depth_ok]
        torch.testing.assert_close(a, b, rtol=RTOL, atol=atol[k], equal_nan=True, msg=lambda m: f"{k}: {m}")


ATOL = {"rgb": 2e-6, "accumulation": 2e-6, "expected_depth": 2e-6, "rgb_var": 1e-7, "rgb_std": 1e-6,
        "depth_var": 1e-7, "depth_std": 1e-6, "depth": 0.0}
KEYS = ["rgb", "accumulation", "depth", "expected_depth", "rgb_var", "rgb_std", "depth_var", "depth_std"]


@pytest.mark.parametrize("num_rays,num_samples", [(4096, 48), (1001, 48), (7, 48), (333, 32), (500, 64), (129, 96)])
def test_active_nerfacto_fast_path(built_library, num_rays, num_samples):
    from uncertainty_nerf_gs_b200.models.outputs import active_nerfacto_outputs

    inp = synthetic.ray_samples(num_rays, num_samples, seed=num_rays)
    ref = oc.active_nerfacto_outputs(**inp)
    out = active_nerfacto_outputs(**_cuda(inp), return_weights=True)
    assert list(out.keys())[:9] == list(ref.keys())
    w_ref = oc.get_weights(inp["density"], inp["deltas"])
    torch.testing.assert_close(out["weights"].cpu(), w_ref, rtol=RTOL, atol=2e-7)
    _check(out, ref, inp, KEYS, ATOL)


@pytest.mark.parametrize("num_samples", [5, 31, 50, 100, 256])
def test_active_nerfacto_generic_path(built_library, num_samples):
    from uncertainty_nerf_gs_b200.models.outputs import active_nerfacto_outputs

    inp = synthetic.ray_samples(257, num_samples, seed=num_samples)
    ref = oc.active_nerfacto_outputs(**inp)
    out = active_nerfacto_outputs(**_cuda(inp), return_weights=True)
    torch.testing.assert_close(out["weights"].cpu(), oc.get_weights(inp["density"], inp["deltas"]),
                               rtol=RTOL, atol=2e-7)
    _check(out, ref, inp, KEYS, ATOL)


def test_chunked_eval_matches_reference_chunk_loop(built_library):
    """Chunk-wide reductions (clip bounds of expected depth, beta NaN guard) follow the eval chunk."""
    from uncertainty_nerf_gs_b200.models.outputs import active_nerfacto_outputs

    chunk = 256
    inp = synthetic.ray_samples(1000, 48, seed=5)
    # make the per-chunk bounds differ and bite: shift rays of the 2nd chunk, empty rays clip to chunk min
    inp["starts"][chunk:2 * chunk] += 3.0
    inp["ends"][chunk:2 * chunk] += 3.0
    ref = oc.render_in_chunks(oc.active_nerfacto_outputs, chunk, inp["density"], inp["deltas"], inp["starts"],
                              inp["ends"], inp["rgb"], inp["beta"])
    out = active_nerfacto_outputs(**_cuda(inp), rays_per_chunk=chunk)
    _check(out, ref, inp, KEYS, ATOL)
    empty = ref["accumulation"][:, 0] == 0
    assert empty.any()
    torch.testing.assert_close(out["expected_depth"].cpu()[empty], ref["expected_depth"][empty], rtol=0, atol=0)


def test_beta_inf_with_nan_in_chunk(built_library):
    """isnan(beta).any() -> nan_to_num(beta) also rewrites +-inf in that chunk, and only there."""
    from uncertainty_nerf_gs_b200.models.outputs import active_nerfacto_outputs

    chunk = 128
    inp = synthetic.ray_samples(256, 48, seed=11, edge_cases=False)
    inp["beta"][3, 7] = float("nan")      # chunk 0 has a NaN ...
    inp["beta"][5, 2] = float("inf")      # ... so this inf becomes FLT_MAX
    inp["beta"][200, 4] = float("inf")    # chunk 1 has no NaN: inf stays
    ref = oc.render_in_chunks(oc.active_nerfacto_outputs, chunk, inp["density"], inp["deltas"], inp["starts"],
                              inp["ends"], inp["rgb"], inp["beta"])
    out = active_nerfacto_outputs(**_cuda(inp), rays_per_chunk=chunk)
    torch.testing.assert_close(out["rgb_var"].cpu(), ref["rgb_var"], rtol=RTOL, atol=1e-7, equal_nan=True)
    assert torch.isfinite(ref["rgb_var"][5]) and not torch.isfinite(ref["rgb_var"][200])


@pytest.mark.parametrize("background", ["last_sample", "random", (0.0, 0.0, 0.0), (1.0, 0.5, 0.25)])
def test_backgrounds(built_library, background):
    from uncertainty_nerf_gs_b200 import ops

    inp = synthetic.ray_samples(512, 48, seed=3)
    w = oc.get_weights(inp["density"], inp["deltas"])
    ref = oc.render_rgb(inp["rgb"], w, background)
    c = _cuda(inp)
    out = ops.composite_rays(c["density"], c["deltas"], c["starts"], c["ends"], c["rgb"], c["beta"],
                             background=background)
    torch.testing.assert_close(out["rgb"].cpu(), ref, rtol=RTOL, atol=2e-6)


def test_training_mode_skips_clamp_and_nan_to_num(built_library):
    from uncertainty_nerf_gs_b200 import ops

    inp = synthetic.ray_samples(300, 48, seed=4)
    inp["rgb"] = inp["rgb"] * 3.0 - 1.0  # outside [0, 1]
    w = oc.get_weights(inp["density"], inp["deltas"])
    ref = oc.render_rgb(inp["rgb"], w, "last_sample", training=True)
    c = _cuda(inp)
    out = ops.composite_rays(c["density"], c["deltas"], c["starts"], c["ends"], c["rgb"], c["beta"],
                             eval_mode=False)
    torch.testing.assert_close(out["rgb"].cpu(), ref, rtol=RTOL, atol=5e-6)


@pytest.mark.parametrize("num_samples", [48, 96, 256])
def test_render_weights_prop_depth(built_library, num_samples):
    from uncertainty_nerf_gs_b200 import ops

    inp = synthetic.ray_samples(700, num_samples, seed=8)
    w = oc.get_weights(inp["density"], inp["deltas"])
    ref_depth = oc.render_depth_median(w, inp["starts"], inp["ends"])
    ref_exp = oc.render_depth_expected(w, inp["starts"], inp["ends"])
    ref_acc = oc.render_accumulation(w)
    ref_dvar = oc.depth_variance(w, inp["starts"], inp["ends"], ref_depth)
    out = ops.render_weights(w.cuda(), inp["starts"].cuda(), inp["ends"].cuda(),
                             want=("accumulation", "depth", "expected_depth", "depth_var", "depth_std"))
    # weights are given: the cumulative sums see identical inputs, the median index must match exactly
    torch.testing.assert_close(out["depth"].cpu(), ref_depth, rtol=0, atol=0)
    torch.testing.assert_close(out["accumulation"].cpu(), ref_acc, rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(out["expected_depth"].cpu(), ref_exp, rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(out["depth_var"].cpu(), ref_dvar, rtol=RTOL, atol=1e-7)


def test_laplace_outputs_unc_deterministic_density(built_library):
    from uncertainty_nerf_gs_b200.models.outputs import laplace_outputs_unc

    inp = synthetic.ray_samples(900, 48, seed=21, edge_cases=False)
    rgb_var = inp.pop("beta") * 1e-3
    ref = oc.laplace_outputs_unc(inp["density"], inp["deltas"], inp["starts"], inp["ends"], inp["rgb"], rgb_var)
    c = _cuda(inp)
    out = laplace_outputs_unc(c["density"], c["deltas"], c["starts"], c["ends"], c["rgb"], rgb_var.cuda())
    assert list(out.keys()) == list(ref.keys())
    full = dict(inp, beta=rgb_var)
    atol = dict(ATOL)
    _check(out, ref, full, ["rgb", "rgb_std", "accumulation", "depth", "depth_std", "expected_depth"], atol)


def test_laplace_sampled_density_depth(built_library):
    """depth from the mean weights of the density draws (laplace_model.py:486-521): the averaged weights
    come from the oracle here; the renderers on top of them are the CUDA ones."""
    from uncertainty_nerf_gs_b200.models.outputs import laplace_outputs_unc

    inp = synthetic.ray_samples(200, 48, seed=22, edge_cases=False)
    rgb_var = inp.pop("beta") * 1e-3
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(20, 200, 48, 1, generator=g)
    density_var = (0.1 * inp["density"]) ** 2
    ref = oc.laplace_outputs_unc(inp["density"], inp["deltas"], inp["starts"], inp["ends"], inp["rgb"], rgb_var,
                                 density_var=density_var, use_deterministic_density=False, density_noise=noise)
    std = torch.maximum(density_var.sqrt(), torch.tensor([1e-10]))
    sampled = torch.relu(inp["density"].unsqueeze(0) + std.unsqueeze(0) * noise)
    avg_w = torch.stack([oc.get_weights(s, inp["deltas"]) for s in sampled]).mean(0)
    c = _cuda(inp)
    out = laplace_outputs_unc(c["density"], c["deltas"], c["starts"], c["ends"], c["rgb"], rgb_var.cuda(),
                              averaged_weights=avg_w.cuda())
    torch.testing.assert_close(out["depth"].cpu(), ref["depth"], rtol=0, atol=0)
    torch.testing.assert_close(out["depth_std"].cpu(), ref["depth_std"], rtol=RTOL, atol=1e-6)
    torch.testing.assert_close(out["expected_depth"].cpu(), ref["expected_depth"], rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(out["accumulation"].cpu(), ref["accumulation"], rtol=RTOL, atol=2e-6)


def test_empty_and_errors(built_library):
    from uncertainty_nerf_gs_b200 import ops

    z = lambda *s: torch.zeros(*s, device="cuda")
    out = ops.composite_rays(z(0, 48), z(0, 48), z(0, 48), z(0, 48), z(0, 48, 3), z(0, 48))
    assert out["rgb"].shape == (0, 3)
    with pytest.raises(RuntimeError):
        ops.composite_rays(torch.zeros(4, 48), torch.zeros(4, 48), torch.zeros(4, 48), torch.zeros(4, 48),
                           torch.zeros(4, 48, 3))
    with pytest.raises(ValueError):
        ops.composite_rays(z(4, 48), z(4, 47), z(4, 48), z(4, 48), z(4, 48, 3))


def test_full_view_properties(built_library):
    """BASELINE size (1297x840 rays x 48): size-independent properties instead of the slow oracle --
    accumulation == sum of returned weights, rgb within [0,1], depth is one of the ray's midpoints,
    and a 2-chunk split of the same rays gives bit-identical per-ray results."""
    from uncertainty_nerf_gs_b200 import ops

    R = 1297 * 840
    inp = synthetic.ray_samples(R, 48, seed=0, device="cuda")
    o = ops.composite_rays(inp["density"], inp["deltas"], inp["starts"], inp["ends"], inp["rgb"], inp["beta"],
                           return_weights=True)
    w = o["weights"]
    torch.testing.assert_close(o["accumulation"], w.sum(dim=1), rtol=1e-5, atol=2e-6)
    assert float(o["rgb"].min()) >= 0.0 and float(o["rgb"].max()) <= 1.0
    steps = ((inp["starts"] + inp["ends"]) / 2)[..., 0]
    assert bool(((steps - o["depth"]).abs().min(dim=1).values == 0).all())
    half = (R // 2) // 8 * 8
    a = ops.composite_rays(*[inp[k][:half] for k in ("density", "deltas", "starts", "ends", "rgb", "beta")])
    b = ops.composite_rays(*[inp[k][half:] for k in ("density", "deltas", "starts", "ends", "rgb", "beta")])
    for k in ("rgb", "accumulation", "depth", "rgb_var", "depth_var"):
        assert torch.equal(torch.cat([a[k], b[k]]), o[k]), k


@pytest.mark.parametrize("num_samples,chunk", [(48, 64), (20, 0), (64, 128)])
def test_batched_members_equal_individual_calls(built_library, num_samples, chunk):
    """ub_composite_rays_batch (one memset, M kernels, one finalize launch) must return, bit for bit, what M
    separate ub_composite_rays calls return -- including the per-chunk clip bounds of every member and the
    inf-beta redo in chunks that held a NaN."""
    from uncertainty_nerf_gs_b200 import ops

    R = 333
    members = []
    for i in range(5):
        m = synthetic.ray_samples(R, num_samples, seed=40 + i, device="cuda")
        if i == 2:
            m["beta"][7, 3] = float("nan")
            m["beta"][9, 1] = float("inf")
            m["starts"][:40] += 3.0
            m["ends"][:40] += 3.0
        members.append(m)
    keys = ("density", "deltas", "starts", "ends", "rgb", "beta")
    many = ops.composite_rays_many([[m[k] for k in keys] for m in members], rays_per_chunk=chunk or None,
                                   background=(0.1, 0.5, 0.9))
    for m, got in zip(members, many):
        one = ops.composite_rays(*[m[k] for k in keys], rays_per_chunk=chunk or None, background=(0.1, 0.5, 0.9))
        for k, v in got.items():
            assert torch.equal(torch.nan_to_num(v, nan=-5.0), torch.nan_to_num(one[k], nan=-5.0)), k
    img = ops.composite_rays_many([[m[k] for k in keys] for m in members[:2]], image_hw=(9, 37))
    assert img[1]["rgb"].shape == (9, 37, 3) and img[0]["depth_std"].shape == (9, 37, 1)
    nine = ops.composite_rays_many([[members[i % 5][k] for k in keys] for i in range(9)])   # more than one C batch
    assert len(nine) == 9 and torch.equal(nine[8]["rgb"], nine[3]["rgb"])


@pytest.mark.parametrize("num_rays,num_samples,chunk", [(4096, 48, None), (1001, 48, 512), (7, 48, None), (70000, 48, 1 << 15),
                                                        (333, 32, None), (500, 64, None), (129, 96, None),
                                                        (257, 5, None), (300, 50, 64), (64, 256, None)])
def test_derived_deltas_equal_given_deltas(built_library, num_rays, num_samples, chunk):
    """``deltas=None`` (ub_composite_rays_args.deltas == NULL): the bin widths are ``ends - starts`` in float32, which is
    what nerfstudio's ``RayBundle.get_ray_samples`` stores in ``RaySamples.deltas`` -- every output bit for bit the one
    of the call that is handed that difference, on the TMA path, the generic path and through the NaN-guard redo of the
    finalize pass; and the oracle on the same inputs within the usual tolerances."""
    from uncertainty_nerf_gs_b200 import ops
    from uncertainty_nerf_gs_b200.models.outputs import active_nerfacto_outputs, active_nerfacto_outputs_many

    inp = synthetic.ray_samples(num_rays, num_samples, seed=7 * num_rays + num_samples)
    inp["deltas"] = inp["ends"] - inp["starts"]
    if chunk is not None:
        inp["beta"][::5, 0] = float("inf")      # inf betas in chunks that also hold NaNs: the finalize pass recomputes them
    dev = _cuda(inp)
    given = active_nerfacto_outputs(**dev, return_weights=True, rays_per_chunk=chunk)
    none = dict(dev, deltas=None)
    derived = active_nerfacto_outputs(**none, return_weights=True, rays_per_chunk=chunk)
    for k in KEYS + ["weights"]:
        assert torch.equal(given[k].view(torch.int32), derived[k].view(torch.int32)), k
    many = active_nerfacto_outputs_many([{k: v for k, v in dev.items() if k != "deltas"}, dev], rays_per_chunk=chunk)
    for k in KEYS:
        assert torch.equal(many[0][k].view(torch.int32), given[k].view(torch.int32)), k
        assert torch.equal(many[1][k].view(torch.int32), given[k].view(torch.int32)), k
    raw = ops.composite_rays(dev["density"], None, dev["starts"], dev["ends"], dev["rgb"], dev["beta"], beta_mode="raw",
                             background=(0.25, 0.5, 0.75), eval_mode=False)
    raw_given = ops.composite_rays(dev["density"], dev["deltas"], dev["starts"], dev["ends"], dev["rgb"], dev["beta"],
                                   beta_mode="raw", background=(0.25, 0.5, 0.75), eval_mode=False)
    for k in KEYS:
        assert torch.equal(raw[k].view(torch.int32), raw_given[k].view(torch.int32)), k
    if chunk is None:
        ref = oc.active_nerfacto_outputs(**inp)
        _check(derived, ref, inp, KEYS, ATOL)
