"""GPU parity for the two "next" rows already wired through the kernels: depth-uncertainty scoring with
ragged (masked) segments (eval_uncertainty.py:415-644) and the nerfacto-laplace sampled-density depth
(laplace_model.py:486-521) with the draws passed in."""
import numpy as np
import pytest
import torch

from oracle import compositing as oc, metrics as om
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _depth_view(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(h, w, generator=g) * 8.0 + 0.5
    gt[torch.rand(h, w, generator=g) < 0.3] = 0.0                      # invalid pixels -> ragged segment
    depth = (gt + 0.3 * torch.randn(h, w, generator=g)).clamp(min=-1.0) / 2.5
    depth[0, :5] = 100.0                                               # exercises the clamp to max gt
    std = torch.clamp(0.2 * torch.rand(h, w, generator=g), min=0.05)   # tie group at the floor
    return depth[..., None], std[..., None], gt


def test_score_depth_batch_vs_oracle(built_library):
    from uncertainty_nerf_gs_b200.metrics import score_depth_batch

    views = [_depth_view(37, 53, s) for s in range(3)]
    scales = [2.5, 2.4, 2.6]
    outs = score_depth_batch(torch.stack([v[0] for v in views]).cuda(), torch.stack([v[1] for v in views]).cuda(),
                             torch.stack([v[2] for v in views]).cuda(), scales, min_depth_std_for_nll=1.0)
    for (d, s, g), a, out in zip(views, scales, outs):
        ref = om.unc_metrics_depth(d, s, g, a, min_depth_std_for_nll=1.0)
        assert np.array_equal(out["coverage_values"], ref["coverage_values"])
        for k in ("err_mae", "err_mse", "err_rmse", "err_var_mae", "err_var_mse", "err_var_rmse", "avg_length_values"):
            np.testing.assert_allclose(out[k], ref[k], rtol=RTOL, atol=0)
        for k in ("ause_mae", "ause_mse", "ause_rmse", "avg_var", "auc_abs_error_values", "auc_length_values",
                  "auc_neg_error_values"):
            np.testing.assert_allclose(out[k], ref[k], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(out["nll_depth"], ref["nll_depth"], rtol=RTOL)
        np.testing.assert_allclose(out["mse_mean"], float(ref["mse"].mean()), rtol=RTOL)


def test_score_depth_batch_resizes_renders_like_the_reference(built_library):
    """Renders of shape [H-1, W-1] (splatfacto's depth) against a [H, W] ground truth: eval_uncertainty.py:442-452
    resizes them bilinearly before scoring; the oracle's resize is pinned to that code (tests/test_oracle_pinned.py).
    The interpolation runs in torch on either side (CUDA vs CPU kernels: equal to float32 rounding), so the
    coverage may differ for an element sitting on an interval end: compared to 2e-3 of the pixel count."""
    from uncertainty_nerf_gs_b200.metrics import resize_like_reference, score_depth_batch

    views = [_depth_view(37, 53, s) for s in range(2)]
    small = [(d[:-1, :-1], s[:-1, :-1], g) for d, s, g in views]
    scales = [2.5, 2.4]
    got = resize_like_reference(torch.stack([v[0][..., 0] for v in small]).cuda(), (37, 53)).cpu()
    want = torch.stack([om.resize_like_reference(v[0][..., 0], (37, 53)) for v in small])
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
    outs = score_depth_batch(torch.stack([v[0] for v in small]).cuda(), torch.stack([v[1] for v in small]).cuda(),
                             torch.stack([v[2] for v in small]).cuda(), scales, min_depth_std_for_nll=1.0)
    for (d, s, g), a, out in zip(small, scales, outs):
        ref = om.unc_metrics_depth(d, s, g, a, min_depth_std_for_nll=1.0)
        cov_ref = np.asarray(ref["coverage_values"], dtype=np.float64)       # fractions of the valid pixels
        np.testing.assert_allclose(np.asarray(out["coverage_values"], dtype=np.float64), cov_ref, rtol=0,
                                   atol=2e-3 * max(1.0, float(cov_ref.max())))
        for k in ("err_mae", "err_mse", "err_rmse", "err_var_mae", "err_var_mse", "err_var_rmse"):
            np.testing.assert_allclose(out[k], ref[k], rtol=1e-4, atol=1e-7)
        for k in ("ause_mae", "ause_mse", "ause_rmse", "avg_var"):
            np.testing.assert_allclose(out[k], ref[k], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(out["nll_depth"], ref["nll_depth"], rtol=1e-4)


@pytest.mark.parametrize("num_samples,draws", [(48, 20), (96, 7), (5, 3)])
def test_average_sampled_weights_with_given_noise(built_library, num_samples, draws):
    from uncertainty_nerf_gs_b200 import ops

    R = 300
    inp = synthetic.ray_samples(R, num_samples, seed=3, edge_cases=False)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(draws, R, num_samples, 1, generator=g)
    dvar = (0.2 * inp["density"]) ** 2
    dvar[7, 2] = float("nan")                                          # NaN std -> 1e-10 (:495-496)
    std = torch.nan_to_num(torch.maximum(dvar.sqrt(), torch.tensor([1e-10])), nan=1e-10)
    sampled = torch.relu(inp["density"].unsqueeze(0) + std.unsqueeze(0) * noise)
    ref = torch.stack([oc.get_weights(s, inp["deltas"]) for s in sampled]).mean(0)
    out = ops.average_sampled_weights(inp["density"].cuda(), dvar.cuda(), inp["deltas"].cuda(), draws,
                                      noise=noise[..., 0].cuda())
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=3e-7)


def test_laplace_outputs_unc_sampled_density_end_to_end(built_library):
    from uncertainty_nerf_gs_b200.models.outputs import laplace_outputs_unc

    R, S, K = 400, 48, 16
    inp = synthetic.ray_samples(R, S, seed=9, edge_cases=False)
    rgb_var = inp.pop("beta") * 1e-3
    noise = torch.randn(K, R, S, 1, generator=torch.Generator().manual_seed(2))
    dvar = (0.1 * inp["density"]) ** 2
    ref = oc.laplace_outputs_unc(inp["density"], inp["deltas"], inp["starts"], inp["ends"], inp["rgb"], rgb_var,
                                 density_var=dvar, use_deterministic_density=False, density_noise=noise)
    c = {k: v.cuda() for k, v in inp.items()}
    out = laplace_outputs_unc(c["density"], c["deltas"], c["starts"], c["ends"], c["rgb"], rgb_var.cuda(),
                              density_var=dvar.cuda(), density_noise=noise[..., 0].cuda(), num_draws=K)
    same = torch.isclose(out["depth"].cpu(), ref["depth"], rtol=RTOL, atol=0)[:, 0]
    assert float(same.float().mean()) > 0.995          # the median index may move on borderline rays
    torch.testing.assert_close(out["accumulation"].cpu(), ref["accumulation"], rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(out["expected_depth"].cpu(), ref["expected_depth"], rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(out["rgb"].cpu(), ref["rgb"], rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(out["depth_std"].cpu()[same], ref["depth_std"][same], rtol=RTOL, atol=1e-6)


def test_philox_draws_are_statistically_sane(built_library):
    """Without passed-in noise the kernel draws N(0,1) itself: compare the mean weights with a large-K
    torch estimate (statistical parity only)."""
    from uncertainty_nerf_gs_b200 import ops

    R, S = 256, 48
    inp = synthetic.ray_samples(R, S, seed=4, edge_cases=False)
    dvar = (0.3 * inp["density"]) ** 2
    out = ops.average_sampled_weights(inp["density"].cuda(), dvar.cuda(), inp["deltas"].cuda(), 400, seed=7).cpu()
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(400, R, S, 1, generator=g)
    sampled = torch.relu(inp["density"].unsqueeze(0) + dvar.sqrt().unsqueeze(0) * noise)
    ref = torch.stack([oc.get_weights(s, inp["deltas"]) for s in sampled]).mean(0)
    assert float((out - ref).abs().mean()) < 2e-3
    assert abs(float(out.sum() - ref.sum())) / float(ref.sum()) < 5e-3


def test_depth_prepare_is_the_reference_torch_lines(built_library):
    """ub_depth_prepare against the reference's own per-view torch lines (scale, masked-assignment clamps, boolean mask),
    bit for bit: ragged views, a view without any valid pixel, a NaN in a ground truth, sizes that are not a
    multiple of the block, pixel order preserved."""
    from uncertainty_nerf_gs_b200 import ops

    g = torch.Generator().manual_seed(5)
    for n in (1, 1023, 1024, 5000):
        b = 4
        gt = torch.rand(b, n, generator=g) * 8 + 0.5
        gt[torch.rand(b, n, generator=g) < 0.4] = 0.0
        gt[1] = 0.0                                             # a view with nothing to score
        if n > 10:
            gt[2, 7] = float("nan")                             # MAX_DEPTH = NaN: `depth > NaN` is false, nothing is clamped above
            gt[3, 3] = -2.0
        depth = torch.randn(b, n, generator=g) * 3
        depth[0, 0] = float("nan")
        depth[3, n // 2] = 1e6
        std = torch.rand(b, n, generator=g)
        scales = [2.5, 1.0, 0.7, 3.25]
        pred, sd, gg, lens = ops.depth_prepare(depth.cuda(), std.cuda(), gt.cuda(), scales)
        want_p, want_s, want_g, want_l = [], [], [], []
        for i in range(b):
            max_d = gt[i].max().float()
            d = scales[i] * depth[i]
            mask = gt[i] > 0
            dm = d[mask]                                        # eval_uncertainty.py:552-560, the reference's own lines:
            dm[dm < 1e-3] = 1e-3                                #   depth[depth < MIN_DEPTH] = MIN_DEPTH
            dm[dm > max_d] = max_d                              #   depth[depth > MAX_DEPTH] = MAX_DEPTH
            want_p.append(dm); want_s.append((scales[i] * std[i])[mask]); want_g.append(gt[i][mask]); want_l.append(int(mask.sum()))
        assert lens == want_l
        for got, want in ((pred, want_p), (sd, want_s), (gg, want_g)):
            w = torch.cat(want)
            assert torch.equal(torch.nan_to_num(got.cpu(), nan=-9.0), torch.nan_to_num(w, nan=-9.0))
