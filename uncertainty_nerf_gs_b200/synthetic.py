"""Seeded synthetic inputs of the shapes BASELINE.json names (no datasets / checkpoints offline).

The field networks, the proposal sampler and ``gsplat.project_gaussians`` are producers
upstream of the hot path and out of scope, so the path is driven with tensors of the shapes
and value ranges those producers emit.  Everything is generated with a
``torch.Generator`` on the requested device: ``"cpu"`` for parity tests (bit-identical inputs
for the CUDA path and the oracle), ``"cuda"`` for the full-size benchmark.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor

NERFACTO_SAMPLES_PER_RAY = 48          # nerfstudio nerfacto default, final proposal level
EVAL_RAYS_PER_CHUNK = 1 << 15          # reference activenerfacto_config.py:38
MIP360_HW = (840, 1297)                # H, W of a Mip-NeRF-360 view at the reference's downscale
BLENDER_HW = (800, 800)


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def ray_samples(num_rays: int, num_samples: int = NERFACTO_SAMPLES_PER_RAY, seed: int = 0,
                device="cpu", edge_cases: bool = True, near: float = 0.05) -> Dict[str, Tensor]:
    """Per-sample field outputs for ``num_rays`` rays: density, deltas, starts, ends, rgb, beta
    (``[R, S, C]`` float32).  With ``edge_cases``: 1 % empty rays (accumulation 0), 1 % rays
    saturating within the first three samples, 0.1 % NaN in beta (exercises the reference's
    ``nan_to_num`` guard, activenerfacto_model.py:105-106)."""
    g = _gen(seed, device)
    R, S = num_rays, num_samples
    kw = dict(generator=g, device=device, dtype=torch.float32)
    deltas = torch.rand(R, S, 1, **kw) * (0.1 - 1e-3) + 1e-3
    starts = torch.cumsum(deltas, dim=1) - deltas + near
    ends = starts + deltas
    density = torch.exp(torch.clamp(2.0 * torch.randn(R, S, 1, **kw), -15.0, 15.0))
    density = density * (torch.rand(R, S, 1, **kw) < 0.9).float()
    rgb = torch.sigmoid(torch.randn(R, S, 3, **kw))
    beta = torch.nn.functional.softplus(torch.randn(R, S, 1, **kw)) + 0.01
    if edge_cases:
        pick = torch.rand(R, **kw)
        density[pick < 0.01] = 0.0
        sat = (pick >= 0.01) & (pick < 0.02)
        density[sat, :3] = 1e4
        nan_mask = torch.rand(R, S, 1, **kw) < 1e-3
        beta = torch.where(nan_mask, torch.full_like(beta, float("nan")), beta)
    return {"density": density, "deltas": deltas, "starts": starts, "ends": ends, "rgb": rgb, "beta": beta}


def scoring_image(height: int, width: int, seed: int = 0, device="cpu", std_floor: float = 0.03
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """``(rgb_pred [H,W,3], rgb_std [H,W,1], rgb_gt [H,W,3])``.  ``std = max(floor, 0.1 U)`` forces a
    large tie group at the floor, so rankings only match under the stable tie-break."""
    g = _gen(seed, device)
    kw = dict(generator=g, device=device, dtype=torch.float32)
    pred = torch.rand(height, width, 3, **kw)
    std = torch.clamp(0.1 * torch.rand(height, width, 1, **kw), min=std_floor)
    gt = torch.clamp(pred + std * torch.randn(height, width, 3, **kw), 0.0, 1.0)
    return pred, std, gt


def member_renders(num_members: int, height: int, width: int, seed: int = 0, device="cpu",
                   with_pred_std: bool = False) -> list:
    """``num_members`` per-view output dicts shaped like nerfacto eval outputs (``[H, W, C]``), as the
    ensemble / MC-dropout reduce receives them.  ``with_pred_std`` adds the active-* keys so the
    ensemble's aleatoric + epistemic branch is exercised."""
    outs = []
    base = _gen(seed, device)
    kw0 = dict(generator=base, device=device, dtype=torch.float32)
    mean_rgb = torch.rand(height, width, 3, **kw0)
    mean_depth = torch.rand(height, width, 1, **kw0) * 5.0 + 0.5
    for m in range(num_members):
        g = _gen(seed * 1000 + 17 + m, device)
        kw = dict(generator=g, device=device, dtype=torch.float32)
        rgb = torch.clamp(mean_rgb + 0.05 * torch.randn(height, width, 3, **kw), 0.0, 1.0)
        depth = mean_depth + 0.1 * torch.randn(height, width, 1, **kw)
        out = {
            "rgb": rgb,
            "accumulation": torch.rand(height, width, 1, **kw),
            "depth": depth,
            "expected_depth": depth + 0.01 * torch.randn(height, width, 1, **kw),
        }
        if with_pred_std:
            out["rgb_var"] = 1e-3 + 1e-2 * torch.rand(height, width, 1, **kw)
            out["rgb_std"] = out["rgb_var"].sqrt()
            out["depth_var"] = 1e-5 + 1e-2 * torch.rand(height, width, 1, **kw)
            out["depth_std"] = out["depth_var"].sqrt()
        out["prop_depth_0"] = depth + 0.05 * torch.randn(height, width, 1, **kw)
        out["prop_depth_1"] = depth + 0.02 * torch.randn(height, width, 1, **kw)
        outs.append(out)
    return outs


def laplace_head(num_points: int, hidden: int = 64, out_dim: int = 3, n_samples: int = 100,
                 seed: int = 0, device="cpu") -> Dict[str, Tensor]:
    """Inputs of the last-layer Laplace MC moments (reference laplace_field.py:528-568): features
    ``x [P, hidden]``, MAP parameters ``mu_q`` (weight row-major then bias, the order of
    ``parameters_to_vector``), diagonal GGN and the standard-normal draws, which are passed in
    so that the CUDA path and the oracle see identical samples."""
    g = _gen(seed, device)
    kw = dict(generator=g, device=device, dtype=torch.float32)
    n_params = hidden * out_dim + out_dim
    return {
        "x": torch.relu(torch.randn(num_points, hidden, **kw)),
        "mu_q": 0.3 * torch.randn(n_params, **kw),
        "ggn": torch.exp(3.0 + 2.0 * torch.randn(n_params, **kw)),
        "eps_draws": torch.randn(n_samples, n_params, **kw),
    }


def splat_scene(num_gaussians: int, height: int, width: int, seed: int = 0, device="cpu",
                mean_scale_px: float = 3.0) -> Dict[str, Tensor]:
    """Projected 2-D Gaussians as ``gsplat.project_gaussians`` would emit them (reference
    activesplatfacto_model.py:221-234): centres, conics, radii, opacities, depths, colours, beta."""
    g = _gen(seed, device)
    kw = dict(generator=g, device=device, dtype=torch.float32)
    G = num_gaussians
    xys = torch.stack([torch.rand(G, **kw) * (width + 32) - 16, torch.rand(G, **kw) * (height + 32) - 16], -1)
    scale = torch.exp(torch.log(torch.tensor(mean_scale_px, device=device)) + 0.7 * torch.randn(G, 2, **kw))
    theta = torch.rand(G, **kw) * 3.14159265
    c, s = torch.cos(theta), torch.sin(theta)
    sx2, sy2 = scale[:, 0] ** 2, scale[:, 1] ** 2
    cov_a = c * c * sx2 + s * s * sy2
    cov_b = c * s * (sx2 - sy2)
    cov_c = s * s * sx2 + c * c * sy2
    det = cov_a * cov_c - cov_b * cov_b
    conics = torch.stack([cov_c / det, -cov_b / det, cov_a / det], -1)
    radii = torch.ceil(3.0 * scale.max(dim=-1).values).to(torch.int32)
    return {
        "xys": xys, "conics": conics, "radii": radii,
        "opacities": torch.sigmoid(1.5 * torch.randn(G, 1, **kw)),
        "depths": torch.rand(G, **kw) * 9.9 + 0.1,
        "rgbs": torch.rand(G, 3, **kw),
        "betas": torch.nn.functional.softplus(torch.randn(G, 1, **kw)) + 0.01,
    }


def gaussians_3d(num_gaussians: int, height: int, width: int, seed: int = 0, device="cpu", sh_degree: int = 3
                 ) -> Dict[str, Tensor]:
    """A 3-D Gaussian scene in front of (and partly behind / beside) a pinhole camera, with the inputs the
    reference hands to ``project_gaussians`` / ``spherical_harmonics`` (activesplatfacto_model.py:205-249):
    ``means, scales`` (already exp'ed), ``quats`` (unnormalised), ``viewmat [3,4]``, intrinsics, SH coefficients."""
    g = _gen(seed, device)
    kw = dict(generator=g, device=device, dtype=torch.float32)
    G = num_gaussians
    means = torch.stack([torch.rand(G, **kw) * 8 - 4, torch.rand(G, **kw) * 6 - 3, torch.rand(G, **kw) * 9 - 1], -1)
    scales = torch.exp(torch.randn(G, 3, **kw) * 0.6 - 3.0)
    quats = torch.randn(G, 4, **kw)
    # camera: small rotation about y and x, translated back along z
    ay, ax = 0.15, -0.08
    ry = torch.tensor([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    rx = torch.tensor([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    rot = (rx @ ry).to(torch.float32)
    viewmat = torch.cat([rot, torch.tensor([[0.1], [-0.2], [1.5]])], dim=1).to(device)
    k = (sh_degree + 1) ** 2
    return {
        "means": means, "scales": scales, "quats": quats, "viewmat": viewmat,
        "fx": 0.9 * width, "fy": 0.95 * width, "cx": width / 2 + 0.3, "cy": height / 2 - 0.4,
        "sh_coeffs": 0.5 * torch.randn(G, k, 3, **kw),
        "camera_position": -(rot.T @ torch.tensor([0.1, -0.2, 1.5])).to(device),
        "opacities": torch.sigmoid(1.5 * torch.randn(G, 1, **kw)),
        "betas": torch.nn.functional.softplus(torch.randn(G, 1, **kw)) + 0.01,
    }
