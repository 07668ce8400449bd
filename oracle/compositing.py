"""Oracle: per-ray front-to-back compositing with the variance term (torch, CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED to the reference's own
code: ``tests/test_oracle_pinned.py`` executes ``ComputeWeightsModule`` / ``SumModule``
(``laplace_model.py:47-62,102-107``), ``ActiveNerfactoModel.get_outputs`` + the inherited chunk loop
and ``NerfactoLaplaceModel.get_outputs_unc`` unmodified (``oracle/ref_exec.py``) and demands bit
equality with the functions below (also against ``tests/golden/ref_composite.npz``).  The renderers
themselves live in nerfstudio 1.1.0 (``model_components/renderers.py``,
``cameras/rays.py:RaySamples.get_weights``; the version ``/root/reference/README.md:23`` installs; not
vendored, not installable here): they are restated from the published code, in ``tests/stubs/site/nerfstudio``
for the reference to call and here for the tests.  The anchors inside the reference are cited per function.

Tensor conventions follow the reference: ``[R, S, C]`` with a trailing channel
axis (``C = 1`` for density / deltas / starts / ends / beta, ``C = 3`` for rgb).
Everything runs in float32 on CPU, i.e. exactly the "reference torch path on
CPU" that BASELINE.json names; in particular ``torch.cumsum`` on CPU float32
accumulates in float64 and rounds every prefix to float32, which decides the
median-depth *index*.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

import torch

Tensor = torch.Tensor
Background = Union[str, Tensor, Sequence[float]]


def get_weights(densities: Tensor, deltas: Tensor) -> Tensor:
    """Volume-rendering weights ``w_i = alpha_i * T_i``.

    Follows the in-repo restatement of nerfstudio's ``RaySamples.get_weights``:
    reference ``nerfuncertainty/models/laplace/laplace_model.py:47-62``
    (``ComputeWeightsModule.forward``); called at
    ``activenerfacto_model.py:94`` and ``laplace_model.py:218,471``.
    """
    dd = deltas * densities
    alphas = 1 - torch.exp(-dd)
    acc = torch.cumsum(dd[..., :-1, :], dim=-2)
    acc = torch.cat([torch.zeros_like(acc[..., :1, :]), acc], dim=-2)
    trans = torch.exp(-acc)
    return torch.nan_to_num(alphas * trans)


def midpoints(starts: Tensor, ends: Tensor) -> Tensor:
    """``steps = (starts + ends) / 2`` -- ``activenerfacto_model.py:111``."""
    return (starts + ends) / 2


def render_accumulation(weights: Tensor) -> Tensor:
    """nerfstudio ``AccumulationRenderer``; call site ``activenerfacto_model.py:102``."""
    return torch.sum(weights, dim=-2)


def render_rgb(rgb: Tensor, weights: Tensor, background_color: Background = "last_sample",
               training: bool = False) -> Tensor:
    """nerfstudio ``RGBRenderer.forward`` (eval mode by default).

    Call sites: ``activenerfacto_model.py:98``, ``laplace_model.py:222,475``.  The
    plain weighted sum is restated in-repo as ``SumModule`` (``laplace_model.py:102-107``).
    Eval mode: ``nan_to_num`` on the input colours and ``clamp_(0, 1)`` on the output.
    ``background_color``: ``"last_sample"`` (nerfacto default), ``"random"`` (returns the
    unblended sum), or a fixed RGB triple.
    """
    if not training:
        rgb = torch.nan_to_num(rgb)
    comp = torch.sum(weights * rgb, dim=-2)
    acc = torch.sum(weights, dim=-2)
    if isinstance(background_color, str) and background_color == "random":
        out = comp
    else:
        if isinstance(background_color, str):
            if background_color != "last_sample":
                raise ValueError(f"unsupported background {background_color!r}")
            bg = rgb[..., -1, :]
        else:
            bg = torch.as_tensor(background_color, dtype=comp.dtype).expand(comp.shape)
        out = comp + bg * (1.0 - acc)
    if not training:
        out = torch.clamp(out, min=0.0, max=1.0)
    return out


def render_depth_median(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """nerfstudio ``DepthRenderer(method="median")``; call sites
    ``activenerfacto_model.py:99-100,150-151``, ``laplace_model.py:509``."""
    steps = midpoints(starts, ends)
    cw = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * 0.5
    idx = torch.searchsorted(cw, split, side="left")
    idx = torch.clamp(idx, 0, steps.shape[-2] - 1)
    return torch.gather(steps[..., 0], dim=-1, index=idx)


def render_depth_expected(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """nerfstudio ``DepthRenderer(method="expected")``; call site
    ``activenerfacto_model.py:101``.  NB the clip bounds are the min/max of
    ``steps`` over the *whole tensor passed in* (one eval chunk in the reference)."""
    eps = 1e-10
    steps = midpoints(starts, ends)
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + eps)
    return torch.clip(depth, steps.min(), steps.max())


def render_uncertainty(betas: Tensor, weights: Tensor) -> Tensor:
    """nerfstudio ``UncertaintyRenderer``: ``sum(weights * betas, dim=-2)``.
    The reference always passes ``weights=weights**2`` (``activenerfacto_model.py:107``,
    ``laplace_model.py:478-479``)."""
    return torch.sum(weights * betas, dim=-2)


def depth_variance(weights: Tensor, starts: Tensor, ends: Tensor, depth: Tensor) -> Tensor:
    """``sum_i w_i (steps_i - depth)^2 + 1e-5`` -- ``activenerfacto_model.py:111-112``,
    ``laplace_model.py:513-514``."""
    steps = midpoints(starts, ends)
    return torch.sum(weights * (steps - depth.unsqueeze(-1)).pow(2), dim=-2) + 1e-5


def active_nerfacto_outputs(density: Tensor, deltas: Tensor, starts: Tensor, ends: Tensor,
                            rgb: Tensor, beta: Tensor,
                            background_color: Background = "last_sample") -> Dict[str, Tensor]:
    """One chunk of ``ActiveNerfactoModel.get_outputs`` in eval mode, downstream of the
    field: ``activenerfacto_model.py:94-127`` (keys in the reference's insertion order)."""
    weights = get_weights(density, deltas)
    out_rgb = render_rgb(rgb, weights, background_color)
    depth = render_depth_median(weights, starts, ends)
    expected = render_depth_expected(weights, starts, ends)
    acc = render_accumulation(weights)
    if torch.isnan(beta).any():
        beta = torch.nan_to_num(beta, 0.0)
    rgb_var = render_uncertainty(beta, weights ** 2)
    dvar = depth_variance(weights, starts, ends, depth)
    return {
        "rgb": out_rgb,
        "accumulation": acc,
        "depth": depth,
        "expected_depth": expected,
        "density": density,
        "rgb_var": rgb_var,
        "rgb_std": rgb_var.sqrt(),
        "depth_var": dvar,
        "depth_std": dvar.sqrt(),
    }


def laplace_outputs_unc(density: Tensor, deltas: Tensor, starts: Tensor, ends: Tensor,
                        rgb: Tensor, rgb_var: Tensor, density_var: Optional[Tensor] = None,
                        use_deterministic_density: bool = True,
                        density_noise: Optional[Tensor] = None,
                        background_color: Background = "last_sample",
                        density_draws: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """One chunk of ``NerfactoLaplaceModel.get_outputs_unc`` downstream of the field:
    ``laplace_model.py:471-530``.

    rgb / rgb_std use the deterministic weights (``:471-480``).  With
    ``use_deterministic_density=False`` the depth / depth_std / expected_depth /
    accumulation use the mean of the weights of 100 density draws (``:486-521``);
    the draws are ``relu(density + density_std * density_noise[k])`` with the standard
    normal ``density_noise [K, R, S, 1]`` passed in so that oracle and GPU see identical
    samples (the reference draws them from torch's global generator).  ``density_draws [K, R, S, 1]`` passes
    the draws of ``Normal(density, density_std).sample((K,))`` themselves (before the relu), which is how the
    reference-executed goldens are replayed bit for bit.
    """
    weights = get_weights(density, deltas)
    out_rgb = render_rgb(rgb, weights, background_color)
    var = render_uncertainty(rgb_var, weights ** 2)
    rgb_std = torch.sqrt(var)
    if not use_deterministic_density:
        assert density_var is not None and (density_noise is not None or density_draws is not None)
        density_std = density_var.sqrt()
        density_std = torch.maximum(density_std, torch.tensor([1e-10]))
        if torch.isnan(density_std).any():
            density_std = torch.nan_to_num(density_std, nan=1e-10)
        if density_draws is None:
            density_draws = density.unsqueeze(0) + density_std.unsqueeze(0) * density_noise
        sampled = torch.relu(density_draws)
        sampled_w = torch.stack([get_weights(s, deltas) for s in sampled], dim=0)
        weights = sampled_w.mean(dim=0)
    depth = render_depth_median(weights, starts, ends)
    dvar = depth_variance(weights, starts, ends, depth)
    if torch.isnan(dvar).any() or torch.isinf(dvar).any():
        raise RuntimeError("depth_var has Nans")
    expected = render_depth_expected(weights, starts, ends)
    acc = render_accumulation(weights)
    return {
        "rgb": out_rgb,
        "rgb_std": rgb_std,
        "accumulation": acc,
        "depth": depth,
        "depth_std": torch.sqrt(dvar),
        "expected_depth": expected,
    }


def render_in_chunks(fn, num_rays_per_chunk: int, *ray_tensors: Tensor, **kwargs) -> Dict[str, Tensor]:
    """The reference's eval chunk loop (nerfstudio ``get_outputs_for_camera_ray_bundle``,
    restated in-repo at ``laplace_model.py:282-297``): slice rays row-major into chunks of
    ``eval_num_rays_per_chunk`` (1 << 15, ``activenerfacto_config.py:38``), run the per-chunk
    function and concatenate.  Per-chunk global reductions (``steps.min()``, the ``isnan().any()``
    guard) therefore see one chunk at a time."""
    n = ray_tensors[0].shape[0]
    pieces: Dict[str, list] = {}
    for lo in range(0, n, num_rays_per_chunk):
        part = fn(*[None if t is None else t[lo:lo + num_rays_per_chunk] for t in ray_tensors], **kwargs)
        for k, v in part.items():
            pieces.setdefault(k, []).append(v)
    return {k: torch.cat(v) for k, v in pieces.items()}
