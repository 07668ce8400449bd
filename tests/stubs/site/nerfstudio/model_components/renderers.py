"""The renderers the reference instantiates / calls (activenerfacto_model.py:81, 98-107; laplace_model.py:189,
222-231, 475-521), restated from nerfstudio 1.1.0 ``model_components/renderers.py``."""
from __future__ import annotations

import contextlib
from typing import Optional, Union

import torch
from torch import nn

BACKGROUND_COLOR_OVERRIDE: Optional[torch.Tensor] = None
_NAMED = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0), "red": (1.0, 0.0, 0.0), "green": (0.0, 1.0, 0.0),
          "blue": (0.0, 0.0, 1.0)}


@contextlib.contextmanager
def background_color_override_context(mode):
    global BACKGROUND_COLOR_OVERRIDE
    old = BACKGROUND_COLOR_OVERRIDE
    try:
        BACKGROUND_COLOR_OVERRIDE = mode
        yield
    finally:
        BACKGROUND_COLOR_OVERRIDE = old


class RGBRenderer(nn.Module):
    def __init__(self, background_color: Union[str, torch.Tensor] = "random"):
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb, weights, background_color="random"):
        comp_rgb = torch.sum(weights * rgb, dim=-2)
        accumulated_weight = torch.sum(weights, dim=-2)
        if BACKGROUND_COLOR_OVERRIDE is not None:
            background_color = BACKGROUND_COLOR_OVERRIDE
        if isinstance(background_color, str) and background_color == "random":
            return comp_rgb
        if isinstance(background_color, str) and background_color == "last_sample":
            background_color = rgb[..., -1, :]
        elif isinstance(background_color, str):
            background_color = torch.tensor(_NAMED[background_color], device=comp_rgb.device)
        background_color = torch.as_tensor(background_color, dtype=comp_rgb.dtype).expand(comp_rgb.shape).to(comp_rgb.device)
        return comp_rgb + background_color * (1.0 - accumulated_weight)

    def forward(self, rgb, weights, ray_indices=None, num_rays=None, background_color=None):
        if background_color is None:
            background_color = self.background_color
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        rgb = self.combine_rgb(rgb, weights, background_color=background_color)
        if not self.training:
            torch.clamp_(rgb, min=0.0, max=1.0)
        return rgb


class AccumulationRenderer(nn.Module):
    @classmethod
    def forward(cls, weights, ray_indices=None, num_rays=None):
        return torch.sum(weights, dim=-2)


class DepthRenderer(nn.Module):
    def __init__(self, method: str = "median"):
        super().__init__()
        self.method = method

    def forward(self, weights, ray_samples, ray_indices=None, num_rays=None):
        if self.method == "median":
            steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
            cumulative_weights = torch.cumsum(weights[..., 0], dim=-1)
            split = torch.ones((*weights.shape[:-2], 1), device=weights.device) * 0.5
            median_index = torch.searchsorted(cumulative_weights, split, side="left")
            median_index = torch.clamp(median_index, 0, steps.shape[-2] - 1)
            return torch.gather(steps[..., 0], dim=-1, index=median_index)
        if self.method == "expected":
            eps = 1e-10
            steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
            depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + eps)
            return torch.clip(depth, steps.min(), steps.max())
        raise NotImplementedError(self.method)


class UncertaintyRenderer(nn.Module):
    @classmethod
    def forward(cls, betas, weights):
        return torch.sum(weights * betas, dim=-2)


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
