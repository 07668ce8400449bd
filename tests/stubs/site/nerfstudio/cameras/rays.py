"""``Frustums`` / ``RaySamples`` / ``RayBundle`` with the members the reference touches
(activenerfacto_model.py:89-112, laplace_model.py:210-231, 432-441, 459-521)."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Optional

import torch


def _slice(obj, idx):
    kw = {}
    for f in fields(obj):
        v = getattr(obj, f.name)
        if torch.is_tensor(v):
            kw[f.name] = v[idx]
        elif hasattr(v, "__dataclass_fields__"):
            kw[f.name] = _slice(v, idx)
        else:
            kw[f.name] = v
    return type(obj)(**kw)


@dataclass
class Frustums:
    origins: Optional[torch.Tensor] = None       # [..., 3]
    directions: Optional[torch.Tensor] = None    # [..., 3]
    starts: Optional[torch.Tensor] = None        # [..., 1]
    ends: Optional[torch.Tensor] = None          # [..., 1]
    pixel_area: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.starts.shape[:-1]

    def get_positions(self):
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        return pos


@dataclass
class RaySamples:
    frustums: Frustums = None
    camera_indices: Optional[torch.Tensor] = None
    deltas: Optional[torch.Tensor] = None        # [..., S, 1]
    spacing_starts: Optional[torch.Tensor] = None
    spacing_ends: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.frustums.shape

    def get_weights(self, densities: torch.Tensor) -> torch.Tensor:
        """alpha-compositing weights from densities ``[..., S, 1]`` (nerfstudio ``cameras/rays.py``)."""
        delta_density = self.deltas * densities
        alphas = 1 - torch.exp(-delta_density)
        transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
        transmittance = torch.cat(
            [torch.zeros((*transmittance.shape[:1], 1, 1), device=densities.device), transmittance], dim=-2)
        transmittance = torch.exp(-transmittance)
        weights = alphas * transmittance
        return torch.nan_to_num(weights)


@dataclass
class RayBundle:
    origins: torch.Tensor = None                 # [..., 3] (image shaped for a camera ray bundle)
    directions: Optional[torch.Tensor] = None
    pixel_area: Optional[torch.Tensor] = None
    camera_indices: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None
    metadata: Optional[dict] = None
    times: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def __len__(self) -> int:
        n = 1
        for s in self.origins.shape[:-1]:
            n *= int(s)
        return n

    def flatten(self) -> "RayBundle":
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.reshape(-1, v.shape[-1]) if torch.is_tensor(v) else v
        return RayBundle(**kw)

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return _slice(self.flatten(), slice(start_idx, end_idx))


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
