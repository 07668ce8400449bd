"""Stand-in *producers* for the hot path (proposal sampler, fields, member models, eval dataloader) -- TEST
INFRASTRUCTURE ONLY.  They hand pre-made tensors to whichever ``get_outputs*`` implementation is under test, so
that the reference's own methods (dev container, CPU, ``oracle/ref_exec.py``) and the plugin replacements
(GPU box, ``models/nerfstudio_plugin.py``) are driven by bit-identical inputs.

Needs ``tests/stubs/ub_stubs.install()`` first (the ``nerfstudio`` stand-in package).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Level = Tuple[Tensor, Tensor, Tensor]          # (weights, starts, ends) of one proposal level, [R, S_i, 1] each


def camera_ray_bundle(height: int, width: int, device="cpu"):
    """An image-shaped ``RayBundle`` whose ``camera_indices`` carry the row-major ray index, which is how the
    stand-in sampler finds the rows of its tensors after ``get_row_major_sliced_ray_bundle``."""
    from nerfstudio.cameras.rays import RayBundle

    idx = torch.arange(height * width, device=device).view(height, width, 1)
    return RayBundle(origins=torch.zeros(height, width, 3, device=device),
                     directions=torch.zeros(height, width, 3, device=device), camera_indices=idx)


def flat_ray_bundle(num_rays: int, device="cpu"):
    from nerfstudio.cameras.rays import RayBundle

    return RayBundle(origins=torch.zeros(num_rays, 3, device=device), directions=torch.zeros(num_rays, 3, device=device),
                     camera_indices=torch.arange(num_rays, device=device).view(num_rays, 1))


class TensorSampler:
    """``proposal_sampler`` stand-in: returns the stored final-level samples and proposal levels of the rays in
    the bundle (``(ray_samples, weights_list, ray_samples_list)`` like nerfstudio's ``ProposalNetworkSampler``)."""

    def __init__(self, deltas: Tensor, starts: Tensor, ends: Tensor, proposal_levels: Sequence[Level] = ()):
        self.deltas, self.starts, self.ends = deltas, starts, ends
        self.levels = list(proposal_levels)

    def __call__(self, ray_bundle, density_fns=None):
        from nerfstudio.cameras.rays import Frustums, RaySamples

        idx = ray_bundle.camera_indices.reshape(-1)
        rs = RaySamples(frustums=Frustums(starts=self.starts[idx], ends=self.ends[idx]), deltas=self.deltas[idx])
        rs.ub_ray_index = idx
        weights_list, samples_list = [], []
        for w, s, e in self.levels:
            weights_list.append(w[idx])
            samples_list.append(RaySamples(frustums=Frustums(starts=s[idx], ends=e[idx])))
        return rs, weights_list, samples_list


class TensorField:
    """``field`` stand-in: ``forward`` / ``forward_unc`` return the stored per-sample tensors of the rays in
    ``ray_samples``.  ``extras`` are passed through under their string keys (``rgb_var``, ``density_var``)."""

    def __init__(self, density: Tensor, rgb: Tensor, **extras: Optional[Tensor]):
        self.density, self.rgb, self.extras = density, rgb, extras
        self.calls: List[dict] = []

    def _out(self, ray_samples):
        from nerfstudio.field_components.field_heads import FieldHeadNames

        idx = ray_samples.ub_ray_index
        out = {FieldHeadNames.RGB: self.rgb[idx], FieldHeadNames.DENSITY: self.density[idx]}
        for k, v in self.extras.items():
            out[k] = None if v is None else v[idx]
        return out

    def forward(self, ray_samples, compute_normals=False, **kwargs):
        self.calls.append(dict(kwargs, compute_normals=compute_normals))
        return self._out(ray_samples)

    def forward_unc(self, ray_samples, compute_normals=False, **kwargs):
        self.calls.append(dict(kwargs, compute_normals=compute_normals))
        out = self._out(ray_samples)
        if kwargs.get("use_deterministic_density"):
            out["density_var"] = None          # laplace_field.py:499-504
        return out


class ReplayModel:
    """A member model of an ensemble / the nerfacto parent of the MC-dropout model: ``get_outputs_for_camera``
    (``_ray_bundle``) returns stored per-view output dicts, one per call."""

    def __init__(self, renders: Sequence[Dict[str, Tensor]]):
        self.renders = list(renders)
        self.calls = 0

    def get_outputs_for_camera(self, camera, obb_box=None):
        return self.get_outputs_for_camera_ray_bundle(camera)

    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle):
        out = self.renders[self.calls % len(self.renders)]
        self.calls += 1
        return dict(out)


def attach_producers(model, inp: Dict[str, Tensor], proposal_levels: Sequence[Level] = (), field=None):
    """Give a stand-in-constructed nerfacto-family model its sampler and field."""
    model.proposal_sampler = TensorSampler(inp["deltas"], inp["starts"], inp["ends"], proposal_levels)
    model.field = field if field is not None else TensorField(inp["density"], inp["rgb"], rgb_var=inp.get("beta"))
    model.density_fns = []
    return model


def proposal_levels(num_rays: int, seed: int, sizes=(16, 24), device="cpu") -> List[Level]:
    g = torch.Generator(device=device).manual_seed(1000 + seed)
    levels = []
    for s in sizes:
        w = torch.rand(num_rays, s, 1, generator=g, device=device)
        w = w / w.sum(dim=1, keepdim=True) * torch.rand(num_rays, 1, 1, generator=g, device=device)
        d = torch.rand(num_rays, s, 1, generator=g, device=device) * 0.2 + 1e-3
        starts = torch.cumsum(d, dim=1) - d + 0.05
        levels.append((w, starts, starts + d))
    return levels
