"""``Model`` / ``ModelConfig``: the chunked per-camera evaluation loop every nerfstudio model inherits
(nerfstudio 1.1.0 ``models/base_model.py``; restated in the reference at laplace_model.py:269-297)."""
from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass, field
from typing import Dict, Optional, Type

import torch
from torch import nn


@dataclass
class ModelConfig:
    _target: Type = field(default_factory=lambda: Model)
    enable_collider: bool = True
    collider_params: Optional[Dict[str, float]] = None
    loss_coefficients: Optional[Dict[str, float]] = None
    eval_num_rays_per_chunk: int = 4096
    prompt: Optional[str] = None

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class Model(nn.Module):
    config: ModelConfig

    def __init__(self, config=None, scene_box=None, num_train_data: int = 0, **kwargs):
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.device_indicator_param = nn.Parameter(torch.empty(0))
        self.populate_modules()

    def populate_modules(self):
        pass

    @property
    def device(self):
        return self.device_indicator_param.device

    def forward(self, ray_bundle):
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)

    @torch.no_grad()
    def get_outputs_for_camera(self, camera, obb_box=None):
        return self.get_outputs_for_camera_ray_bundle(
            camera.generate_rays(camera_indices=0, keep_shape=True, obb_box=obb_box))

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle):
        num_rays_per_chunk = self.config.eval_num_rays_per_chunk
        image_height, image_width = camera_ray_bundle.origins.shape[:2]
        num_rays = len(camera_ray_bundle)
        outputs_lists = defaultdict(list)
        for i in range(0, num_rays, num_rays_per_chunk):
            ray_bundle = camera_ray_bundle.get_row_major_sliced_ray_bundle(i, i + num_rays_per_chunk)
            outputs = self.forward(ray_bundle=ray_bundle)
            for output_name, output in outputs.items():
                if not isinstance(output, torch.Tensor):
                    continue
                outputs_lists[output_name].append(output)
        return {name: torch.cat(parts).view(image_height, image_width, -1) for name, parts in outputs_lists.items()}


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
