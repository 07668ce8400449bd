"""``Cameras``: a real class so that ``isinstance(camera, Cameras)`` (activesplatfacto_model.py:152) works,
with the pinhole members ``get_outputs`` reads (``:197-203, 231-232``)."""
from __future__ import annotations

import torch


class Cameras:
    def __init__(self, camera_to_worlds, fx, fy, cx, cy, width, height):
        t = lambda v: torch.as_tensor(v).reshape(1, 1)
        self.camera_to_worlds = torch.as_tensor(camera_to_worlds, dtype=torch.float32).reshape(1, 3, 4)
        self.fx, self.fy, self.cx, self.cy = (t(float(v)).float() for v in (fx, fy, cx, cy))
        self.width, self.height = t(int(width)).long(), t(int(height)).long()
        self.ray_bundle = None   # tests may attach the camera ray bundle ``generate_rays`` should return

    @property
    def shape(self):
        return self.camera_to_worlds.shape[:-2]

    @property
    def device(self):
        return self.camera_to_worlds.device

    def rescale_output_resolution(self, scaling_factor):
        if scaling_factor == 1 or scaling_factor == 1.0:
            return
        self.fx, self.fy = self.fx * scaling_factor, self.fy * scaling_factor
        self.cx, self.cy = self.cx * scaling_factor, self.cy * scaling_factor
        self.width = (self.width * scaling_factor).to(torch.int64)
        self.height = (self.height * scaling_factor).to(torch.int64)

    def generate_rays(self, camera_indices=0, keep_shape=True, obb_box=None, **_):
        assert self.ray_bundle is not None, "stub Cameras: attach .ray_bundle first"
        return self.ray_bundle


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
