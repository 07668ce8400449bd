"""``SplatfactoModel`` / ``SplatfactoModelConfig`` with the members ``ActiveSplatfactoModel.get_outputs`` reads
(activesplatfacto_model.py:142-367; nerfstudio 1.1.0 ``models/splatfacto.py``): the Gaussian parameter dict and
its accessor properties, background handling, the identity camera optimiser, ``get_empty_outputs`` and the
ground-truth compositing helpers the scorer calls (eval_uncertainty.py:320-322).  ``populate_modules`` takes the
Gaussians from ``kwargs['seed_gaussians']`` (a dict of tensors) instead of an SfM point cloud."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Type

import torch

from nerfstudio.models.base_model import Model, ModelConfig


@dataclass
class SplatfactoModelConfig(ModelConfig):
    _target: Type = field(default_factory=lambda: SplatfactoModel)
    background_color: str = "random"
    sh_degree: int = 3
    sh_degree_interval: int = 1000
    rasterize_mode: str = "classic"
    output_depth_during_training: bool = False
    num_downscales: int = 2
    resolution_schedule: int = 3000
    use_scale_regularization: bool = False
    max_gauss_ratio: float = 10.0
    ssim_lambda: float = 0.2
    camera_optimizer: Any = None


class _IdentityCameraOptimizer(torch.nn.Module):
    def apply_to_camera(self, camera):
        return camera.camera_to_worlds

    def get_loss_dict(self, loss_dict):
        return None


class SplatfactoModel(Model):
    config: SplatfactoModelConfig

    def populate_modules(self):
        seed = self.kwargs.get("seed_gaussians")
        assert seed is not None, "stub SplatfactoModel: pass seed_gaussians={means, scales, quats, features_dc, features_rest, opacities}"
        self.gauss_params = torch.nn.ParameterDict({k: torch.nn.Parameter(v.clone()) for k, v in seed.items()})
        self.camera_optimizer = _IdentityCameraOptimizer()
        self.crop_box = None
        self.step = 30000
        self.background_color = torch.tensor([0.1490, 0.1647, 0.2157]) if self.config.background_color == "random" \
            else torch.tensor({"white": [1.0, 1.0, 1.0], "black": [0.0, 0.0, 0.0]}[self.config.background_color])

    means = property(lambda self: self.gauss_params["means"])
    scales = property(lambda self: self.gauss_params["scales"])
    quats = property(lambda self: self.gauss_params["quats"])
    features_dc = property(lambda self: self.gauss_params["features_dc"])
    features_rest = property(lambda self: self.gauss_params["features_rest"])
    opacities = property(lambda self: self.gauss_params["opacities"])
    num_points = property(lambda self: self.gauss_params["means"].shape[0])

    def _get_downscale_factor(self):
        return 1

    def _downscale_if_required(self, image):
        return image

    def get_empty_outputs(self, width: int, height: int, background: torch.Tensor) -> Dict[str, Any]:
        rgb = background.repeat(height, width, 1)
        depth = background.new_ones(*rgb.shape[:2], 1) * 10
        accumulation = background.new_zeros(*rgb.shape[:2], 1)
        return {"rgb": rgb, "depth": depth, "accumulation": accumulation, "background": background}

    @staticmethod
    def get_gt_img(image: torch.Tensor):
        if image.dtype == torch.uint8:
            image = image.float() / 255.0
        return image

    def composite_with_background(self, image, background):
        if image.shape[2] == 4:
            alpha = image[..., -1].unsqueeze(-1).repeat((1, 1, 3))
            return alpha * image[..., :3] + (1 - alpha) * background
        return image


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
