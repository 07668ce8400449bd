"""``NerfactoModel`` / ``NerfactoModelConfig`` with the members the reference's subclasses rely on
(nerfstudio 1.1.0 ``models/nerfacto.py``): config defaults, the four renderers, and ``get_outputs`` of plain
nerfacto (the members of seed ensembles, branch B of ensemble_pipeline.py:186-189).  ``populate_modules`` does
not build fields or samplers -- tests attach stand-ins for those producers."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Literal, Optional, Type

import torch

from nerfstudio.field_components.field_heads import FieldHeadNames
from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer
from nerfstudio.models.base_model import Model, ModelConfig


@dataclass
class NerfactoModelConfig(ModelConfig):
    _target: Type = field(default_factory=lambda: NerfactoModel)
    near_plane: float = 0.05
    far_plane: float = 1000.0
    background_color: Any = "last_sample"
    hidden_dim: int = 64
    hidden_dim_color: int = 64
    hidden_dim_transient: int = 64
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    num_proposal_samples_per_ray: tuple = (256, 96)
    num_nerf_samples_per_ray: int = 48
    num_proposal_iterations: int = 2
    use_proposal_weight_anneal: bool = True
    use_appearance_embedding: bool = True
    use_average_appearance_embedding: bool = True
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    orientation_loss_mult: float = 0.0001
    pred_normal_loss_mult: float = 0.001
    use_gradient_scaling: bool = False
    predict_normals: bool = False
    disable_scene_contraction: bool = False
    implementation: Literal["tcnn", "torch"] = "tcnn"
    appearance_embed_dim: int = 32
    average_init_density: float = 1.0
    camera_optimizer: Optional[Any] = None
    eval_num_rays_per_chunk: int = 1 << 15


class NerfactoModel(Model):
    config: NerfactoModelConfig

    def populate_modules(self):
        self.renderer_rgb = RGBRenderer(background_color=self.config.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="median")
        self.renderer_expected_depth = DepthRenderer(method="expected")
        self.density_fns = []
        self.proposal_sampler = None
        self.field = None

    def get_outputs(self, ray_bundle):
        ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
        field_outputs = self.field.forward(ray_samples, compute_normals=self.config.predict_normals)
        weights = ray_samples.get_weights(field_outputs[FieldHeadNames.DENSITY])
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        rgb = self.renderer_rgb(rgb=field_outputs[FieldHeadNames.RGB], weights=weights)
        with torch.no_grad():
            depth = self.renderer_depth(weights=weights, ray_samples=ray_samples)
        expected_depth = self.renderer_expected_depth(weights=weights, ray_samples=ray_samples)
        accumulation = self.renderer_accumulation(weights=weights)
        outputs = {"rgb": rgb, "accumulation": accumulation, "depth": depth, "expected_depth": expected_depth}
        if self.training:
            outputs["weights_list"] = weights_list
            outputs["ray_samples_list"] = ray_samples_list
        for i in range(self.config.num_proposal_iterations):
            outputs[f"prop_depth_{i}"] = self.renderer_depth(weights=weights_list[i], ray_samples=ray_samples_list[i])
        return outputs


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
