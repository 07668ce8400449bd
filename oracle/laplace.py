"""Oracle: last-layer diagonal-Laplace Monte-Carlo moments (torch, CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates
``nerfuncertainty/models/laplace/laplace_field.py:528-568`` (``sample_laplace``) and the rgb-head
post-processing at ``:478-484``.  The only change: the standard-normal draws are an argument instead of
``torch.randn`` on the global generator, so that the CUDA path can be fed identical samples.
PINNED: ``tests/test_oracle_pinned.py::test_live_sample_laplace`` runs the reference's ``sample_laplace`` unmodified with the
global generator seeded, so that its ``torch.randn`` returns the draws handed to the oracle, and demands bit equality
(also ``tests/golden/ref_laplace.npz``).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

Tensor = torch.Tensor


def posterior_samples(mu_q: Tensor, diag_ggn: Tensor, eps_draws: Tensor, prior_prec: float = 1.0,
                      eps: float = 1e-9) -> Tensor:
    """``laplace_field.py:541-547``: ``mu + randn * 1/sqrt(ggn + prior_prec + eps)`` -> ``[n_samples, n_params]``."""
    precision = diag_ggn + prior_prec
    post_std = 1 / torch.sqrt(precision + eps)
    return mu_q.view(1, -1) + eps_draws * post_std.view(1, -1)


def sample_laplace(x: Tensor, sampled_params: Tensor, out_dim: int, activation: Callable[[Tensor], Tensor]
                   ) -> Tuple[Tensor, Tensor, Tensor]:
    """``laplace_field.py:549-565``: loop over sampled parameter vectors (weight ``[out, in]`` row-major
    then bias, the order of ``parameters_to_vector``), accumulate E[y] and E[y^2].  Returns
    ``(mean, mean2, sigma2)``."""
    hidden = x.shape[-1]
    n = sampled_params.shape[0]
    mu = 0.0
    mu2 = 0.0
    for theta in sampled_params:
        w = theta[: out_dim * hidden].view(out_dim, hidden)
        b = theta[out_dim * hidden:]
        pred = activation(torch.nn.functional.linear(x, w, b))
        mu = mu + pred
        mu2 = mu2 + pred ** 2
    mu = mu / n
    mu2 = mu2 / n
    return mu, mu2, mu2 - mu ** 2


def rgb_variance_from_sigma2(sigma2_rgb: Tensor) -> Tensor:
    """``laplace_field.py:478-482``: clamp negatives to 0 and average over the channel axis."""
    return torch.relu(sigma2_rgb).mean(dim=-1)[..., None]
