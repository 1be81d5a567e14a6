"""API shapes of ambersim/trajopt/base.py (reference :12-172), on torch tensors.

Only the signatures matter to the hot path; the default `grad`/`hess` use torch autograd where the
reference uses jax.grad / jax.hessian (reference base.py:119-172).
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

import torch


@dataclasses.dataclass
class TrajectoryOptimizerParams:
    """Inputs of a trajectory optimiser (initial iterates, seeds, ...). Reference base.py:12-36."""


@dataclasses.dataclass
class TrajectoryOptimizer:
    """A trajectory optimiser for mechanical systems. Reference base.py:39-78."""

    def optimize(self, params: TrajectoryOptimizerParams) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (xs_star (N+1, nq+nv), us_star (N, nu))."""
        raise NotImplementedError


@dataclasses.dataclass
class CostFunctionParams:
    """Generic cost-function parameters. Reference base.py:86-88."""


class CostFunction:
    """cost / grad / hess of a trajectory. Reference base.py:91-172."""

    def cost(self, xs: torch.Tensor, us: torch.Tensor, params: CostFunctionParams) -> Tuple[torch.Tensor, CostFunctionParams]:
        """xs (N+1, nq+nv), us (N, nu) -> (scalar cost, new params)."""
        raise NotImplementedError

    def grad(self, xs, us, params):
        """Default: autograd of `cost` wrt (xs, us). Returns (g_xs, g_us, g_params, new_params)."""
        xs_ = xs.detach().clone().requires_grad_(True)
        us_ = us.detach().clone().requires_grad_(True)
        val, _ = self.cost(xs_, us_, params)
        g_xs, g_us = torch.autograd.grad(val, (xs_, us_))
        return g_xs, g_us, params, params

    def hess(self, xs, us, params):
        """Default: autograd Hessian blocks (H_xsxs, H_xsus, H_xsparams, H_usus, H_usparams, H_paramsall, new_params)."""
        fn = lambda a, b: self.cost(a, b, params)[0]
        (h_xx, h_xu), (_, h_uu) = torch.autograd.functional.hessian(fn, (xs.detach(), us.detach()))
        return h_xx, h_xu, params, h_uu, params, params, params
