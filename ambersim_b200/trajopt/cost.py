"""StaticGoalQuadraticCost (reference ambersim/trajopt/cost.py:13-178) on torch tensors.

`cost` is the formula the CUDA engine fuses into the rollout (cost.py:62-85):
    0.5 * [ sum_{t<N} (x_t-xg)' Q (x_t-xg) + (x_N-xg)' Qf (x_N-xg) + sum_{t<N} u_t' R u_t ]
`device_cost()` hands the matrices to the engine (abr_cost_create). `grad`/`hess` are the closed
forms (host-side torch; not on the sampling hot path).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from ambersim_b200 import _abi, _lib
from ambersim_b200.trajopt.base import CostFunction, CostFunctionParams


class _CostHandle:
    def __init__(self, Q, Qf, R, xg, device: int):
        host, keep = _abi.pack_cost(Q, Qf, R, xg)
        ptr = C.c_void_p()
        _lib.check(_lib.lib().abr_cost_create(C.byref(host), device, C.byref(ptr)))
        self.ptr = ptr
        del keep

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().abr_cost_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


class StaticGoalQuadraticCost(CostFunction):
    """Quadratic distance-to-a-fixed-goal cost with static dense weights."""

    def __init__(self, Q, Qf, R, xg) -> None:
        """Q, Qf (nx, nx) state / terminal weights, R (nu, nu) control weight, xg (nx,) goal state."""
        self.Q, self.Qf, self.R, self.xg = (torch.as_tensor(a, dtype=torch.float32) for a in (Q, Qf, R, xg))
        self._handles = {}

    def device_cost(self, device: int = 0) -> _CostHandle:
        h = self._handles.get(device)
        if h is None:
            h = self._handles[device] = _CostHandle(_np(self.Q), _np(self.Qf), _np(self.R), _np(self.xg), device)
        return h

    def _w(self, like: torch.Tensor):
        return (a.to(device=like.device, dtype=like.dtype) for a in (self.Q, self.Qf, self.R, self.xg))

    @staticmethod
    def batch_quadform(bs: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
        """b' A b over the last axis of bs."""
        return ((bs @ A) * bs).sum(-1)

    @staticmethod
    def batch_matmul(bs: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
        """b' A over the last axis of bs."""
        return bs @ A

    def cost(self, xs: torch.Tensor, us: torch.Tensor, params: CostFunctionParams = None) -> Tuple[torch.Tensor, CostFunctionParams]:
        Q, Qf, R, xg = self._w(xs)
        err = xs - xg
        running = self.batch_quadform(err[..., :-1, :], Q).sum(-1)
        terminal = self.batch_quadform(err[..., -1, :], Qf)
        effort = self.batch_quadform(us, R).sum(-1)
        return 0.5 * (running + terminal + effort), params

    def grad(self, xs, us, params=None):
        Q, Qf, R, xg = self._w(xs)
        err = xs - xg
        # d/dx of 0.5 x'Ax is 0.5 (A + A') x; the reference writes x'A, identical for symmetric weights
        g_xs = torch.cat((self.batch_matmul(err[:-1], Q), (Qf @ err[-1])[None, :]), dim=-2)
        g_us = self.batch_matmul(us, R)
        return g_xs, g_us, params, params

    def hess(self, xs, us, params=None):
        Q, Qf, R, _ = self._w(xs)
        N, nu = us.shape
        nx = Q.shape[0]
        eye_x = torch.eye(N + 1, dtype=xs.dtype, device=xs.device)
        blocks = Q[None].repeat(N + 1, 1, 1)
        blocks[-1] = Qf
        # H[t, i, s, j] = delta(t, s) * W_t[i, j]
        h_xx = torch.einsum("ts,tij->tisj", eye_x, blocks)
        h_uu = torch.einsum("ts,ij->tisj", torch.eye(N, dtype=xs.dtype, device=xs.device), R)
        h_xu = torch.zeros((N + 1, nx, N, nu), dtype=xs.dtype, device=xs.device)
        empty = CostFunctionParams()
        return h_xx, h_xu, empty, h_uu, empty, empty, params
