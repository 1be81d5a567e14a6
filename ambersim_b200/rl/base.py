"""MJX-style environment base on the CUDA engine (reference ambersim/rl/base.py:14-153).

`MjxEnv.pipeline_init(qpos, qvel)` and `MjxEnv.pipeline_step(data, ctrl)` keep the reference
signatures; leading batch dimensions replace brax's VmapWrapper. `VectorEnvStepper` is the
in-place, auto-resetting hot loop used for env-step throughput (brax AutoResetWrapper semantics:
`where(done, first_state, state)` folded into the next step's prologue).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional

import numpy as np
import torch

from ambersim_b200 import _lib, mjx
from ambersim_b200.utils.mjcf import MjModel


@dataclasses.dataclass
class State:
    """Environment state (reference rl/base.py:14-32)."""

    pipeline_state: mjx.Data
    obs: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
    metrics: Dict[str, torch.Tensor] = dataclasses.field(default_factory=dict)
    info: Dict[str, Any] = dataclasses.field(default_factory=dict)

    def replace(self, **kw) -> "State":
        return dataclasses.replace(self, **kw)


mjx._register_dataclass_pytree(State)


class MjxEnv(ABC):
    """API for an engine-backed system for training and inference.

    `data_fields`: names of derived `mjx.Data` fields (`mjx.DERIVED_FIELDS`: xpos, xquat, cvel, contact_dist, efc_force, ...)
    that `compute_obs` / `compute_reward` read (the pattern ambersim/rl/base.py:98-125 is designed for). They are filled for
    the whole batch by the launch that does the physics; leave it empty when obs / reward only need qpos / qvel."""

    data_fields: tuple = ()

    def __init__(self, mj_model: MjModel, physics_steps_per_control_step: int = 1, device=None) -> None:
        assert physics_steps_per_control_step >= 1
        self.model = mj_model
        self.sys = mjx.device_put(mj_model)
        self._physics_steps_per_control_step = physics_steps_per_control_step
        self._device = device

    @property
    def dt(self) -> float:
        """Time per env step."""
        return self.sys.opt.timestep * self._physics_steps_per_control_step

    @property
    def observation_size(self) -> int:
        return self.reset(0).obs.shape[-1]

    @property
    def action_size(self) -> int:
        return self.sys.nu

    @property
    def backend(self) -> str:
        return "mjx"

    def pipeline_init(self, qpos: torch.Tensor, qvel: torch.Tensor) -> mjx.Data:
        """Initialises the physics state: ctrl = 0, then mjx.forward (reference rl/base.py:81-86)."""
        dev = mjx._dev(self._device if not (isinstance(qpos, torch.Tensor) and qpos.is_cuda) else qpos.device)
        qpos = torch.as_tensor(qpos, dtype=torch.float32, device=dev)
        qvel = torch.as_tensor(qvel, dtype=torch.float32, device=dev)
        batch = tuple(qpos.shape[:-1])
        data = mjx.make_data(self.sys, device=dev)
        data = data.replace(qpos=qpos, qvel=qvel, ctrl=torch.zeros(*batch, self.sys.nu, device=dev),
                            qacc_warmstart=torch.zeros(*batch, self.sys.nv, device=dev),
                            time=torch.zeros(batch, device=dev))
        return mjx.forward(self.sys, data, fields=self.data_fields)

    def pipeline_step(self, data: mjx.Data, ctrl: torch.Tensor) -> mjx.Data:
        """Holds ctrl and takes physics_steps_per_control_step physics steps (reference rl/base.py:88-96)."""
        ctrl = torch.as_tensor(ctrl, dtype=torch.float32, device=data.qpos.device)
        return mjx.step(self.sys, data.replace(ctrl=ctrl), nsubsteps=self._physics_steps_per_control_step, fields=self.data_fields)

    def compute_reward(self, data: mjx.Data, info: Dict[str, Any]) -> torch.Tensor:
        raise NotImplementedError

    def compute_obs(self, data: mjx.Data, info: Dict[str, Any]) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def reset(self, rng) -> State:
        """Resets the environment."""

    @abstractmethod
    def step(self, state: State, action: torch.Tensor) -> State:
        """Takes a step in the environment."""


class VectorEnvStepper:
    """E environments stepped in place with auto-reset: the physics share of a brax training step.

    Buffers (qpos, qvel, qacc_warmstart, time) live on the device; `step(ctrl, done)` first blends in
    the cached first state where `done` is set (AutoResetWrapper), then runs nsubsteps x mjx.step.
    """

    def __init__(self, model: mjx.Model, qpos0: torch.Tensor, qvel0: torch.Tensor, nsubsteps: int = 1):
        self.model, self.nsubsteps = model, int(nsubsteps)
        dev = qpos0.device
        self.E = qpos0.shape[0]
        d0 = mjx.forward(model, mjx.Data(qpos=qpos0, qvel=qvel0, ctrl=torch.zeros(self.E, model.nu, device=dev),
                                         qacc=torch.zeros(self.E, model.nv, device=dev),
                                         qacc_warmstart=torch.zeros(self.E, model.nv, device=dev),
                                         time=torch.zeros(self.E, device=dev)))
        self.first_qpos, self.first_qvel, self.first_warm = d0.qpos.contiguous(), d0.qvel.contiguous(), d0.qacc_warmstart.contiguous()
        self.qpos, self.qvel, self.warm = self.first_qpos.clone(), self.first_qvel.clone(), self.first_warm.clone()
        self.time = torch.zeros(self.E, device=dev)
        self._h = model.handle(dev.index or 0)

    def step(self, ctrl: torch.Tensor, done: Optional[torch.Tensor] = None) -> None:
        dev = self.qpos.device
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        mask = None if done is None else done.to(torch.uint8).contiguous()
        ctrl = ctrl.to(torch.float32).contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().abr_env_step_dev(self._h.ptr, p(self.qpos), p(self.qvel), p(self.warm), p(self.time), p(ctrl),
                                               self.E, self.nsubsteps, p(mask), p(self.first_qpos), p(self.first_qvel),
                                               p(self.first_warm), stream))
