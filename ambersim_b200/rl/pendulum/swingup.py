"""Torque-limited pendulum swing-up (reference ambersim/rl/pendulum/swingup.py:14-122).

States x = (theta, dtheta); observations y = (cos theta, sin theta, dtheta); action = motor torque.
obs / reward are elementwise host-framework math (torch); the physics is the CUDA engine. A leading
batch dimension on `rng`-derived states replaces brax's VmapWrapper.
"""
from __future__ import annotations

import dataclasses
import math
from pathlib import Path
from typing import Any, Dict, Optional, Union

import torch

from ambersim_b200 import mjx
from ambersim_b200.rl.base import MjxEnv, State
from ambersim_b200.utils.io_utils import load_mj_model_from_file


@dataclasses.dataclass
class PendulumSwingupConfig:
    """Config of the swing-up task (reference swingup.py:14-36)."""

    model_path: Union[Path, str] = "models/pendulum/scene.xml"
    physics_steps_per_control_step: int = 1
    stdev_obs: float = 0.0
    theta_cost_weight: float = 1.0
    theta_dot_cost_weight: float = 0.1
    control_cost_weight: float = 0.001
    qpos_hi: float = math.pi
    qpos_lo: float = -math.pi
    qvel_hi: float = 2.0
    qvel_lo: float = -2.0


class PendulumSwingupEnv(MjxEnv):
    """Swing a pendulum from hanging to upright under a torque limit."""

    def __init__(self, config: Optional[PendulumSwingupConfig] = None, num_envs: Optional[int] = None, device=None) -> None:
        self.config = config or PendulumSwingupConfig()
        self.num_envs = num_envs
        super().__init__(load_mj_model_from_file(self.config.model_path), self.config.physics_steps_per_control_step, device)

    def compute_obs(self, data: mjx.Data, info: Dict[str, Any]) -> torch.Tensor:
        theta = data.qpos[..., 0]
        return torch.stack((torch.cos(theta), torch.sin(theta), data.qvel[..., 0]), dim=-1)

    def compute_reward(self, data: mjx.Data, info: Dict[str, Any]) -> torch.Tensor:
        """Maximal at theta = pi (upright), zero velocity, zero torque."""
        theta, theta_dot, tau = data.qpos[..., 0], data.qvel[..., 0], data.ctrl[..., 0]
        err = theta - math.pi
        err = torch.atan2(torch.sin(err), torch.cos(err))
        cfg = self.config
        return -cfg.theta_cost_weight * err**2 - cfg.theta_dot_cost_weight * theta_dot**2 - cfg.control_cost_weight * tau**2

    def _generator(self, rng, dev):
        if isinstance(rng, torch.Generator):
            return rng
        g = torch.Generator(device=dev)
        g.manual_seed(int(rng))
        return g

    def reset(self, rng) -> State:
        dev = mjx._dev(self._device)
        g = self._generator(rng, dev)
        batch = () if self.num_envs is None else (self.num_envs,)
        cfg = self.config
        u = lambda n, lo, hi: lo + (hi - lo) * torch.rand(*batch, n, generator=g, device=dev)
        data = self.pipeline_init(u(self.sys.nq, cfg.qpos_lo, cfg.qpos_hi), u(self.sys.nv, cfg.qvel_lo, cfg.qvel_hi))
        obs = self.compute_obs(data, {})
        zero = torch.zeros(batch, device=dev)
        return State(data, obs, zero, zero.clone(), {"reward": zero.clone()}, {"rng": g, "step": 0})

    def step(self, state: State, action: torch.Tensor) -> State:
        data = self.pipeline_step(state.pipeline_state, action)
        obs = self.compute_obs(data, state.info)
        if self.config.stdev_obs:
            obs = obs + torch.randn(obs.shape, generator=state.info["rng"], device=obs.device) * self.config.stdev_obs
        reward = self.compute_reward(data, state.info)
        info = dict(state.info)
        info["step"] = info["step"] + 1
        metrics = dict(state.metrics)
        metrics["reward"] = reward
        return state.replace(pipeline_state=data, obs=obs, reward=reward, done=torch.zeros_like(reward), metrics=metrics, info=info)
