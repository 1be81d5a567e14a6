"""Training wrappers and the fused task environment.

The reference trains its `MjxEnv`s through brax's training wrappers (`ambersim/rl/base.py:9`,
`examples/rl/pendulum/ex_swingup.py`): `VmapWrapper` (batch of envs), `EpisodeWrapper` (episode length,
truncation flag) and `AutoResetWrapper` (`where(done, first_state, state)`). Here a leading batch dimension
replaces `VmapWrapper`, and the other two are provided twice:

* `EpisodeWrapper` / `AutoResetWrapper`: torch restatements that wrap ANY `MjxEnv` (obs / reward stay the
  user's host-framework code, one physics launch per step plus elementwise torch ops);
* `FusedQuadraticTaskEnv`: the same semantics for `QuadraticTaskEnv` in ONE launch per env step
  (`abr_env_task_step_dev`: physics, obs, reward, done, episode counter and the auto-reset blend fused into
  the step kernel's epilogue). `tests/test_gpu_parity.py` checks the two against each other.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Optional

import numpy as np
import torch

from ambersim_b200 import _lib, mjx
from ambersim_b200.rl.base import MjxEnv, State
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.utils.mjcf import MjModel


class Wrapper:
    """Forwards everything to the wrapped env (brax.envs.base.Wrapper)."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def reset(self, rng) -> State:
        return self.env.reset(rng)

    def step(self, state: State, action: torch.Tensor) -> State:
        return self.env.step(state, action)


class VmapWrapper(Wrapper):
    """brax.envs.wrappers.training.VmapWrapper for API compatibility: the engine's envs are batched natively (a leading dimension on
    every `State` / `mjx.Data` field replaces `jax.vmap`), so this only checks / records the batch size."""

    def __init__(self, env, batch_size: Optional[int] = None):
        super().__init__(env)
        self.batch_size = batch_size

    def reset(self, rng) -> State:
        state = self.env.reset(rng)
        if self.batch_size is not None and tuple(state.done.shape[:1]) != (self.batch_size,):
            raise ValueError(f"VmapWrapper(batch_size={self.batch_size}) around an env that resets {tuple(state.done.shape)} worlds")
        return state


class EpisodeWrapper(Wrapper):
    """Episode length and action repeat: info['steps'] counts env steps, `done` is raised at
    `episode_length` and info['truncation'] marks episodes ended by the length alone."""

    def __init__(self, env, episode_length: int, action_repeat: int = 1):
        super().__init__(env)
        self.episode_length, self.action_repeat = int(episode_length), int(action_repeat)

    def reset(self, rng) -> State:
        state = self.env.reset(rng)
        info = dict(state.info)
        info["steps"] = torch.zeros_like(state.done, dtype=torch.int32)
        info["truncation"] = torch.zeros_like(state.done)
        return state.replace(info=info)

    def step(self, state: State, action: torch.Tensor) -> State:
        reward = torch.zeros_like(state.reward)
        for _ in range(self.action_repeat):
            state = self.env.step(state, action)
            reward = reward + state.reward
        steps = state.info["steps"] + self.action_repeat
        over = steps >= self.episode_length
        info = dict(state.info)
        info["truncation"] = torch.where(over, 1.0 - state.done, torch.zeros_like(state.done))
        info["steps"] = steps
        return state.replace(reward=reward, done=torch.where(over, torch.ones_like(state.done), state.done), info=info)


class AutoResetWrapper(Wrapper):
    """Resets finished envs to the state their episode started from: where(done, first_state, state) on the
    physics state and the observation; reward and done keep the finished step's values."""

    def reset(self, rng) -> State:
        state = self.env.reset(rng)
        info = dict(state.info)
        info["first_pipeline_state"] = state.pipeline_state
        info["first_obs"] = state.obs
        return state.replace(info=info)

    def step(self, state: State, action: torch.Tensor) -> State:
        info = dict(state.info)
        if "steps" in info:
            info["steps"] = torch.where(state.done.bool(), torch.zeros_like(info["steps"]), info["steps"])
        state = self.env.step(state.replace(done=torch.zeros_like(state.done), info=info), action)
        done = state.done.bool()

        def where_done(first, cur):
            d = done.reshape(done.shape + (1,) * (cur.dim() - done.dim()))
            return torch.where(d, first, cur)

        first, cur = state.info["first_pipeline_state"], state.pipeline_state
        data = cur.replace(**{f.name: where_done(getattr(first, f.name), getattr(cur, f.name)) for f in dataclasses.fields(cur)
                              if getattr(cur, f.name) is not None and getattr(first, f.name) is not None})  # derived fields are optional
        return state.replace(pipeline_state=data, obs=where_done(state.info["first_obs"], state.obs))


class QuadraticTaskEnv(MjxEnv):
    """Track a goal state with a floating-base robot: obs = (qpos, qvel),
    reward = -(0.5 (x - xg)' Q (x - xg) + 0.5 u' R u) on the stepped state, done = base height < z_min.
    `reset` starts from `qpos0` (+ uniform joint jitter) at rest; a leading batch of `num_envs`."""

    def __init__(self, mj_model: MjModel, reward: StaticGoalQuadraticCost, qpos0, num_envs: int, z_min: float = -math.inf,
                 jitter: float = 0.0, physics_steps_per_control_step: int = 1, device=None) -> None:
        super().__init__(mj_model, physics_steps_per_control_step, device)
        self.reward_fn, self.num_envs, self.z_min, self.jitter = reward, int(num_envs), float(z_min), float(jitter)
        self.qpos0 = np.asarray(qpos0, dtype=np.float32)

    def compute_obs(self, data: mjx.Data, info=None) -> torch.Tensor:
        return torch.cat((data.qpos, data.qvel), dim=-1)

    def compute_reward(self, data: mjx.Data, info=None) -> torch.Tensor:
        Q, _, R, xg = self.reward_fn._w(data.qpos)
        err = torch.cat((data.qpos, data.qvel), dim=-1) - xg
        return -0.5 * (self.reward_fn.batch_quadform(err, Q) + self.reward_fn.batch_quadform(data.ctrl, R))

    def reset(self, rng) -> State:
        dev = mjx._dev(self._device)
        g = rng if isinstance(rng, torch.Generator) else torch.Generator(device=dev).manual_seed(int(rng))
        qpos = torch.tensor(self.qpos0, device=dev).repeat(self.num_envs, 1)
        nj = self.sys.nq - 7
        qpos[:, 7:] += (torch.rand((self.num_envs, nj), generator=g, device=dev) - 0.5) * 2 * self.jitter
        data = self.pipeline_init(qpos, torch.zeros(self.num_envs, self.sys.nv, device=dev))
        zero = torch.zeros(self.num_envs, device=dev)
        return State(data, self.compute_obs(data), zero, zero.clone(), {}, {})

    def step(self, state: State, action: torch.Tensor) -> State:
        data = self.pipeline_step(state.pipeline_state, action)
        done = (~(data.qpos[..., 2] >= self.z_min)).to(torch.float32)  # a non-finite height terminates too
        return state.replace(pipeline_state=data, obs=self.compute_obs(data), reward=self.compute_reward(data), done=done)


class FusedQuadraticTaskEnv:
    """`AutoResetWrapper(EpisodeWrapper(QuadraticTaskEnv, episode_length))` as ONE launch per env step.

    The state tensors are updated IN PLACE (the returned `State` aliases the env's buffers): this is the
    training hot loop, not a functional API. `step` never synchronises."""

    def __init__(self, env: QuadraticTaskEnv, episode_length: int, randomization: Optional[torch.Tensor] = None):
        """randomization: optional (num_envs, 2) tensor of per-env {contact friction scale, actuator strength scale} or (num_envs, 4)
        with {joint damping scale, joint armature scale} as well (abr_env_set_randomization_ex): every env then steps its own
        variant of the model."""
        self.randomization = randomization
        if not torch.equal(env.reward_fn.Q, torch.diag(torch.diagonal(env.reward_fn.Q))) or \
                not torch.equal(env.reward_fn.R, torch.diag(torch.diagonal(env.reward_fn.R))):
            raise NotImplementedError("the fused reward takes diagonal Q and R")
        self.env, self.episode_length = env, int(episode_length)

    @property
    def unwrapped(self):
        return self.env

    def reset(self, rng) -> State:
        E = self.env.num_envs
        if self.randomization is not None:  # installed before pipeline_init so the cached first state's warm start sees it too
            dev0 = mjx._dev(self.env._device)
            self._dr = self.randomization.to(device=dev0, dtype=torch.float32).reshape(E, -1).contiguous()
            mjx.set_randomization(self.env.sys, self._dr)
        s = self.env.reset(rng)
        d = s.pipeline_state
        dev = d.qpos.device
        self._first = tuple(t.contiguous().clone() for t in (d.qpos, d.qvel, d.qacc_warmstart))
        self._buf = dict(qpos=d.qpos.contiguous().clone(), qvel=d.qvel.contiguous().clone(), warm=d.qacc_warmstart.contiguous().clone(),
                         time=torch.zeros(E, device=dev), steps=torch.zeros(E, dtype=torch.int32, device=dev),
                         obs=s.obs.contiguous().clone(), reward=torch.zeros(E, device=dev),
                         done=torch.zeros(E, dtype=torch.uint8, device=dev), trunc=torch.zeros(E, dtype=torch.uint8, device=dev))
        self._h = self.env.sys.handle(dev.index or 0)
        self._ch = self.env.reward_fn.device_cost(dev.index or 0)
        return self._state(torch.zeros(E, self.env.sys.nu, device=dev))

    def _state(self, ctrl) -> State:
        b = self._buf
        data = mjx.Data(qpos=b["qpos"], qvel=b["qvel"], ctrl=ctrl, qacc=torch.zeros_like(b["qvel"]), qacc_warmstart=b["warm"], time=b["time"])
        return State(data, b["obs"], b["reward"], b["done"], {}, {"steps": b["steps"], "truncation": b["trunc"]})

    def step(self, state: Optional[State], action: torch.Tensor) -> State:
        b, e = self._buf, self.env
        dev = b["qpos"].device
        ctrl = action.to(torch.float32).contiguous()
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().abr_env_task_step_dev(
            self._h.ptr, p(b["qpos"]), p(b["qvel"]), p(b["warm"]), p(b["time"]), p(ctrl), e.num_envs, e._physics_steps_per_control_step,
            p(self._first[0]), p(self._first[1]), p(self._first[2]), self._ch.ptr, e.z_min, self.episode_length, p(b["steps"]), p(b["obs"]),
            p(b["reward"]), p(b["done"]), p(b["trunc"]), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return self._state(ctrl)
