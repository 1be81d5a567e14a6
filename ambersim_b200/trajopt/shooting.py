"""shoot + vanilla predictive sampling on the CUDA engine (reference ambersim/trajopt/shooting.py).

Same names and argument meaning as the reference:
    shoot(m, x0, us) -> xs                                        (shooting.py:22-48)
    VanillaPredictiveSampler(model, cost_function, nsamples, stdev).optimize(params) -> (xs*, us*)
                                                                  (shooting.py:96-157)
The reference vmaps `shoot` over samples and materialises every trajectory; here one fused kernel
rolls all samples, evaluates the quadratic cost in-flight, and only the winner's trajectory leaves
the GPU. Leading batch dimensions stand in for jax.vmap.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional, Tuple, Union

import numpy as np
import torch

from ambersim_b200 import _lib, mjx
from ambersim_b200.trajopt.base import CostFunction, TrajectoryOptimizer, TrajectoryOptimizerParams
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost

Array = Union[torch.Tensor, np.ndarray]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _rollout(m: mjx.Model, x0: Array, us: Array, cost: Optional[StaticGoalQuadraticCost], want_xs: bool, want_cost: bool):
    """Common path of `shoot`: device tensors go through abr_rollout_dev on the current stream,
    host arrays through abr_rollout_host (copies inside the call)."""
    L = _lib.lib()
    nx, nu = m.nx, m.nu
    on_device = isinstance(us, torch.Tensor) and us.is_cuda
    if on_device:
        dev = us.device
        us_t = us.to(torch.float32)
        x0_t = torch.as_tensor(x0, dtype=torch.float32, device=dev)
        N = us_t.shape[-2]
        batch = tuple(us_t.shape[:-2]) if us_t.dim() - 2 >= x0_t.dim() - 1 else tuple(x0_t.shape[:-1])
        W = int(np.prod(batch)) if len(batch) else 1
        us_f = us_t.expand(*batch, N, nu).reshape(W, N, nu).contiguous()
        if x0_t.dim() == 1:
            x0_f, stride = x0_t.contiguous(), 0
        else:
            x0_f, stride = x0_t.expand(*batch, nx).reshape(-1, nx).contiguous(), nx
        h = m.handle(dev.index or 0)
        xs = torch.empty((W, N + 1, nx), dtype=torch.float32, device=dev) if want_xs else None
        costs = torch.empty((W,), dtype=torch.float32, device=dev) if want_cost else None
        ch = cost.device_cost(dev.index or 0).ptr if want_cost else None
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.abr_rollout_dev(h.ptr, _ptr(x0_f), stride, _ptr(us_f), N * nu, W, N, _ptr(xs), ch, _ptr(costs), stream))
        return (xs.reshape(*batch, N + 1, nx) if want_xs else None), (costs.reshape(batch) if want_cost else None)
    us_n = np.ascontiguousarray(us.cpu().numpy() if isinstance(us, torch.Tensor) else us, dtype=np.float32)
    x0_n = np.ascontiguousarray(x0.cpu().numpy() if isinstance(x0, torch.Tensor) else x0, dtype=np.float32)
    N = us_n.shape[-2]
    batch = us_n.shape[:-2] if us_n.ndim - 2 >= x0_n.ndim - 1 else x0_n.shape[:-1]
    W = int(np.prod(batch)) if len(batch) else 1
    us_f = np.ascontiguousarray(np.broadcast_to(us_n, (*batch, N, nu)).reshape(W, N, nu))
    if x0_n.ndim == 1:
        x0_f, stride = x0_n, 0
    else:
        x0_f, stride = np.ascontiguousarray(np.broadcast_to(x0_n, (*batch, nx)).reshape(-1, nx)), nx
    h = m.handle()
    xs = np.empty((W, N + 1, nx), dtype=np.float32) if want_xs else None
    costs = np.empty((W,), dtype=np.float32) if want_cost else None
    ch = cost.device_cost(h.device).ptr if want_cost else None
    _lib.check(L.abr_rollout_host(h.ptr, _np_ptr(x0_f), stride, _np_ptr(us_f), N * nu, W, N, _np_ptr(xs), ch, _np_ptr(costs)))
    return (xs.reshape(*batch, N + 1, nx) if want_xs else None), (costs.reshape(batch) if want_cost else None)


def shoot(m: mjx.Model, x0: Array, us: Array) -> Array:
    """Rolls the model forward from x0 under the zero-order-hold controls us.

    Args:
        m: the model.
        x0 (shape=(nq+nv,) or (..., nq+nv)): initial state(s), x = [qpos; qvel].
        us (shape=(N, nu) or (..., N, nu)): control sequence(s).

    Returns:
        xs (shape=(..., N+1, nq+nv)): state trajectory, row 0 is x0 verbatim.
    """
    xs, _ = _rollout(m, x0, us, None, True, False)
    return xs


def shoot_cost(m: mjx.Model, x0: Array, us: Array, cost_function: StaticGoalQuadraticCost) -> Array:
    """Fused `cost(shoot(m, x0, us), us)` without materialising xs (engine extension)."""
    _, costs = _rollout(m, x0, us, cost_function, False, True)
    return costs


@dataclasses.dataclass
class ShootingParams(TrajectoryOptimizerParams):
    """Inputs of shooting methods (reference shooting.py:58-73)."""

    x0: Array = None  # shape=(nq+nv,) or (B, nq+nv)
    us_guess: Array = None  # shape=(N, nu) or (B, N, nu)

    @property
    def N(self) -> int:
        """Number of time steps (zero-order-hold parameterisation)."""
        return self.us_guess.shape[-2]


@dataclasses.dataclass
class ShootingAlgorithm(TrajectoryOptimizer):
    """A shooting-based trajectory optimiser (reference shooting.py:76-90)."""

    def optimize(self, params: ShootingParams) -> Tuple[Array, Array]:
        raise NotImplementedError


@dataclasses.dataclass
class VanillaPredictiveSamplerParams(ShootingParams):
    """Inputs of predictive sampling (reference shooting.py:96-100).

    key: random seed: an int, or an integer array/tensor (a JAX-PRNGKey-like uint32[2], or (B,2)).
    noise: optional standard normals (S-1, N, nu) / (B, S-1, N, nu) supplied by the caller instead of
        the engine's counter-based generator ("parity mode": pass jax.random.normal's values).
    """

    key: Union[int, Array] = 0
    noise: Optional[Array] = None


def _seed_of(key) -> int:
    """An int, or ONE jax-style key (two uint32 words). A batch of keys is rejected: batched problems draw independent
    noise from the one key through the problem index of the counter (the reference vmaps over split keys instead)."""
    if isinstance(key, (int, np.integer)):
        return int(key) & 0xFFFFFFFFFFFFFFFF
    k = key.detach().cpu().numpy() if isinstance(key, torch.Tensor) else np.asarray(key)
    if k.size > 2:
        raise ValueError(f"key must be an int or one (2,) key, got shape {tuple(k.shape)}: pass one key for the whole batch")
    k = k.reshape(-1).astype(np.uint64)
    if k.size == 1:
        return int(k[0])
    return int((k[0] << np.uint64(32)) | (k[1] & np.uint64(0xFFFFFFFF)))


@dataclasses.dataclass
class VanillaPredictiveSampler(ShootingAlgorithm):
    """Vanilla predictive sampling (reference shooting.py:103-157): fixed model, ZOH controls,
    isotropic normal noise around the guess (sample 0 is the guess itself), quadratic cost."""

    model: mjx.Model = None
    cost_function: CostFunction = None
    nsamples: int = 1
    stdev: float = 0.0

    def optimize(self, params: VanillaPredictiveSamplerParams, sample_offset: int = 0, nsamples_total: Optional[int] = None,
                 return_info: bool = False):
        """Returns (xs_star (N+1, nq+nv), us_star (N, nu)); batched inputs give batched outputs.

        sample_offset / nsamples_total: this call evaluates global samples
        [sample_offset, sample_offset + nsamples) of nsamples_total (multi-GPU sharding; the noise is
        keyed by the global sample id, so the union over ranks equals the single-GPU result).
        return_info: also return dict(best_idx (global), best_cost, costs).
        """
        m, S, L = self.model, int(self.nsamples), _lib.lib()
        S_total = S if nsamples_total is None else int(nsamples_total)
        nx, nu = m.nx, m.nu
        if not isinstance(self.cost_function, StaticGoalQuadraticCost):
            return self._optimize_generic(params, return_info)
        on_device = isinstance(params.us_guess, torch.Tensor) and params.us_guess.is_cuda
        seed = _seed_of(params.key)
        ug_nd = params.us_guess.dim() if isinstance(params.us_guess, torch.Tensor) else np.ndim(params.us_guess)
        if ug_nd not in (2, 3):
            raise ValueError(f"us_guess must be (N, nu) or (B, N, nu), got {ug_nd} dimensions")
        if on_device:
            dev = params.us_guess.device
            ug = params.us_guess.to(torch.float32)
            x0 = torch.as_tensor(params.x0, dtype=torch.float32, device=dev)
            batched = ug.dim() == 3
            N = ug.shape[-2]
            ug_f = ug.reshape(-1, N, nu).contiguous()
            B = ug_f.shape[0]
            x0_f = x0.expand(B, nx).contiguous() if x0.dim() == 1 else x0.reshape(-1, nx).contiguous()
            nz = None
            if params.noise is not None:
                nz = torch.as_tensor(params.noise, dtype=torch.float32, device=dev).expand(B, S_total - 1, N, nu).contiguous()
            h = m.handle(dev.index or 0)
            f = dict(dtype=torch.float32, device=dev)
            xs_star, us_star = torch.empty((B, N + 1, nx), **f), torch.empty((B, N, nu), **f)
            best_idx, best_cost = torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), **f)
            costs = torch.empty((B, S), **f) if return_info else None
            ch = self.cost_function.device_cost(dev.index or 0).ptr
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.abr_predictive_sample_dev(h.ptr, ch, _ptr(x0_f), _ptr(ug_f), _ptr(nz), seed, B, S, N, float(self.stdev),
                                                   int(sample_offset), S_total, _ptr(xs_star), _ptr(us_star), _ptr(best_idx),
                                                   _ptr(best_cost), _ptr(costs), stream))
        else:
            ug = np.ascontiguousarray(_to_np(params.us_guess), dtype=np.float32)
            x0 = np.ascontiguousarray(_to_np(params.x0), dtype=np.float32)
            batched = ug.ndim == 3
            N = ug.shape[-2]
            ug_f = np.ascontiguousarray(ug.reshape(-1, N, nu))
            B = ug_f.shape[0]
            x0_f = np.ascontiguousarray(np.broadcast_to(x0, (B, nx)) if x0.ndim == 1 else x0.reshape(-1, nx))
            nz = None
            if params.noise is not None:
                nz = np.ascontiguousarray(np.broadcast_to(_to_np(params.noise), (B, S_total - 1, N, nu)), dtype=np.float32)
            h = m.handle()
            xs_star, us_star = np.empty((B, N + 1, nx), np.float32), np.empty((B, N, nu), np.float32)
            best_idx, best_cost = np.empty((B,), np.int32), np.empty((B,), np.float32)
            costs = np.empty((B, S), np.float32) if return_info else None
            ch = self.cost_function.device_cost(h.device).ptr
            _lib.check(L.abr_predictive_sample_host(h.ptr, ch, _np_ptr(x0_f), _np_ptr(ug_f), _np_ptr(nz), seed, B, S, N,
                                                    float(self.stdev), int(sample_offset), S_total, _np_ptr(xs_star),
                                                    _np_ptr(us_star), _np_ptr(best_idx), _np_ptr(best_cost), _np_ptr(costs)))
        if not batched:
            xs_star, us_star = xs_star[0], us_star[0]
        if return_info:
            info = dict(best_idx=best_idx if batched else best_idx[0], best_cost=best_cost if batched else best_cost[0],
                        costs=costs if batched else costs[0])
            return xs_star, us_star, info
        return xs_star, us_star

    def mpc(self, params: VanillaPredictiveSamplerParams, nticks: int):
        """Receding-horizon loop on the device (engine extension; how `optimize` is driven in practice): nticks x
        {solve from the current state with key + tick; the plant takes one physics step under us*[0]; the guess
        becomes us* shifted by one step}. Equivalent to calling `optimize` in a Python loop, without the host
        round trips. Returns (xs (nticks+1, nx) visited states, us (nticks, nu) applied controls, info)."""
        m, L = self.model, _lib.lib()
        if not isinstance(self.cost_function, StaticGoalQuadraticCost):
            raise NotImplementedError("mpc() needs a StaticGoalQuadraticCost (fused on the device)")
        dev = params.us_guess.device if isinstance(params.us_guess, torch.Tensor) and params.us_guess.is_cuda else mjx._dev()
        f = dict(dtype=torch.float32, device=dev)
        ug = torch.as_tensor(params.us_guess, **f).clone().contiguous()
        x = torch.as_tensor(params.x0, **f).clone().contiguous()
        if ug.dim() != 2 or x.dim() != 1:
            raise ValueError("mpc() takes one problem: x0 (nx,), us_guess (N, nu)")
        N = ug.shape[0]
        xs_log, us_log = torch.empty((nticks + 1, m.nx), **f), torch.empty((nticks, m.nu), **f)
        cost_log, idx_log = torch.empty((nticks,), **f), torch.empty((nticks,), dtype=torch.int32, device=dev)
        h = m.handle(dev.index or 0)
        ch = self.cost_function.device_cost(dev.index or 0).ptr
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.abr_mpc_dev(h.ptr, ch, _ptr(x), _ptr(ug), _seed_of(params.key), int(self.nsamples), N, float(self.stdev), int(nticks),
                                 _ptr(xs_log), _ptr(us_log), _ptr(cost_log), _ptr(idx_log), stream))
        return xs_log, us_log, dict(best_cost=cost_log, best_idx=idx_log, us_guess=ug, x=x)

    def _optimize_generic(self, params: VanillaPredictiveSamplerParams, return_info: bool):
        """User-defined CostFunction: sample + clip on the host framework, roll every sample on the
        engine, evaluate `cost_function.cost` on the returned trajectories (reference order,
        shooting.py:140-156). Unbatched problems only."""
        m, S = self.model, int(self.nsamples)
        ug = torch.as_tensor(params.us_guess, dtype=torch.float32)
        dev = ug.device if ug.is_cuda else mjx._dev()
        ug = ug.to(dev)
        N = ug.shape[-2]
        if params.noise is not None:
            nz = torch.as_tensor(params.noise, dtype=torch.float32, device=dev)
        else:
            g = torch.Generator(device=dev)
            g.manual_seed(_seed_of(params.key) & 0x7FFFFFFFFFFFFFFF)
            nz = torch.randn((S - 1, N, m.nu), generator=g, device=dev)
        noise = torch.cat((torch.zeros((1, N, m.nu), device=dev), nz * self.stdev), dim=0)
        lim = torch.as_tensor(np.asarray(m.actuator_ctrlrange), dtype=torch.float32, device=dev)
        us_samples = torch.clamp(ug + noise, lim[:, 0], lim[:, 1])
        xs_samples = shoot(m, torch.as_tensor(params.x0, dtype=torch.float32, device=dev), us_samples)
        costs = torch.stack([self.cost_function.cost(xs_samples[i], us_samples[i], None)[0] for i in range(S)])
        costs_nan_first = torch.where(torch.isnan(costs), torch.full_like(costs, -float("inf")), costs)
        best = int(torch.argmin(costs_nan_first))
        if return_info:
            return xs_samples[best], us_samples[best], dict(best_idx=best, best_cost=costs[best], costs=costs)
        return xs_samples[best], us_samples[best]


@dataclasses.dataclass
class SplineShootingParams(VanillaPredictiveSamplerParams):
    """`us_guess` holds K control KNOTS (K, nu) / (B, K, nu), not N zero-order-hold controls: the alternate parameterisation
    `ShootingParams.N` anticipates (reference shooting.py:66-73), so the horizon is given explicitly."""

    horizon: int = 0

    @property
    def N(self) -> int:
        return int(self.horizon)


@dataclasses.dataclass
class SplinePredictiveSampler(VanillaPredictiveSampler):
    """Predictive sampling over control knots (engine extension): the K knots are spread evenly over the N-step
    horizon and interpolated ("zoh", "linear" or "cubic" Catmull-Rom) into the per-step controls; the Gaussian noise is drawn
    on the KNOTS (stdev per knot entry, sample 0 = the un-noised knots) and interpolated the same way, so a sample's controls
    are interp(knots + stdev z) clipped to actuator_ctrlrange. The rollouts, costs, argmin and winner run fused on the engine
    exactly as for `VanillaPredictiveSampler` (through its caller-supplied-noise path). Returns (xs_star, knots_star)."""

    interp: str = "linear"

    def weights(self, K: int, N: int, device=None) -> torch.Tensor:
        """(N, K) interpolation matrix: row t holds the weights of the knots in the control applied at step t."""
        W = torch.zeros((N, K), dtype=torch.float64)
        if K == 1:
            W[:, 0] = 1.0
        else:
            for t in range(N):
                s = t * (K - 1) / max(1, N - 1) if N > 1 else 0.0  # knot coordinate of step t: first knot at t = 0, last at t = N - 1
                k = min(int(np.floor(s)), K - 2)
                a = s - k
                if self.interp == "zoh":
                    W[t, k if a < 1.0 else k + 1] = 1.0
                elif self.interp == "linear":
                    W[t, k] += 1.0 - a
                    W[t, k + 1] += a
                elif self.interp == "cubic":  # Catmull-Rom through the knots, end knots repeated
                    c = [(-a**3 + 2 * a**2 - a) / 2, (3 * a**3 - 5 * a**2 + 2) / 2, (-3 * a**3 + 4 * a**2 + a) / 2, (a**3 - a**2) / 2]
                    for j, cj in zip((k - 1, k, k + 1, k + 2), c):
                        W[t, min(max(j, 0), K - 1)] += cj
                else:
                    raise ValueError(f"interp must be 'zoh', 'linear' or 'cubic', got {self.interp!r}")
        return W.to(dtype=torch.float32, device=device)

    def optimize(self, params: SplineShootingParams, return_info: bool = False):
        if not isinstance(self.cost_function, StaticGoalQuadraticCost):
            raise NotImplementedError("SplinePredictiveSampler needs a StaticGoalQuadraticCost (fused on the device)")
        knots = torch.as_tensor(params.us_guess, dtype=torch.float32)
        dev = knots.device if knots.is_cuda else mjx._dev()
        knots = knots.to(dev)
        batched = knots.dim() == 3
        kb = knots if batched else knots[None]
        B, K, nu = kb.shape
        N, S = int(params.N), int(self.nsamples)
        if N <= 0:
            raise ValueError("SplineShootingParams.horizon must be the number of physics steps")
        W = self.weights(K, N, dev)
        g = torch.Generator(device=dev)
        g.manual_seed(_seed_of(params.key) & 0x7FFFFFFFFFFFFFFF)
        z = torch.randn((B, max(S - 1, 1), K, nu), generator=g, dtype=torch.float32, device=dev)  # row s - 1 = sample s (a spare row when S = 1)
        dense = dataclasses.replace(params, us_guess=torch.einsum("tk,bku->btu", W, kb), noise=torch.einsum("tk,bsku->bstu", W, z[:, :S - 1]),
                                    x0=torch.as_tensor(params.x0, dtype=torch.float32, device=dev).expand(B, self.model.nx))
        vanilla = VanillaPredictiveSampler(model=self.model, cost_function=self.cost_function, nsamples=S, stdev=self.stdev)
        xs, us, info = vanilla.optimize(VanillaPredictiveSamplerParams(x0=dense.x0, us_guess=dense.us_guess, key=params.key, noise=dense.noise),
                                        return_info=True)
        idx = info["best_idx"].long()
        pick = z[torch.arange(B, device=dev), (idx - 1).clamp(min=0)] * (idx > 0).float()[:, None, None]
        knots_star = kb + float(self.stdev) * pick
        if not batched:
            xs, us, knots_star = xs[0], us[0], knots_star[0]
            info = {k: (v[0] if hasattr(v, "dim") and v.dim() > 0 else v) for k, v in info.items()}
        if return_info:
            info["us_star"] = us
            return xs, knots_star, info
        return xs, knots_star


@dataclasses.dataclass
class FiniteDifferenceShooting(ShootingAlgorithm):
    """Gradient-based single shooting on the engine (extension; the reference shapes its API for it, base.py:67-69 and
    cost.py:87-178, without shipping one): the gradient of the trajectory cost with respect to every control entry is a
    central difference over 2 N nu perturbed rollouts (ONE engine launch), the step is the best of a geometric ladder of
    step sizes along the clipped negative gradient (a second launch), and the iterate only moves when the cost drops.
    Zero-order-hold controls, clipped to actuator_ctrlrange like the sampler (shooting.py:146-148)."""

    model: mjx.Model = None
    cost_function: CostFunction = None
    iterations: int = 10
    eps: float = 1e-3
    nsteps: int = 16          # candidate step sizes per iteration: alpha0 * 2^-k
    alpha0: float = 1.0       # largest step, in units of the mean control range (or 1 for unlimited controls)

    def _limits(self, dev):
        lim = torch.as_tensor(np.asarray(self.model.actuator_ctrlrange), dtype=torch.float32, device=dev)
        limited = torch.as_tensor(np.asarray(self.model.actuator_ctrllimited) > 0, device=dev)
        lo = torch.where(limited, lim[:, 0], torch.full_like(lim[:, 0], -float("inf")))
        hi = torch.where(limited, lim[:, 1], torch.full_like(lim[:, 1], float("inf")))
        width = torch.where(limited, lim[:, 1] - lim[:, 0], torch.ones_like(lo))
        return lo, hi, float(width.mean())

    def gradient(self, x0: Array, us: Array) -> Tuple[torch.Tensor, torch.Tensor]:
        """(cost(us), d cost / d us (N, nu)) by central differences: 2 N nu + 1 rollouts in one launch."""
        if not isinstance(self.cost_function, StaticGoalQuadraticCost):
            raise NotImplementedError("FiniteDifferenceShooting needs a StaticGoalQuadraticCost (fused on the device)")
        us = torch.as_tensor(us, dtype=torch.float32)
        dev = us.device if us.is_cuda else mjx._dev()
        us = us.to(dev)
        x0 = torch.as_tensor(x0, dtype=torch.float32, device=dev)
        N, nu = us.shape
        eye = torch.eye(N * nu, dtype=torch.float32, device=dev).reshape(N * nu, N, nu) * float(self.eps)
        batch = torch.cat((us[None], us[None] + eye, us[None] - eye), dim=0)
        costs = shoot_cost(self.model, x0, batch, self.cost_function)
        g = (costs[1:1 + N * nu] - costs[1 + N * nu:]) / (2.0 * float(self.eps))
        return costs[0], g.reshape(N, nu)

    def optimize(self, params: ShootingParams, return_info: bool = False):
        """Returns (xs_star (N+1, nq+nv), us_star (N, nu)) after `iterations` gradient steps from `us_guess`."""
        us = torch.as_tensor(params.us_guess, dtype=torch.float32)
        dev = us.device if us.is_cuda else mjx._dev()
        x0 = torch.as_tensor(params.x0, dtype=torch.float32, device=dev)
        lo, hi, width = self._limits(dev)
        us = torch.clamp(us.to(dev), lo, hi)
        history = []
        alphas = float(self.alpha0) * width * (0.5 ** torch.arange(int(self.nsteps), dtype=torch.float32, device=dev))
        for _ in range(int(self.iterations)):
            cost, g = self.gradient(x0, us)
            history.append(cost)
            gmax = g.abs().max()
            direction = torch.where(gmax > 0, g / gmax, torch.zeros_like(g))
            cand = torch.clamp(us[None] - alphas[:, None, None] * direction[None], lo, hi)
            costs = shoot_cost(self.model, x0, torch.cat((us[None], cand), dim=0), self.cost_function)
            key = torch.where(torch.isnan(costs), torch.full_like(costs, float("inf")), costs)  # a diverged candidate never wins
            best = torch.argmin(key)  # index 0 = stay: the cost never increases
            us = torch.cat((us[None], cand), dim=0)[best]
        xs = shoot(self.model, x0, us)
        if return_info:
            history.append(shoot_cost(self.model, x0, us[None], self.cost_function)[0])
            return xs, us, dict(costs=torch.stack(history))
        return xs, us


def _to_np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
