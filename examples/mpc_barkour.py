"""Receding-horizon predictive sampling on the engine (what ambersim's VanillaPredictiveSampler.optimize is for):
a Barkour-class quadruped recovers its home stance from a scrambled leg configuration. Needs a CUDA device.

    python examples/mpc_barkour.py [nsamples] [horizon] [ticks]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from ambersim_b200 import mjx
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams
from ambersim_b200.utils.io_utils import load_mj_model_from_file

S, N, T = (int(a) for a in (sys.argv[1:4] + ["4096", "32", "200"][len(sys.argv) - 1:]))
mj = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
model = mjx.device_put(mj)  # same call as mjx.device_put(mujoco.MjModel) in the reference
nx = mj.nq + mj.nv
goal = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])   # stand still at the home pose ...
x0 = goal.copy()
rng = np.random.default_rng(0)
x0[7:mj.nq] += rng.uniform(-0.15, 0.15, mj.nq - 7)                # ... starting from a scrambled leg configuration
w = np.ones(nx)
w[mj.nq:] = 0.05
cost = StaticGoalQuadraticCost(np.diag(w), 10 * np.diag(w), 0.001 * np.eye(mj.nu), goal)
sampler = VanillaPredictiveSampler(model=model, cost_function=cost, nsamples=S, stdev=0.15)
f = dict(dtype=torch.float32, device="cuda")
params = VanillaPredictiveSamplerParams(key=0, x0=torch.tensor(x0, **f), us_guess=torch.tensor(mj.key_ctrl("home"), **f).repeat(N, 1))

# one solve, exactly the reference's call
xs_star, us_star = sampler.optimize(params)
print("one solve:", tuple(xs_star.shape), tuple(us_star.shape))

# the whole loop on the device: solve, one plant step under us*[0], shifted guess; no host round trips
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sampler.mpc(params, 5)
start.record()
xs, us, info = sampler.mpc(params, T)
end.record()
torch.cuda.synchronize()
err0 = float(np.abs(x0[7:mj.nq] - goal[7:mj.nq]).max())
err1 = float((xs[-1, 7:mj.nq].cpu() - torch.tensor(goal[7:mj.nq], dtype=torch.float32)).abs().max())
print(f"{T} MPC ticks of {S} samples x {N} steps: {start.elapsed_time(end) / T:.3f} ms per tick; "
      f"max joint error {err0:.3f} -> {err1:.3f}; winner cost {float(info['best_cost'][0]):.3f} -> {float(info['best_cost'][-1]):.3f}")
