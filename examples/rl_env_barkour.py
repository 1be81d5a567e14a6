"""A batch of auto-resetting training environments stepped by ONE launch per env step (obs, reward, done, episode counter
and reset fused into the physics kernel), with per-env domain randomisation. Needs a CUDA device.

    python examples/rl_env_barkour.py [num_envs] [steps]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from ambersim_b200.rl.wrappers import FusedQuadraticTaskEnv, QuadraticTaskEnv
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.utils.io_utils import load_mj_model_from_file

E, T = (int(a) for a in (sys.argv[1:3] + ["8192", "1000"][len(sys.argv) - 1:]))
mj = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
nx = mj.nq + mj.nv
goal = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
reward = StaticGoalQuadraticCost(np.eye(nx), np.eye(nx), 0.01 * np.eye(mj.nu), goal)
g = torch.Generator(device="cuda").manual_seed(0)
dr = torch.stack((torch.empty(E, device="cuda").uniform_(0.5, 1.25, generator=g),      # contact friction scale
                  torch.empty(E, device="cuda").uniform_(0.8, 1.2, generator=g)), 1)   # actuator strength scale
env = FusedQuadraticTaskEnv(QuadraticTaskEnv(mj, reward, mj.key_qpos("home"), num_envs=E, z_min=0.15, jitter=0.05,
                                             physics_steps_per_control_step=4), episode_length=250, randomization=dr)
state = env.reset(0)
home = torch.tensor(mj.key_ctrl("home"), dtype=torch.float32, device="cuda")
lim = torch.tensor(mj.actuator_ctrlrange, dtype=torch.float32, device="cuda")
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
episodes = torch.zeros((), device="cuda")
ret = torch.zeros(E, device="cuda")
start.record()
for t in range(T):
    action = torch.clamp(home + 0.3 * torch.randn((E, mj.nu), generator=g, device="cuda"), lim[:, 0], lim[:, 1])  # a "policy"
    state = env.step(state, action)
    ret += state.reward
    episodes += state.done.sum()
end.record()
torch.cuda.synchronize()
ms = start.elapsed_time(end)
print(f"{E} envs x {T} env steps (4 physics steps each, random policy included): {E * T / ms * 1e3:.3e} env-steps/s, "
      f"{int(episodes)} episodes finished, mean reward per step {float(ret.mean()) / T:.3f}")
