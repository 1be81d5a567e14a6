"""Rollout rate of the convex-collision fixture on the generic kernels, with and without contacts (where the step's time goes).
usage: python tools/prof_convex.py [worlds] [steps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import _rollout
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mj = load_mj_model_from_file("tests/models/blocks.xml")
q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]); q0[2] = 0.33
nx = mj.nq + mj.nv
cost = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
f = dict(dtype=torch.float32, device="cuda")
lim = torch.tensor(mj.actuator_ctrlrange, **f)
us = torch.clamp(torch.tensor(mj.key_ctrl("home"), **f) + 0.2 * torch.randn((W, N, mj.nu), generator=g, **f), lim[:, 0], lim[:, 1])
x0 = torch.tensor(q0, **f).repeat(W, 1)
for name, flags in (("contacts on", 0), ("contacts off", 16)):
    m = mjx.device_put(mj)
    m = m.replace(opt=m.opt.replace(disableflags=int(m.opt.disableflags) | flags))
    for r in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _, costs = _rollout(m, x0, us, cost, False, True); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"blocks {W}x{N} {name}: {ms:.2f} ms -> {W*N/ms*1e3:.3e} world-steps/s ({m.describe()})")
