"""One fused-rollout launch of the C2 workload at a reduced horizon (for ncu captures).
usage: python tools/prof_rollout.py [lanes] [steps] [worlds] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import _rollout
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
W = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
model = sys.argv[5] if len(sys.argv) > 5 else "barkour"
path, key = {"barkour": ("models/barkour_standin/barkour_vb_standin.xml", "home"), "biped": ("models/biped_standin/biped_exo_standin.xml", "stand")}[model]
mj = load_mj_model_from_file(path); m = mjx.device_put(mj)
import os
if os.environ.get("ABR_PROF_DISABLE"):  # extra mjtDisableBit flags (e.g. 256 = no warm start, 1 = no constraints): smaller code footprint per step
    m = m.replace(opt=m.opt.replace(disableflags=int(m.opt.disableflags) | int(os.environ["ABR_PROF_DISABLE"])))
if os.environ.get("ABR_PROF_LS"):
    m = m.replace(opt=m.opt.replace(ls_iterations=int(os.environ["ABR_PROF_LS"])))
if lanes: m.set_lanes(lanes)
g = torch.Generator(device="cuda"); g.manual_seed(1)
f = dict(dtype=torch.float32, device="cuda")
lim = torch.tensor(mj.actuator_ctrlrange, **f)
us = torch.clamp(torch.tensor(mj.key_ctrl(key), **f) + 0.1 * torch.randn((W, N, mj.nu), generator=g, **f), lim[:, 0], lim[:, 1])
q0 = np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)])
x0 = torch.tensor(q0, **f).repeat(W, 1)
x0[:, 7:mj.nq] += (torch.rand((W, mj.nq - 7), generator=g, **f) - 0.5) * 0.1
nx = mj.nq + mj.nv
cost = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
for r in range(reps):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); _, costs = _rollout(m, x0, us, cost, False, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f"{model} {W}x{N} lanes={lanes}: {ms:.3f} ms -> {W*N/ms*1e3:.4e} world-steps/s ({ms*1e3/max(N,1):.1f} us/step) cost mean {float(costs.mean()):.4f}")
