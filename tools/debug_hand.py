import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.shooting import shoot
from oracle.oracle import Oracle
np.set_printoptions(precision=5, suppress=True, linewidth=200)
t32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")
mj = load_mj_model_from_file("models/barrett_hand/bh280.xml")
for flags in (16, 272):
    m = mjx.device_put(mj); m = m.replace(opt=m.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, integrator=0, solver=2, disableflags=flags))
    o = Oracle(mj, m.opt)
    rng = np.random.default_rng(31)
    W, N = 12, 25
    lo, hi = mj.jnt_range[:, 0], mj.jnt_range[:, 1]
    span = np.where(hi > lo, hi - lo, 1.0)
    x0 = np.concatenate([np.where(hi > lo, lo, -0.5) + span * rng.uniform(-0.15, 1.15, (W, mj.nq)), 0.5 * rng.normal(size=(W, mj.nv))], axis=1)
    us = rng.normal(size=(W, N, mj.nu)) * 1.5
    xs = shoot(m, t32(x0), t32(us)).cpu().numpy()
    m.set_lanes(8); gen = shoot(m, t32(x0), t32(us)).cpu().numpy(); m.set_lanes(0)
    ref = o.rollout(x0, us); ref32 = o.rollout(x0, us, prec=1)
    print("flags", flags, m.describe())
    for t in (1, 2, 5):
        print(f" t={t}: hand-f64 {np.abs(xs[:,t]-ref[:,t]).max(axis=1)}")
        print(f"       gen-f64  {np.abs(gen[:,t]-ref[:,t]).max(axis=1)}")
        print(f"       f32-f64  {np.abs(ref32[:,t]-ref[:,t]).max(axis=1)}")
    w = int(np.argmax(np.abs(xs[:,1]-ref[:,1]).max(axis=1)))
    print(" worst world", w, "x0", x0[w], "\n  hand", xs[w,1], "\n  gen ", gen[w,1], "\n  f64 ", ref[w,1], "\n  f32 ", ref32[w,1])
