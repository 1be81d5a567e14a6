"""Why does a sample of the reference's N(0,1) sampler test come out NaN on the device? Rebuilds that sample's controls with the
numpy Philox restatement and rolls it through the float32 / float64 oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams, shoot
from oracle.oracle import Oracle, quad_cost
from tests import _philox
DEV = torch.device("cuda")
mj = load_mj_model_from_file("models/barrett_hand/bh280.xml")
model = mjx.device_put(mj)
model = model.replace(opt=model.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, integrator=0, solver=2, disableflags=16))
o = Oracle(mj, model.opt)
nx = 16
cf = StaticGoalQuadraticCost(Q=torch.eye(nx), Qf=10.0 * torch.eye(nx), R=0.01 * torch.eye(4), xg=torch.zeros(nx))
ps = VanillaPredictiveSampler(model=model, cost_function=cf, nsamples=100, stdev=0.01)
g = torch.Generator(device=DEV).manual_seed(0)
x0 = torch.randn((10, nx), generator=g, device=DEV); ug = torch.randn((10, 10, 4), generator=g, device=DEV)
key = torch.tensor([0, 7])
xs, us, info = ps.optimize(VanillaPredictiveSamplerParams(key=key, x0=x0, us_guess=ug), return_info=True)
costs = info["costs"].cpu().numpy()
from ambersim_b200.trajopt.shooting import _seed_of
seed = _seed_of(key)
lim = mj.actuator_ctrlrange.astype(np.float32)
for b in range(10):
    bad = np.nonzero(~np.isfinite(costs[b]))[0]
    print(f"problem {b}: x0 qpos {np.round(x0[b,:8].cpu().numpy(),2)} nonfinite samples {bad.tolist()[:10]} of {len(bad)}; guess cost {costs[b,0]:.4g} min finite {np.nanmin(costs[b]):.4g}")
    for s_ in bad[:3]:
        z = _philox.normals(seed, int(s_), b, np.arange(40)).reshape(10, 4)
        u = np.clip(ug[b].cpu().numpy() + z * np.float32(0.01), lim[:, 0], lim[:, 1])
        a64 = o.rollout(x0[b].cpu().numpy().astype(np.float64), u.astype(np.float64))
        a32 = o.rollout(x0[b].cpu().numpy().astype(np.float64), u.astype(np.float64), prec=1)
        gx = shoot(model, x0[b], torch.tensor(u, device=DEV)).cpu().numpy()
        print(f"   sample {s_}: oracle f64 max|x| {np.abs(a64).max():.4g} finite {np.isfinite(a64).all()}; oracle f32 max|x| {np.nanmax(np.abs(a32)):.4g} finite {np.isfinite(a32).all()}; gpu finite per step {np.isfinite(gx).all(axis=1).astype(int).tolist()} max {np.nanmax(np.abs(gx)):.4g}")
        print("      f64 |x| per step", np.round(np.abs(a64).max(axis=1), 1).tolist())
        print("      f32 |x| per step", np.round(np.abs(a32).max(axis=1), 1).tolist())
        print("      gpu |x| per step", np.round(np.abs(gx).max(axis=1), 1).tolist())
        for lanes in (4, 8, 16, 32):
            model.set_lanes(lanes)
            gl = shoot(model, x0[b], torch.tensor(u, device=DEV)).cpu().numpy()
            print(f"      lanes {lanes}: finite per step {np.isfinite(gl).all(axis=1).astype(int).tolist()}")
        model.set_lanes(0)
