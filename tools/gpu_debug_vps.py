import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams, shoot, shoot_cost
from oracle.oracle import Oracle, quad_cost
mj = load_mj_model_from_file("models/barrett_hand/bh280.xml")
m = mjx.device_put(mj); m = m.replace(opt=m.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, disableflags=16))
cf = StaticGoalQuadraticCost(torch.eye(16), 10*torch.eye(16), 0.01*torch.eye(4), torch.zeros(16))
ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=100, stdev=0.01)
g = torch.Generator(device="cuda").manual_seed(0)
B, N = 10, 10
x0 = torch.randn((B, 16), generator=g, device="cuda"); ug = torch.randn((B, N, 4), generator=g, device="cuda")
xs, us, info = ps.optimize(VanillaPredictiveSamplerParams(key=torch.tensor([0, 7]), x0=x0, us_guess=ug), return_info=True)
cs, _ = cf.cost(xs, us, None)
xg = shoot(m, x0, ug); cg, _ = cf.cost(xg, ug, None)
print("best_idx", info["best_idx"].tolist())
print("best_cost (device)", info["best_cost"].tolist())
print("cost(xs*,us*) torch ", cs.tolist())
print("cost(guess)   torch ", cg.tolist())
print("device cost of sample 0", info["costs"][:, 0].tolist())
print("fused shoot_cost(guess)", shoot_cost(m, x0, ug, cf).tolist())
o = Oracle(mj, m.opt)
ref = o.rollout(x0.cpu().numpy().astype(np.float64), ug.cpu().numpy().astype(np.float64))
print("oracle cost(guess)", quad_cost(ref, ug.cpu().numpy(), np.eye(16), 10*np.eye(16), 0.01*np.eye(4), 0.0).tolist())
print("max |xs_guess - oracle|", np.abs(xg.cpu().numpy() - ref).max(axis=(1, 2)))
print("nan in costs:", int(torch.isnan(info["costs"]).sum()))
