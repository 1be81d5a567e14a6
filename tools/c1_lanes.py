"""C1 (the reference test's own configuration: bh280, contacts off, 100 samples x horizon 10) solve latency by generic group size.
usage: python tools/c1_lanes.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams
from ambersim_b200.utils.io_utils import load_mj_model_from_file
hj = load_mj_model_from_file("models/barrett_hand/bh280.xml")
f = dict(dtype=torch.float32, device="cuda")
for lanes in (0, 4, 8, 16, 32):
    for S, N in ((100, 10), (4096, 32)):
        hm = mjx.device_put(hj)
        hm = hm.replace(opt=hm.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, integrator=0, solver=2, disableflags=16))
        if lanes: hm.set_lanes(lanes)
        nx = hj.nq + hj.nv
        cf = StaticGoalQuadraticCost(np.eye(nx), 10.0 * np.eye(nx), 0.01 * np.eye(hj.nu), np.zeros(nx))
        ps = VanillaPredictiveSampler(model=hm, cost_function=cf, nsamples=S, stdev=0.01)
        prm = VanillaPredictiveSamplerParams(key=0, x0=torch.zeros(nx, **f), us_guess=torch.zeros((N, hj.nu), **f))
        for _ in range(3): ps.optimize(prm)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ps.optimize(prm)
        b.record(); torch.cuda.synchronize()
        print(f"lanes {lanes:2d}  {S} x {N}: {a.elapsed_time(b)/10:.3f} ms")
