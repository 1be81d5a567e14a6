"""CTA-size sweep of the limb rollout kernels (ABR_LIMB_TPB probe): the same launch at 1..8 warps per CTA.
usage: python tools/tpb_sweep.py [model worlds steps]...   (default: the C2 / 65536-world / C3 shapes, shortened)"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import _rollout

MODELS = {"barkour": ("models/barkour_standin/barkour_vb_standin.xml", "home"), "biped": ("models/biped_standin/biped_exo_standin.xml", "stand"),
          "exo": ("models/biped_standin/exo_legs_standin.xml", "stand")}
cfgs = [("barkour", 4096, 400), ("barkour", 8192, 200), ("barkour", 16384, 200), ("barkour", 65536, 100), ("biped", 16384, 100), ("exo", 16384, 100)]
if len(sys.argv) > 3:
    a = sys.argv[1:]
    cfgs = [(a[i], int(a[i + 1]), int(a[i + 2])) for i in range(0, len(a) - 2, 3)]
tpbs = [int(t) for t in os.environ.get("SWEEP_TPBS", "32,64,96,128,160,192,224,256").split(",")]
f = dict(dtype=torch.float32, device="cuda")
for model, W, N in cfgs:
    path, key = MODELS[model]
    mj = load_mj_model_from_file(path); m = mjx.device_put(mj)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    lim = torch.tensor(mj.actuator_ctrlrange, **f)
    us = torch.clamp(torch.tensor(mj.key_ctrl(key), **f) + 0.1 * torch.randn((W, N, mj.nu), generator=g, **f), lim[:, 0], lim[:, 1])
    q0 = np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)])
    x0 = torch.tensor(q0, **f).repeat(W, 1)
    x0[:, 7:mj.nq] += (torch.rand((W, mj.nq - 7), generator=g, **f) - 0.5) * 0.1
    nx = mj.nq + mj.nv
    cost = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
    ref = None
    for tpb in tpbs:
        os.environ["ABR_LIMB_TPB"] = str(tpb)
        best = 1e30
        for r in range(4):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _, costs = _rollout(m, x0, us, cost, False, True); e1.record(); torch.cuda.synchronize()
            if r: best = min(best, e0.elapsed_time(e1))
        if ref is None: ref = costs.clone()
        same = bool(torch.equal(ref, costs))
        print(f"{model:8s} {W:6d}x{N:4d} tpb {tpb:3d} grid {((W << int(np.log2(max(1, m.lanes_per_world() if hasattr(m, 'lanes_per_world') else 4)))) + tpb - 1) // tpb:5d}: {best:8.3f} ms -> {W * N / best * 1e3:.4e} world-steps/s  bit-equal to tpb {tpbs[0]}: {same}", flush=True)
    os.environ.pop("ABR_LIMB_TPB", None)
