"""Development check run on the GPU box: stage-by-stage parity vs the oracle + quick timings."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from ambersim_b200 import mjx, _lib
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from ambersim_b200.trajopt.shooting import shoot
from oracle.oracle import Oracle

np.set_printoptions(precision=5, suppress=True, linewidth=160)
MODELS = {"pendulum": ("models/pendulum/scene.xml", None), "bh280": ("models/barrett_hand/bh280.xml", None),
          "barkour": ("models/barkour_standin/barkour_vb_standin.xml", "home"),
          "biped": ("models/biped_standin/biped_exo_standin.xml", "stand")}
FIELDS = ["xpos", "xquat", "xipos", "cinert", "cdof", "qM", "cvel", "cdof_dot", "contact_dist", "contact_pos", "contact_frame",
          "qfrc_smooth", "qacc_smooth", "efc_J", "efc_D", "efc_aref", "qacc", "efc_force", "qfrc_constraint"]
rng = np.random.default_rng(0)
print("device", torch.cuda.get_device_name(0))
for name, (path, key) in MODELS.items():
    mj = load_mj_model_from_file(path)
    if name == "bh280":
        mj.opt = mj.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, disableflags=16)
    m = mjx.device_put(mj)
    o = Oracle(mj)
    q = mj.key_qpos(key) if key else mj.qpos0.copy()
    if name == "bh280": q = q + rng.uniform(0.0, 0.5, mj.nq)
    elif name == "pendulum": q = q + 0.7
    else: q[7:] += rng.uniform(-0.1, 0.1, mj.nq - 7); q[2] -= 0.004
    v = rng.normal(size=mj.nv) * 0.3
    c = (mj.key_ctrl(key) if key else np.zeros(mj.nu)) + rng.normal(size=mj.nu) * 0.1
    w = rng.normal(size=mj.nv)
    ref = o.forward(q, v, c, w)
    got = mjx.debug_forward(m, q, v, c, w, names=FIELDS)
    print(f"== {name}: nq={mj.nq} nv={mj.nv} nefc={o.nefc}")
    for f in FIELDS:
        r = ref[f].ravel(); g = got[f].ravel().astype(np.float64)
        if f == "subtree_com": continue
        if r.size != g.size:
            print(f"  {f:16s} SIZE MISMATCH {r.size} vs {g.size}"); continue
        if r.size == 0: continue
        err = np.abs(r - g).max(); scale = max(1e-9, np.abs(r).max())
        print(f"  {f:16s} maxabs {err:.3e}  rel {err/scale:.3e}")
    # rollout parity
    N = 50
    us = np.clip(c + 0.1 * rng.normal(size=(4, N, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
    x0 = np.concatenate([q, v])
    xs_ref = o.rollout(x0, us)
    xs_ref32 = o.rollout(x0, us, prec=1)
    for lanes in (4, 8, 16, 32):
        m.set_lanes(lanes)
        xs = shoot(m, torch.tensor(x0, dtype=torch.float32, device="cuda"), torch.tensor(us, dtype=torch.float32, device="cuda")).cpu().numpy()
        e = np.abs(xs - xs_ref).max(axis=(0, 2))
        print(f"  rollout G={lanes:2d}: |gpu-f64| t=1 {e[1]:.2e} t=10 {e[10]:.2e} t={N} {e[N]:.2e}   (oracle f32 vs f64 t={N}: {np.abs(xs_ref32-xs_ref).max(axis=(0,2))[N]:.2e})")
    xs_h = shoot(m, x0.astype(np.float32), us.astype(np.float32))
    print("  host-API rollout == device-API:", np.array_equal(xs_h, xs))
    m.set_lanes(0)

# ---- quick throughput on the Barkour stand-in (C2 shape, shortened)
mj = load_mj_model_from_file(MODELS["barkour"][0]); m = mjx.device_put(mj)
W, N = 4096, 200
c0 = mj.key_ctrl("home"); q0 = mj.key_qpos("home")
g = torch.Generator(device="cuda"); g.manual_seed(1)
lim = torch.tensor(mj.actuator_ctrlrange, dtype=torch.float32, device="cuda")
us = torch.clamp(torch.tensor(c0, dtype=torch.float32, device="cuda") + 0.1 * torch.randn((W, N, mj.nu), generator=g, device="cuda"), lim[:, 0], lim[:, 1])
x0 = torch.tensor(np.concatenate([q0, np.zeros(mj.nv)]), dtype=torch.float32, device="cuda").repeat(W, 1)
x0[:, 7:19] += (torch.rand((W, 12), generator=g, device="cuda") - 0.5) * 0.1
from ambersim_b200.trajopt.shooting import _rollout
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
cost = StaticGoalQuadraticCost(np.eye(mj.nq + mj.nv), 10 * np.eye(mj.nq + mj.nv), 0.01 * np.eye(mj.nu), np.concatenate([q0, np.zeros(mj.nv)]))
for lanes in (32, 16, 8, 4):
    m.set_lanes(lanes)
    for rep in range(2):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _, costs = _rollout(m, x0, us, cost, False, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"barkour {W}x{N} G={lanes:2d}: {ms:8.2f} ms  -> {W*N/ms/1e3:.3e} world-steps/s   cost mean {float(costs.mean()):.3f} nan {int(torch.isnan(costs).sum())}")
tf = __import__("ctypes").c_double(); ms = __import__("ctypes").c_double()
_lib.check(_lib.lib().abr_ffma_peak(0, tf, ms)); print("FFMA peak TFLOP/s", tf.value, "ms", ms.value)
