"""Stand-in CPU numbers for BASELINE.md's table: the float32 oracle port on all host cores, bounded samples of C1 / C3 / C4.
(test infrastructure timing, like bench.py's cpu_baseline leg).  usage: python tools/cpu_configs.py"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from oracle.oracle import Oracle, quad_cost

nth = len(os.sched_getaffinity(0))
rng = np.random.default_rng(0)
# C1: bh280, contacts off, 100 samples x horizon 10 (the reference test's configuration)
hj = load_mj_model_from_file("models/barrett_hand/bh280.xml")
o = Oracle(hj, hj.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, integrator=0, solver=2, disableflags=16))
nx = hj.nq + hj.nv
us = np.clip(0.01 * rng.standard_normal((100, 10, hj.nu)), hj.actuator_ctrlrange[:, 0], hj.actuator_ctrlrange[:, 1])
x0 = np.zeros((100, nx))
o.rollout(x0, us, prec=1, nthreads=nth)
t0 = time.perf_counter()
for _ in range(20):
    xs = o.rollout(x0, us, prec=1, nthreads=nth)
    c = quad_cost(xs, us, np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(hj.nu), np.zeros(nx)); int(np.argmin(c))
print(f"C1 bh280 VPS 100x10 on the CPU oracle port ({nth} threads): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per solve")
# C3: biped stand-in
bj = load_mj_model_from_file("models/biped_standin/biped_exo_standin.xml")
ob = Oracle(bj)
W, N = nth * 16, 200
x0 = np.tile(np.concatenate([bj.key_qpos("stand"), np.zeros(bj.nv)]), (W, 1))
us = np.clip(bj.key_ctrl("stand") + 0.1 * rng.standard_normal((W, N, bj.nu)), bj.actuator_ctrlrange[:, 0], bj.actuator_ctrlrange[:, 1])
t0 = time.perf_counter(); ob.rollout(x0, us, prec=1, nthreads=nth, return_xs=False); dt = time.perf_counter() - t0
print(f"C3 biped on the CPU oracle port ({nth} threads): {W * N / dt:.4e} world-steps/s ({W} worlds x {N} steps)")
# C4: Barkour 4096 samples x 32 steps (+ the warm-start forward): one solve
mj = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
om = Oracle(mj)
x0 = np.tile(np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]), (4096, 1))
us = np.clip(mj.key_ctrl("home") + 0.1 * rng.standard_normal((4096, 32, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
t0 = time.perf_counter(); om.rollout(x0, us, prec=1, nthreads=nth, return_xs=False); dt = time.perf_counter() - t0
print(f"C4 Barkour 4096x32 rollouts on the CPU oracle port ({nth} threads): {dt * 1e3:.1f} ms per solve ({4096 * 32 / dt:.4e} world-steps/s)")
