"""What do the limb / hand kernels' fast arithmetic forms (rcp.approx.ftz, sqrt.approx.ftz, the branch-free sincos; csrc/Makefile
FASTMATH) change and cost? Runs the same rollouts through the shipped library and through the `make precise` build
(ambersim_b200/libabr_precise.so: correctly rounded reciprocal / square root, libdevice sincosf, IEEE division) and compares both with
the float64 oracle and with each other; times the C2 and C3 launches with each.

usage: make -C ambersim_b200/csrc precise -j8 && python tools/precise_vs_fast.py     (needs a GPU; the oracle is the checker here)"""
import json, os, subprocess, sys, tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
MODELS = {"barkour": ("models/barkour_standin/barkour_vb_standin.xml", "home"), "biped": ("models/biped_standin/biped_exo_standin.xml", "stand")}


def child(out_path):
    import numpy as np, torch
    from ambersim_b200 import mjx
    from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
    from ambersim_b200.trajopt.shooting import _rollout, shoot
    from ambersim_b200.utils.io_utils import load_mj_model_from_file
    from oracle.oracle import Oracle

    res, arrays = {}, {}
    f = dict(dtype=torch.float32, device="cuda")
    for name, (path, key) in MODELS.items():
        mj = load_mj_model_from_file(path); m = mjx.device_put(mj)
        res[name + "_kernels"] = m.describe()
        rng = np.random.default_rng(0)
        W, N = 64, 20
        x0 = np.tile(np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)]), (W, 1))
        x0[:, 7:mj.nq] += rng.uniform(-0.05, 0.05, (W, mj.nq - 7))
        x0[:, mj.nq:] += 0.1 * rng.standard_normal((W, mj.nv))
        us = np.clip(mj.key_ctrl(key) + 0.1 * rng.standard_normal((W, N, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
        xs = shoot(m, torch.tensor(x0, **f), torch.tensor(us, **f)).cpu().numpy()
        ref = Oracle(mj).rollout(x0, us, nthreads=8)
        res[name + "_max_abs_err_vs_f64_oracle_64x20"] = float(np.abs(xs - ref).max())
        res[name + "_max_abs_err_vs_f64_oracle_first_step"] = float(np.abs(xs[:, 1] - ref[:, 1]).max())
        arrays[name] = xs
        # timing at the bench shapes
        Wb, Nb = (4096, 1000) if name == "barkour" else (16384, 1000)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        lim = torch.tensor(mj.actuator_ctrlrange, **f)
        usb = torch.clamp(torch.tensor(mj.key_ctrl(key), **f) + 0.1 * torch.randn((Wb, Nb, mj.nu), generator=g, **f), lim[:, 0], lim[:, 1])
        q0 = np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)])
        x0b = torch.tensor(q0, **f).repeat(Wb, 1)
        x0b[:, 7:mj.nq] += (torch.rand((Wb, mj.nq - 7), generator=g, **f) - 0.5) * 0.1
        nx = mj.nq + mj.nv
        cost = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
        best = 1e30
        for r in range(4):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); _, costs = _rollout(m, x0b, usb, cost, False, True); e1.record(); torch.cuda.synchronize()
            if r: best = min(best, e0.elapsed_time(e1))
        res[name + f"_{Wb}x{Nb}_ms"] = best
        res[name + f"_{Wb}x{Nb}_world_steps_per_s"] = Wb * Nb / best * 1e3
        res[name + "_costs_finite"] = bool(torch.isfinite(costs).all())
        arrays[name + "_costs"] = costs.cpu().numpy()
    np.savez(out_path + ".npz", **arrays)
    json.dump(res, open(out_path + ".json", "w"))


def main():
    import numpy as np
    libs = {"fast (shipped)": ROOT / "ambersim_b200" / "libabr.so", "precise": ROOT / "ambersim_b200" / "libabr_precise.so"}
    tmp = tempfile.mkdtemp()
    got = {}
    for tag, lib in libs.items():
        assert lib.exists(), f"{lib} is missing (make -C ambersim_b200/csrc precise)"
        out = os.path.join(tmp, tag.split()[0])
        subprocess.run([sys.executable, __file__, "--child", out], check=True, env=dict(os.environ, ABR_LIB=str(lib)))
        got[tag] = (json.load(open(out + ".json")), np.load(out + ".npz"))
    keys = list(got["precise"][0].keys())
    print(f"{'':58s} {'fast (shipped)':>16s} {'precise':>16s}")
    for k in keys:
        a, b = got["fast (shipped)"][0][k], got["precise"][0][k]
        if isinstance(a, float):
            print(f"{k:58s} {a:16.6g} {b:16.6g}" + (f"   ({100 * (a / b - 1):+.1f} %)" if k.endswith("per_s") else ""))
        elif isinstance(a, bool):
            print(f"{k:58s} {str(a):>16s} {str(b):>16s}")
    for name in MODELS:
        d = np.abs(got["fast (shipped)"][1][name] - got["precise"][1][name])
        print(f"{name}: max |fast - precise| over the 64 x 20-step rollouts = {d.max():.3e} (first step {d[:, 1].max():.3e})")
        ca, cb = got["fast (shipped)"][1][name + "_costs"], got["precise"][1][name + "_costs"]
        print(f"{name}: full-horizon costs, median relative difference fast vs precise = {np.median(np.abs(ca - cb) / np.abs(cb)):.3e} (chaotic over 1000 contact steps)")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()
