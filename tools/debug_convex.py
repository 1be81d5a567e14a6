"""Engine vs oracle on the convex-pair fixture over many random states: classifies every pair whose contacts equal neither the float64
nor the float32 oracle (development aid). A GPU point is `explained` when one of the two oracles reports an active point at the same
place with the same depth (the manifold then differs only in WHICH of the clipped points were kept)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils import mjcf
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from oracle.oracle import Oracle
np.set_printoptions(precision=5, suppress=True, linewidth=200)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 600
mj = load_mj_model_from_file("tests/models/blocks.xml"); m = mjx.device_put(mj); o = Oracle(mj)
rng = np.random.default_rng(5)
qs = np.tile(mj.key_qpos("home"), (E, 1))
qs[:, 0] = rng.uniform(-0.2, 0.45, E); qs[:, 1] = rng.uniform(-0.2, 0.2, E); qs[:, 2] = rng.uniform(0.3, 0.42, E)
quat = np.array([1, 0, 0, 0]) + 0.2 * rng.normal(size=(E, 4)); qs[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
qs[:, 7:] += rng.uniform(-0.6, 0.6, (E, 3))
vs = 0.3 * rng.normal(size=(E, mj.nv)); cs = mj.key_ctrl("home") + 0.2 * rng.normal(size=(E, mj.nu))
t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
d = mjx.Data(qpos=t(qs), qvel=t(vs), ctrl=t(cs), qacc=torch.zeros((E, mj.nv), device="cuda"), qacc_warmstart=torch.zeros((E, mj.nv), device="cuda"), time=torch.zeros(E, device="cuda"))
f = mjx.forward(m, d, fields=("contact_dist", "contact_pos", "contact_frame"))
spans = []; c0 = 0
for k in mj.pair_kind:
    spans.append((c0, c0 + mjcf.PAIR_NCON[int(k)], int(k))); c0 += mjcf.PAIR_NCON[int(k)]
def close(r, a, b, gd, gp, gf, tol):
    return np.abs(r["contact_dist"][a:b] - gd[a:b]).max() < tol and np.abs(r["contact_pos"][a:b] - gp[a:b]).max() < tol and np.abs(r["contact_frame"][a:b] - gf[a:b]).max() < 10 * tol
n_neither = n_unexplained = n_normal = 0
for e in range(E):
    ref = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv))
    gd, gp, gf = f.contact_dist[e].cpu().numpy(), f.contact_pos[e].cpu().numpy(), f.contact_frame[e].cpu().numpy()
    r32 = None
    for a, b, kind in spans:
        rd = ref["contact_dist"][a:b]
        if not (rd < 0).any() and not (gd[a:b] < 0).any(): continue
        if close(ref, a, b, gd, gp, gf, 5e-6): continue
        if r32 is None: r32 = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv), prec=1)
        if close(r32, a, b, gd, gp, gf, 2e-5): continue
        n_neither += 1
        if np.abs(gf[a, 0] - ref["contact_frame"][a, 0]).max() > 1e-3 and np.abs(gf[a, 0] - r32["contact_frame"][a, 0]).max() > 1e-3:
            n_normal += 1
            print(f"world {e} kind {kind}: NORMAL differs from both oracles", gf[a, 0], ref["contact_frame"][a, 0], r32["contact_frame"][a, 0], gd[a:b], rd)
            continue
        cand = [(r["contact_pos"][i], r["contact_dist"][i]) for r in (ref, r32) for i in range(a, b) if r["contact_dist"][i] < 0.5]
        for i in range(a, b):
            if gd[i] > 0.5: continue
            if not any(np.abs(gp[i] - p).max() < 1e-4 and abs(gd[i] - dd) < 1e-4 for p, dd in cand):
                n_unexplained += 1
                print(f"world {e} kind {kind} contact {i}: gpu point {gp[i]} depth {gd[i]:.5f} not among the oracles' points", [(p.round(4), round(float(dd), 5)) for p, dd in cand])
print(f"{E} worlds: {n_neither} pairs equal to neither oracle, {n_normal} with another normal, {n_unexplained} gpu points not among the oracles' points")
