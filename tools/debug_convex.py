"""Engine vs oracle on the convex-pair fixture: prints every pair whose contacts differ (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from ambersim_b200 import mjx
from ambersim_b200.utils import mjcf
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from oracle.oracle import Oracle
np.set_printoptions(precision=5, suppress=True, linewidth=200)
mj = load_mj_model_from_file("tests/models/blocks.xml"); m = mjx.device_put(mj); o = Oracle(mj)
rng = np.random.default_rng(5); E = 96
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
for e in range(E):
    ref = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv))
    r32 = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv), prec=1)
    gd, gp, gf = f.contact_dist[e].cpu().numpy(), f.contact_pos[e].cpu().numpy(), f.contact_frame[e].cpu().numpy()
    for a, b, kind in spans:
        rd = ref["contact_dist"][a:b]
        if not (rd < 0).any() and not (gd[a:b] < 0).any(): continue
        if np.abs(rd - gd[a:b]).max() < 2e-5 and np.abs(ref["contact_pos"][a:b] - gp[a:b]).max() < 2e-5 and np.abs(ref["contact_frame"][a:b] - gf[a:b]).max() < 2e-4: continue
        print(f"--- world {e} pair kind {kind} contacts {a}:{b}")
        print(" oracle f64 dist", rd, " f32 oracle", r32["contact_dist"][a:b], " gpu", gd[a:b])
        print(" oracle pos\n", ref["contact_pos"][a:b], "\n f32 oracle pos\n", r32["contact_pos"][a:b], "\n gpu pos\n", gp[a:b])
        print(" normals", ref["contact_frame"][a, 0], gf[a, 0])
