"""Plane - convex collision (boxes and convex meshes as vertex sets, up to 4 contacts per pair): SURVEY 8(f) rank 2, first slice.
The oracle restates mjx collision_convex.plane_convex / _manifold_points [MEMORY, MJX 3.1.x; parity unpinned like the rest];
these CPU tests pin it against answers that follow from the definition, the GPU tests compare the engine with it."""
import numpy as np
import pytest

from ambersim_b200.utils import mjcf
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from oracle.oracle import Oracle

BOX = """<mujoco><option timestep="0.002"/><worldbody><geom type="plane" size="1 1 .1" euler="{tilt}"/>
  <body pos="0 0 1"><freejoint/><inertial pos="0 0 0" mass="2" diaginertia=".02 .03 .04"/>
  <geom type="box" size=".1 .2 .05" friction="0.9 0.01 0.001"/></body></worldbody></mujoco>"""


def _box(tmp_path, tilt="0 0 0"):
    f = tmp_path / "box.xml"
    f.write_text(BOX.format(tilt=tilt))
    return load_mj_model_from_file(str(f))


def test_flat_box_gives_its_four_bottom_corners(tmp_path):
    m = _box(tmp_path)
    assert m.pair_kind.tolist() == [mjcf.PAIR_PLANE_CONVEX] and m.nvert == 8
    o = Oracle(m)
    assert o.ncon == 4 and o.nefc == 16
    depth = 0.003
    f = o.forward([0.3, -0.2, 0.05 - depth, 1, 0, 0, 0], np.zeros(6))
    # _manifold_points picks a = first masked vertex, b = furthest from a (the diagonal), c = furthest from the line ab, d = furthest
    # from the edges bc / ac: on an exactly rectangular face the last choice TIES between a and the fourth corner and jnp.argmax takes
    # the first, so the literal algorithm yields three corners and one duplicate (reported inactive, dist = 1)
    d = f["contact_dist"]
    act = d < 0
    assert np.allclose(d[act], -depth, atol=1e-7) and act.sum() >= 3 and np.all(d[~act] == 1.0)
    corners = {(round(x, 6), round(y, 6)) for x, y in f["contact_pos"][act, :2]}
    assert corners <= {(0.2, -0.4), (0.4, -0.4), (0.2, 0.0), (0.4, 0.0)} and len(corners) == act.sum()
    assert np.allclose(f["contact_pos"][act, 2], -depth / 2, atol=1e-7)  # midway between the vertex and the plane
    assert np.allclose(f["contact_frame"][:, 0], [0, 0, 1]) and np.allclose(np.linalg.det(f["contact_frame"]), 1.0)
    assert np.isclose(f["efc_force"].sum() > 0, True) and np.all(f["efc_force"] >= 0)


def test_box_on_an_edge_and_on_a_corner_and_clear_of_the_plane(tmp_path):
    m = _box(tmp_path)
    o = Oracle(m)
    # rotated 30 degrees about x: the two -y bottom corners are lowest, well separated from the other two (> 1 mm skin)
    ang = np.deg2rad(30)
    quat = [np.cos(ang / 2), np.sin(ang / 2), 0, 0]
    R = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    low = min((R @ v)[2] for v in m.vert)
    f = o.forward([0, 0, -low - 0.002, *quat], np.zeros(6))
    d = f["contact_dist"]
    assert np.sum(np.isclose(d, -0.002, atol=1e-6)) == 2 and np.sum(d == 1.0) == 2  # a vertex picked twice is an inactive contact
    # on one corner: a single active contact
    quat = np.array([0.9, 0.25, 0.3, 0.1]); quat /= np.linalg.norm(quat)
    from oracle.independent import _qmat
    R = _qmat(quat)
    low = min((R @ v)[2] for v in m.vert)
    f = o.forward([0, 0, -low - 0.001, *quat], np.zeros(6))
    assert np.sum(f["contact_dist"] < 0) == 1 and np.isclose(f["contact_dist"].min(), -0.001, atol=1e-6)
    # clear of the plane: nothing active, distances are vertex heights
    f = o.forward([0, 0, 0.5, 1, 0, 0, 0], np.zeros(6))
    assert np.all(f["contact_dist"] > 0) and np.isclose(f["contact_dist"].min(), 0.45) and np.all(f["efc_force"] == 0)


def test_mesh_geom_is_its_convex_hull(load_model):
    m = load_model("boxbot")
    g = m.names["geom"].index("foot")
    assert m.geom_type[g] == mjcf.GEOM_MESH and m.geom_vertnum[g] == 8  # 9 vertices in the file, one inside the hull
    v = m.vert[m.geom_vertadr[g]:m.geom_vertadr[g] + 8]
    assert not any(np.allclose(p, [0, 0, 0.02]) for p in v)
    assert m.geom_vertnum.tolist() == [0, 8, 8, 0] and m.pair_kind.tolist() == [5, 5, 0]


def test_box_settles_on_a_tilted_plane_and_the_contacts_carry_its_weight(tmp_path):
    m = _box(tmp_path, tilt="0.05 -0.04 0")
    o = Oracle(m, m.opt.replace(iterations=50, ls_iterations=50))
    x0 = np.concatenate([[0, 0, 0.06, 1, 0, 0, 0], np.zeros(6)])
    xs = o.rollout(x0, np.zeros((1500, 0)))
    q, v = xs[-1, :7], xs[-1, 7:]
    assert np.abs(v).max() < 0.05  # friction holds it on the 3.7 degree slope
    f = o.forward(q, v)
    assert np.sum(f["contact_dist"] < 0) >= 3
    # net contact force balances gravity
    assert np.allclose(f["qfrc_constraint"][:3], [0, 0, 2 * 9.81], atol=0.05 * 2 * 9.81)


@pytest.mark.gpu
def test_engine_matches_oracle_on_convex_contacts(load_model):
    import torch

    from ambersim_b200 import mjx

    mj = load_model("boxbot")
    m = mjx.device_put(mj)
    assert "generic kernels" in m.describe()
    o = Oracle(mj)
    rng = np.random.default_rng(3)
    E = 32
    qs = np.tile(mj.key_qpos("home"), (E, 1))
    qs[:, 2] = rng.uniform(0.02, 0.3, E)
    quat = np.array([1, 0, 0, 0]) + 0.25 * rng.normal(size=(E, 4))
    qs[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    qs[:, 7:] += rng.uniform(-0.4, 0.4, (E, 2))
    vs = 0.3 * rng.normal(size=(E, mj.nv))
    cs = mj.key_ctrl("home") + 0.2 * rng.normal(size=(E, mj.nu))
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    names = ("contact_dist", "contact_pos", "contact_frame", "efc_D", "efc_aref", "efc_force", "qfrc_constraint", "qacc_smooth")
    d = mjx.Data(qpos=t(qs), qvel=t(vs), ctrl=t(cs), qacc=torch.zeros((E, mj.nv), device="cuda"), qacc_warmstart=torch.zeros((E, mj.nv), device="cuda"),
                 time=torch.zeros(E, device="cuda"))
    f = mjx.forward(m, d, fields=names)
    active = 0
    for e in range(E):
        ref = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv))
        active += int((ref["contact_dist"] < 0).sum())
        gd, gp = f.contact_dist[e].cpu().numpy(), f.contact_pos[e].cpu().numpy()
        for c0, c1 in ((0, 4), (4, 8), (8, 9)):  # box pair, mesh pair, sphere
            if (ref["contact_dist"][c0:c1] < 0).any():
                assert np.abs(ref["contact_dist"][c0:c1] - gd[c0:c1]).max() < 2e-6
                assert np.abs(ref["contact_pos"][c0:c1] - gp[c0:c1]).max() < 2e-6
            else:
                # nothing penetrates: every vertex carries the mask offset -1e6, which swallows the float32 scores (ulp 0.06), so the
                # manifold picked among the SEPARATED vertices depends on the precision (in MJX's float32 too); the contacts are inactive
                assert np.all(gd[c0:c1] > 0)
        assert np.abs(ref["contact_frame"] - f.contact_frame[e].cpu().numpy()).max() < 2e-6
        for n in ("efc_D", "efc_aref", "efc_force", "qfrc_constraint", "qacc_smooth"):
            r, g = ref[n].ravel(), getattr(f, n)[e].cpu().numpy().ravel()
            assert np.abs(r - g).max() <= 5e-4 * max(1e-6, np.abs(r).max()) + 1e-5, (n, e)
        assert np.abs(ref["qacc"] - f.qacc[e].cpu().numpy()).max() <= 1e-3 * max(1.0, np.abs(ref["qacc"]).max())
    assert active >= 20  # the sample really exercises box, mesh and sphere contacts
    # teacher-forced steps: each device step from the oracle's own state
    from ambersim_b200.trajopt.shooting import shoot

    x = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    x[2] = 0.12
    for k in range(40):
        u = mj.key_ctrl("home") + 0.3 * rng.normal(size=mj.nu)
        nxt = o.rollout(x, u[None])[1]
        got = shoot(m, t(x), t(u[None])).cpu().numpy()[1]
        assert np.abs(got - nxt).max() < 5e-4 * max(1.0, np.abs(nxt).max()), k
        x = nxt
