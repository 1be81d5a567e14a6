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


# ------------------------------------------------------------------------------------------------------------------------------
# sphere - convex, capsule - convex, convex - convex (second slice). The oracle restates
# mjx collision_convex.sphere_convex / capsule_convex / convex_convex [MEMORY, MJX 3.1.x; parity unpinned]; the CPU tests pin it
# against answers that follow from the geometry, the GPU test compares the engine's generic kernels with it.
PAIR = """<mujoco><option timestep="0.002"/><worldbody><geom name="base" type="box" size=".2 .2 .1" pos="0 0 .1"/>
  <body pos="0 0 1"><freejoint/><inertial pos="0 0 0" mass="1" diaginertia=".01 .01 .01"/>{geom}</body></worldbody></mujoco>"""


def _pair(tmp_path, geom):
    f = tmp_path / "pair.xml"
    f.write_text(PAIR.format(geom=geom))
    return load_mj_model_from_file(str(f))


def test_hull_topology_of_a_box_and_of_a_mesh(load_model):
    faces, normals, edges = mjcf.convex_topology(mjcf.box_vertices([0.1, 0.2, 0.05]))
    assert len(faces) == 6 and all(len(f) == 4 for f in faces) and len(edges) == 12
    v = mjcf.box_vertices([0.1, 0.2, 0.05])
    for f, n in zip(faces, normals):
        c = v[f].mean(axis=0)
        assert np.isclose(abs(n).max(), 1.0) and np.dot(c, n) > 0  # axis-aligned, outward
        for a, b in zip(f, f[1:] + f[:1]):  # counter-clockwise seen from outside
            assert np.dot(np.cross(v[a] - c, v[b] - c), n) > 0
    m = load_model("blocks")
    g = m.names["geom"].index("ramp")
    assert m.geom_vertnum[g] == 6 and m.geom_facenum[g] == 5 and m.geom_edgenum[g] == 9  # a triangular prism
    assert sorted(m.face_vertnum[m.geom_faceadr[g]:m.geom_faceadr[g] + 5].tolist()) == [3, 3, 4, 4, 4]
    assert m.pair_kind.tolist() == [mjcf.PAIR_CONVEX_CONVEX, mjcf.PAIR_CAPSULE_CONVEX, mjcf.PAIR_SPHERE_CONVEX] + [mjcf.PAIR_CONVEX_CONVEX] * 3
    # Euler's formula on every hull
    for g in range(m.ngeom):
        if m.geom_facenum[g]:
            assert m.geom_vertnum[g] - m.geom_edgenum[g] + m.geom_facenum[g] == 2


def test_sphere_against_a_box_face_edge_and_corner(tmp_path):
    m = _pair(tmp_path, '<geom type="sphere" size=".1"/>')
    assert m.pair_kind.tolist() == [mjcf.PAIR_SPHERE_CONVEX] and (m.pair_geom1[0], m.pair_geom2[0]) == (1, 0)
    o = Oracle(m)
    assert o.ncon == 1
    f = o.forward([0.05, 0.02, 0.2 + 0.1 - 0.01, 1, 0, 0, 0], np.zeros(6))  # over the top face, 1 cm deep
    assert np.isclose(f["contact_dist"][0], -0.01) and np.allclose(f["contact_pos"][0], [0.05, 0.02, 0.195])
    assert np.allclose(f["contact_frame"][0, 0], [0, 0, -1])  # from the sphere (geom 1 of the pair) to the box
    assert f["qfrc_constraint"][2] > 0 and np.allclose(f["qfrc_constraint"][[0, 1]], 0, atol=1e-9)
    f = o.forward([0.25, 0.0, 0.25, 1, 0, 0, 0], np.zeros(6))  # beyond the edge x = z = 0.2
    assert np.isclose(f["contact_dist"][0], np.hypot(0.05, 0.05) - 0.1) and np.allclose(f["contact_frame"][0, 0], [-np.sqrt(0.5), 0, -np.sqrt(0.5)], atol=1e-4)
    f = o.forward([0.25, 0.25, 0.25, 1, 0, 0, 0], np.zeros(6))  # beyond the corner
    assert np.isclose(f["contact_dist"][0], np.sqrt(3) * 0.05 - 0.1) and np.allclose(f["contact_frame"][0, 0], -np.ones(3) / np.sqrt(3), atol=1e-4)
    f = o.forward([0.0, 0.0, 0.5, 1, 0, 0, 0], np.zeros(6))  # clear
    # clear of the box: the top face is not "reached behind", so the literal algorithm measures against a side face's top edge
    assert np.isclose(f["contact_dist"][0], np.hypot(0.3, 0.2) - 0.1) and np.all(f["efc_force"] == 0)


def test_capsule_lying_standing_and_overhanging(tmp_path):
    m = _pair(tmp_path, '<geom type="capsule" size=".05 .1" euler="0 90 0"/>')
    assert m.pair_kind.tolist() == [mjcf.PAIR_CAPSULE_CONVEX]
    o = Oracle(m)
    assert o.ncon == 2
    f = o.forward([0.0, 0.0, 0.2 + 0.05 - 0.004, 1, 0, 0, 0], np.zeros(6))  # lying on the top face: one contact under each end
    assert np.allclose(f["contact_dist"], -0.004) and np.allclose(sorted(f["contact_pos"][:, 0]), [-0.1, 0.1]) and np.allclose(f["contact_pos"][:, 2], 0.198)
    assert np.allclose(f["contact_frame"][:, 0], [0, 0, -1])
    m = _pair(tmp_path, '<geom type="capsule" size=".05 .1"/>')
    f = Oracle(m).forward([0.05, 0.0, 0.2 + 0.15 - 0.004, 1, 0, 0, 0], np.zeros(6))  # standing: the lower end only
    assert np.isclose(f["contact_dist"][0], -0.004) and f["contact_dist"][1] > 0.19
    m = _pair(tmp_path, '<geom type="capsule" size=".05 .3" euler="0 90 0"/>')
    f = Oracle(m).forward([0.0, 0.0, 0.2 + 0.05 - 0.004, 1, 0, 0, 0], np.zeros(6))
    # longer than the box: the axis passes within a radius of the face's edges, and the literal algorithm then reports ONE edge
    # contact (the face contacts are dropped: the TODO in mjx's capsule_convex)
    assert np.isclose(f["contact_dist"][0], -0.004) and f["contact_dist"][1] == 1.0 and np.isclose(abs(f["contact_pos"][0, 0]), 0.2)


def test_box_on_box_face_contact_and_crossed_edges(tmp_path):
    m = _pair(tmp_path, '<geom type="box" size=".05 .08 .03"/>')
    assert m.pair_kind.tolist() == [mjcf.PAIR_CONVEX_CONVEX] and (m.pair_geom1[0], m.pair_geom2[0]) == (0, 1)
    o = Oracle(m)
    assert o.ncon == 4 and o.nefc == 16
    for yaw in (0.0, 0.3):
        f = o.forward([0.02, 0.01, 0.2 + 0.03 - 0.002, np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)], np.zeros(6))
        d = f["contact_dist"]
        act = d < 0
        assert act.sum() >= 3 and np.allclose(d[act], -0.002, atol=1e-9) and np.all(d[~act] == 1.0)
        assert np.allclose(f["contact_frame"][:, 0], [0, 0, 1])  # from the base (geom 1 of the pair) up to the small box
        c, s = np.cos(yaw), np.sin(yaw)
        corners = {(round(0.02 + c * x - s * y, 5), round(0.01 + s * x + c * y, 5)) for x in (-0.05, 0.05) for y in (-0.08, 0.08)}
        got = {(round(x, 5), round(y, 5)) for x, y in f["contact_pos"][act, :2]}
        assert got <= corners and len(got) == act.sum()  # the small box's bottom corners
        assert np.allclose(f["contact_pos"][act, 2], 0.199) and f["qfrc_constraint"][2] > 0
    # crossed edges: the small box's edge (local corner x = -.05, z = -.03, running along its y axis) is laid across the base's edge
    # x = z = 0.2 (which runs along y), perpendicular to it, its two faces symmetric about the diagonal n = (1, 0, 1) / sqrt 2
    n = np.array([1.0, 0.0, 1.0]) / np.sqrt(2)
    t = np.array([1.0, 0.0, -1.0]) / np.sqrt(2)
    yh = np.array([0.0, 1.0, 0.0])
    ax, az = (n - yh) / np.sqrt(2), (n + yh) / np.sqrt(2)
    if np.dot(np.cross(ax, t), az) < 0:
        ax, az = az, ax
    R = np.stack([ax, t, az], axis=1)  # columns: images of the box's x, y, z axes
    w = 0.5 * np.sqrt(max(1e-12, 1 + np.trace(R)))
    q = np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    pos = np.array([0.2, 0.0, 0.2]) - 0.003 * n - R @ np.array([-0.05, 0.0, -0.03])
    f = o.forward([*pos, *q], np.zeros(6))
    d = f["contact_dist"]
    assert np.isclose(d[0], -0.003, atol=1e-6) and np.all(d[1:] == 1.0)  # one edge - edge contact
    assert np.allclose(f["contact_frame"][0, 0], n, atol=1e-6)
    assert np.allclose(f["contact_pos"][0], np.array([0.2, 0, 0.2]) - 0.0015 * n, atol=1e-6)


def test_box_dropped_on_a_box_comes_to_rest_with_its_weight_carried(tmp_path):
    m = _pair(tmp_path, '<geom type="box" size=".05 .08 .03" friction="0.9 0.01 0.001"/>')
    o = Oracle(m, m.opt.replace(iterations=50, ls_iterations=50))
    q0 = np.array([0.03, -0.02, 0.2 + 0.03 + 0.01, np.cos(0.2), 0.02, 0.03, np.sin(0.2)])
    q0[3:] /= np.linalg.norm(q0[3:])
    xs = o.rollout(np.concatenate([q0, np.zeros(6)]), np.zeros((1200, 0)))
    q, v = xs[-1, :7], xs[-1, 7:]
    assert np.abs(v).max() < 0.02 and abs(q[2] - 0.23) < 2e-3  # resting on the base's top face
    f = o.forward(q, v)
    assert np.sum(f["contact_dist"] < 0) >= 3
    assert np.allclose(f["qfrc_constraint"][:3], [0, 0, 9.81], atol=0.05 * 9.81)


def test_oversized_or_flat_hulls_are_refused_like_mjx(tmp_path):
    flat = '<geom type="mesh" mesh="flat"/>'
    xml = PAIR.format(geom=flat).replace("<worldbody>", '<asset><mesh name="flat" vertex="0 0 0  .1 0 0  .1 .1 0  0 .1 0"/></asset><worldbody>')
    f = tmp_path / "flat.xml"
    f.write_text(xml)
    m = load_mj_model_from_file(str(f))
    assert m.n_unsupported_pairs == 1 and "hull" in m.unsupported_reason
    from ambersim_b200 import mjx
    assert mjx.device_put(m).n_unsupported_pairs == 1  # the device handle refuses it (NotImplementedError) unless contacts are disabled


def _compare_contacts_with_the_oracles(mj, m, o, qs, vs, cs):
    """mjx.forward on a batch of states: every pair in contact must equal the float64 oracle (5e-6) or, where _manifold_points breaks a
    structural tie by rounding, the float32 oracle (MJX's own precision, 2e-5). Returns (pairs in contact per kind, ties, neither)."""
    import torch

    from ambersim_b200 import mjx

    E = len(qs)
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    names = ("contact_dist", "contact_pos", "contact_frame", "efc_force", "qfrc_constraint", "qacc_smooth")
    d = mjx.Data(qpos=t(qs), qvel=t(vs), ctrl=t(cs), qacc=torch.zeros((E, mj.nv), device="cuda"), qacc_warmstart=torch.zeros((E, mj.nv), device="cuda"),
                 time=torch.zeros(E, device="cuda"))
    f = mjx.forward(m, d, fields=names)
    spans = []
    c0 = 0
    for k in mj.pair_kind:
        spans.append((c0, c0 + mjcf.PAIR_NCON[int(k)], int(k)))
        c0 += mjcf.PAIR_NCON[int(k)]
    seen = {mjcf.PAIR_SPHERE_CONVEX: 0, mjcf.PAIR_CAPSULE_CONVEX: 0, mjcf.PAIR_CONVEX_CONVEX: 0}
    flips, checked, ties = [], 0, 0

    def close(r, a, b, gd, gp, gf, tol):
        return (np.abs(r["contact_dist"][a:b] - gd[a:b]).max() < tol and np.abs(r["contact_pos"][a:b] - gp[a:b]).max() < tol
                and np.abs(r["contact_frame"][a:b] - gf[a:b]).max() < 10 * tol)

    for e in range(E):
        ref = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv))
        gd, gp, gf = f.contact_dist[e].cpu().numpy(), f.contact_pos[e].cpu().numpy(), f.contact_frame[e].cpu().numpy()
        r32 = None
        world_ok = True
        for a, b, kind in spans:
            rd = ref["contact_dist"][a:b]
            if not (rd < 0).any() and not (gd[a:b] < 0).any():
                continue  # separated in both: which (inactive) candidates are reported is precision dependent
            checked += 1
            if close(ref, a, b, gd, gp, gf, 5e-6):
                seen[kind] += 1
                continue
            # _manifold_points breaks structural ties by rounding (the two off-diagonal corners of a rectangular contact patch are
            # equally far from the diagonal), so the float64 oracle may keep another corner than float32 arithmetic does: such pairs
            # must then equal the FLOAT32 oracle (MJX's own precision)
            world_ok = False
            if r32 is None:
                r32 = o.forward(f.qpos[e].cpu().numpy(), vs[e], cs[e], np.zeros(mj.nv), prec=1)
            if close(r32, a, b, gd, gp, gf, 2e-5):
                ties += 1
                seen[kind] += 1
            else:
                flips.append((e, kind, np.round(rd, 4).tolist(), np.round(r32["contact_dist"][a:b], 4).tolist(), np.round(gd[a:b], 4).tolist()))
        if world_ok:
            for n in ("efc_force", "qfrc_constraint", "qacc_smooth"):
                r, g = ref[n].ravel(), getattr(f, n)[e].cpu().numpy().ravel()
                assert np.abs(r - g).max() <= 5e-4 * max(1e-6, np.abs(r).max()) + 1e-5, (n, e)
            assert np.abs(ref["qacc"] - f.qacc[e].cpu().numpy()).max() <= 1e-3 * max(1.0, np.abs(ref["qacc"]).max())
    print(f"convex pairs in contact: {checked} checked, {ties} equal to the float32 oracle only (rounding-decided ties), {len(flips)} equal to neither", flips)
    return seen, checked, ties, flips


@pytest.mark.gpu
def test_engine_matches_oracle_on_sphere_capsule_and_convex_pairs(load_model):
    import torch

    from ambersim_b200 import mjx

    mj = load_model("blocks")
    m = mjx.device_put(mj)
    assert "generic kernels" in m.describe()
    o = Oracle(mj)
    rng = np.random.default_rng(5)
    E = int(__import__("os").environ.get("ABR_SOAK_E", "96"))  # a soak run sets ABR_SOAK_E higher
    qs = np.tile(mj.key_qpos("home"), (E, 1))
    qs[:, 0] = rng.uniform(-0.2, 0.45, E)
    qs[:, 1] = rng.uniform(-0.2, 0.2, E)
    qs[:, 2] = rng.uniform(0.3, 0.42, E)
    quat = np.array([1, 0, 0, 0]) + 0.2 * rng.normal(size=(E, 4))
    qs[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    qs[:, 7:] += rng.uniform(-0.6, 0.6, (E, 3))
    vs = 0.3 * rng.normal(size=(E, mj.nv))
    cs = mj.key_ctrl("home") + 0.2 * rng.normal(size=(E, mj.nu))
    seen, checked, ties, flips = _compare_contacts_with_the_oracles(mj, m, o, qs, vs, cs)
    assert all(v >= 5 for v in seen.values()), seen  # every new pair function is exercised in contact
    # what equals neither oracle are ties again (the device contracts multiply-adds, the float32 oracle on the CPU does not): a few per cent
    assert len(flips) <= 0.05 * checked, (len(flips), checked, flips)
    # teacher-forced steps: each device step from the oracle's own state, while the robot settles on the pedestal
    from ambersim_b200.trajopt.shooting import shoot

    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    x = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    x[2] = 0.36
    for k in range(60):
        u = mj.key_ctrl("home") + 0.3 * rng.normal(size=mj.nu)
        nxt = o.rollout(x, u[None])[1]
        got = shoot(m, t(x), t(u[None])).cpu().numpy()[1]
        assert np.abs(got - nxt).max() < 1e-3 * max(1.0, np.abs(nxt).max()), k
        x = nxt


@pytest.mark.gpu
def test_bh280_with_its_real_small_hulls_and_contacts_on(load_model):
    """The reference's Barrett hand with contacts ON, on the tractable part of its real collision geometry (`tests/models/bh280_hulls.xml`:
    the 7 collision hulls of at most 48 vertices, 19 hull - hull pairs, 76 contact slots, from the reference's OBJ files): finger tips
    closing on each other. contact_dist / pos / frame and the resulting accelerations against the oracle, then steps."""
    import torch

    from ambersim_b200 import mjx
    from ambersim_b200.trajopt.shooting import shoot

    mj = load_model("bh280_hulls")
    assert mj.npair == 19 and mj.n_unsupported_pairs == 0 and set(mj.pair_kind.tolist()) == {mjcf.PAIR_CONVEX_CONVEX}
    m = mjx.device_put(mj)
    m = m.replace(opt=m.opt.replace(timestep=0.002, iterations=2, ls_iterations=6))
    o = Oracle(mj, m.opt)
    assert (o.ncon, o.ne, o.nl, o.nefc) == (76, 4, 8, 316) and "generic kernels" in m.describe()
    rng = np.random.default_rng(7)
    E = 48
    qs = np.zeros((E, 8))
    qs[:, 0] = rng.uniform(2.2, 2.44, E); qs[:, 1] = rng.uniform(0.6, 0.84, E)      # finger 3 closed
    qs[:, 2] = rng.uniform(1.1, 1.9, E); qs[:, 3] = rng.uniform(2.25, 2.43, E); qs[:, 4] = rng.uniform(0.6, 0.84, E)  # finger 1 swung in and closed
    qs[:, 5] = qs[:, 2] + rng.uniform(-0.05, 0.05, E); qs[:, 6] = rng.uniform(1.0, 2.4, E); qs[:, 7] = rng.uniform(0.2, 0.8, E)
    vs = 0.5 * rng.normal(size=(E, 8))
    cs = rng.uniform(-1, 1, (E, mj.nu))
    seen, checked, ties, flips = _compare_contacts_with_the_oracles(mj, m, o, qs, vs, cs)
    assert seen[mjcf.PAIR_CONVEX_CONVEX] >= 8 and len(flips) <= max(1, 0.1 * checked), (seen, checked, flips)
    # teacher-forced steps from a touching pose
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    x = np.concatenate([[2.34, 0.7, 1.5, 2.35, 0.7, 1.5, 1.4, 0.5], np.zeros(8)])
    for k in range(25):
        u = rng.uniform(-1, 1, mj.nu)
        nxt = o.rollout(x, u[None])[1]
        got = shoot(m, t(x), t(u[None])).cpu().numpy()[1]
        assert np.abs(got - nxt).max() < 2e-3 * max(1.0, np.abs(nxt).max()), k
        x = nxt


@pytest.mark.gpu
def test_c_abi_refuses_inconsistent_hull_tables(load_model):
    """The collision functions index fixed-size per-lane arrays with the hull tables: abr_model_create checks them once."""
    import copy
    import ctypes as C

    from ambersim_b200 import _abi, _lib

    mj = load_model("blocks")
    for field, value, what in (("face_vertnum", 9, "corners"), ("face_vert", 99, "vertex set"),
                               ("edge_vert", -1, "vertex set"), ("pair_kind", 12, "pair")):
        bad = copy.copy(mj)
        arr = np.array(getattr(mj, field)).copy()
        arr.ravel()[0] = value
        setattr(bad, field, arr)
        host, keep = _abi.pack_model(bad, bad.opt)
        ptr = C.c_void_p()
        rc = _lib.lib().abr_model_create(C.byref(host), 0, C.byref(ptr))
        assert rc == _lib.ABR_EINVAL and what in _lib.lib().abr_last_error().decode(), (field, rc, _lib.lib().abr_last_error())
        assert not ptr.value


def test_hull_topology_of_random_point_clouds():
    """convex_topology on random hulls: closed polyhedra (Euler), outward unit normals, every polygon planar, convex and counter-clockwise,
    every edge shared by exactly two faces."""
    rng = np.random.default_rng(3)
    for trial in range(12):
        n = int(rng.integers(5, 40))
        pts = rng.normal(size=(n, 3)) * rng.uniform(0.05, 0.3, 3)
        if trial % 3 == 0:  # some coplanar groups: a prism-like cloud, so that triangles get merged into polygons
            pts = np.concatenate([np.c_[rng.normal(size=(6, 2)) * 0.1, np.full(6, -0.1)], np.c_[rng.normal(size=(6, 2)) * 0.1, np.full(6, 0.1)]])
        v = mjcf.convex_vertices(pts)
        faces, normals, edges = mjcf.convex_topology(v)
        assert len(v) - len(edges) + len(faces) == 2
        c = v.mean(axis=0)
        count = {}
        for f, nn in zip(faces, normals):
            assert 3 <= len(f) <= mjcf.MAX_FACE_VERTS and np.isclose(np.linalg.norm(nn), 1.0)
            P = v[f]
            assert np.abs((P - P[0]) @ nn).max() < 1e-9  # planar
            assert np.dot(P.mean(axis=0) - c, nn) > 0  # outward
            assert np.all((v - P[0]) @ nn < 1e-9)  # a supporting plane of the hull
            for a, b, d in zip(f, f[1:] + f[:1], f[2:] + f[:2]):
                assert np.dot(np.cross(v[b] - v[a], v[d] - v[b]), nn) > -1e-12  # convex, counter-clockwise seen from outside
                count[(min(a, b), max(a, b))] = count.get((min(a, b), max(a, b)), 0) + 1
        assert sorted(count) == [tuple(e) for e in edges.tolist()] and set(count.values()) == {2}


@pytest.mark.gpu
def test_model_without_actuators_steps_and_shoots(tmp_path):
    """nu = 0 (a passive box dropped on a box): mjx.step and shoot take (.., 0)-shaped controls; the rollout matches the oracle."""
    import torch

    from ambersim_b200 import mjx
    from ambersim_b200.trajopt.shooting import shoot

    mj = _pair(tmp_path, '<geom type="box" size=".05 .08 .03" friction="0.9 0.01 0.001"/>')
    assert mj.nu == 0
    m = mjx.device_put(mj)
    o = Oracle(mj)
    q0 = np.array([0.03, -0.02, 0.2 + 0.03 + 0.004, np.cos(0.2), 0.02, 0.03, np.sin(0.2)])
    q0[3:] /= np.linalg.norm(q0[3:])
    x0 = np.concatenate([q0, np.zeros(6)])
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    d = mjx.step(m, mjx.Data(qpos=t(np.tile(q0, (3, 1))), qvel=torch.zeros(3, 6, device="cuda"), ctrl=torch.zeros(3, 0, device="cuda"),
                             qacc=torch.zeros(3, 6, device="cuda"), qacc_warmstart=torch.zeros(3, 6, device="cuda"), time=torch.zeros(3, device="cuda")))
    ref = o.rollout(x0, np.zeros((1, 0)))[1]
    assert d.ctrl.shape == (3, 0) and np.abs(d.qpos[1].cpu().numpy() - ref[:7]).max() < 1e-5
    xs = shoot(m, t(x0), torch.zeros(20, 0, device="cuda")).cpu().numpy()
    refx = o.rollout(x0, np.zeros((20, 0)))
    assert xs.shape == (21, 13) and np.abs(xs[:8] - refx[:8]).max() < 5e-4
