"""Algorithm-independent cross-checks: the oracle and the engine share the loader's flattened
model and MJX's algorithms, so a common-mode mistake passes every oracle-vs-engine parity test. These tests restate
the same physics along different routes (oracle/independent.py) and compare:

* the loader's mj_setConst constants vs numerically differentiated kinematics + a kinetic-energy mass matrix;
* the oracle's CRBA + RNE + factorisation `qacc_smooth` vs Featherstone's articulated-body algorithm;
* the oracle's converged Newton solve vs a generic scipy minimiser of the primal constraint cost, in contact.

Parity with MJX itself stays UNPINNED (no MJX here); these pin the oracle against physics it must obey."""
import numpy as np
import pytest

from oracle import independent as ind
from oracle.oracle import Oracle

ALL = ["pendulum", "bh280", "barkour", "biped", "exolegs", "tripod", "tripod3", "fixedbase"]
KEYS = {"barkour": "home", "biped": "stand", "exolegs": "stand", "tripod": "home", "tripod3": "home"}


def _random_state(m, name, rng, spread=0.3):
    q = m.key_qpos(KEYS[name]) if name in KEYS else m.qpos0.copy()
    if m.jnt_type[0] == 0:
        q[7:] += rng.uniform(-spread, spread, m.nq - 7)
        quat = rng.normal(size=4)
        q[3:7] = quat / np.linalg.norm(quat)
    else:
        q = q + rng.uniform(0.0, 0.5, m.nq)
    return q, rng.normal(size=m.nv)


@pytest.mark.parametrize("name", ALL)
def test_setconst_against_numeric_route(load_model, name):
    m = load_model(name)
    sc = ind.setconst_numeric(m)
    assert np.allclose(sc["qM0"], m.qM0, rtol=0, atol=1e-8 * np.abs(m.qM0).max())
    assert np.allclose(sc["dof_invweight0"], m.dof_invweight0, rtol=1e-7)
    assert np.allclose(sc["body_invweight0"], m.body_invweight0, rtol=1e-7, atol=1e-12)
    assert np.isclose(sc["meaninertia"], m.stat.meaninertia, rtol=1e-8)
    assert np.array_equal(sc["body_subtreemass"], m.body_subtreemass)


def test_setconst_pendulum_closed_form(load_model):
    """One hinge: dof_invweight0 = 1 / I_axis, body_invweight0 = (r^2 / I / 3 summed over the two in-plane axes, 1 / I / 3)."""
    m = load_model("pendulum")
    b = int(m.dof_bodyid[0])
    I = float(m.qM0[0, 0])
    assert np.isclose(m.dof_invweight0[0], 1.0 / I, rtol=1e-12)
    assert np.isclose(m.stat.meaninertia, I, rtol=1e-12)
    xpos, xrot, xipos, _ = ind.kinematics(m, m.qpos0)
    j = int(m.body_jntadr[b])
    r = xipos[b] - (xpos[b] + xrot[b] @ m.jnt_pos[j])
    axis = xrot[b] @ m.jnt_axis[j]
    lever = np.cross(axis, r)
    assert np.isclose(m.body_invweight0[b, 0], lever @ lever / I / 3, rtol=1e-10)
    assert np.isclose(m.body_invweight0[b, 1], 1.0 / I / 3, rtol=1e-10)


@pytest.mark.parametrize("name", ALL)
def test_forward_dynamics_against_articulated_body_algorithm(load_model, name):
    m = load_model(name)
    rng = np.random.default_rng(11)
    o = Oracle(m, m.opt.replace(disableflags=1))  # constraints off: qacc is qacc_smooth
    for _ in range(3):
        q, v = _random_state(m, name, rng)
        c = rng.normal(size=m.nu) * 0.3
        f = o.forward(q, v, c)
        tau = f["qfrc_passive"] + f["qfrc_actuator"]
        qa = ind.aba(m, q, v, tau, m.opt.gravity)
        # model parameters reach the oracle as float32 (the ABI's precision): 1e-6 relative
        assert np.abs(qa - f["qacc_smooth"]).max() < 2e-6 * max(1.0, np.abs(f["qacc_smooth"]).max())
        M = ind.mass_matrix_energy(m, q)
        assert np.abs(M - f["qM"]).max() < 1e-6 * np.abs(f["qM"]).max()


@pytest.mark.parametrize("name", ["barkour", "biped", "exolegs", "tripod", "bh280", "fixedbase"])
def test_constraint_solve_against_generic_minimiser(load_model, name):
    m = load_model(name)
    rng = np.random.default_rng(13)
    base = m.opt.replace(disableflags=16) if name in ("bh280", "fixedbase") else m.opt
    conv = Oracle(m, base.replace(solver=2, iterations=200, ls_iterations=50, tolerance=1e-14))
    for trial in range(3):
        if name in ("bh280", "fixedbase"):
            q, v = rng.uniform(-0.3, 0.8, m.nq), rng.normal(size=m.nv)  # past the lower joint limits, equalities active
        else:
            q, v = _random_state(m, name, rng, spread=0.15)
            q[3:7] = m.key_qpos(KEYS[name])[3:7]
            q[2] -= 0.01 + 0.01 * trial  # feet pressed into the floor
        c = rng.normal(size=m.nu) * 0.2
        f = conv.forward(q, v, c)
        assert f["efc_D"].size and (f["efc_D"] > 0).any()
        a = ind.constraint_minimum(f["qM"], f["qacc_smooth"], f["efc_J"], f["efc_D"], f["efc_aref"], conv.ne)
        scale = max(1.0, np.abs(f["qacc"]).max())
        assert np.abs(a - f["qacc"]).max() < 1e-6 * scale
        # and the single-iteration configuration the benchmarks run lowers the same cost from the same start
        one = Oracle(m, base).forward(q, v, c, qacc_warmstart=f["qacc_smooth"])

        def cost(x):
            r = f["efc_J"] @ x - f["efc_aref"]
            act = (np.arange(r.size) < conv.ne) | (r < 0)
            e = x - f["qacc_smooth"]
            return 0.5 * e @ f["qM"] @ e + 0.5 * np.sum(f["efc_D"] * r * r * act)

        assert cost(a) <= cost(one["qacc"]) + 1e-9 * abs(cost(a)) <= cost(f["qacc_smooth"]) + 1e-9 * abs(cost(a))


@pytest.mark.parametrize("seed", range(8))
def test_random_models_setconst_and_aba(tmp_path, seed):
    """The same three routes on random trees (tilted inertial frames, slides, forks, random anchors and axes)."""
    from ambersim_b200.utils.io_utils import load_mj_model_from_file
    from tests._randmodel import random_limb_model

    xml, home, _ = random_limb_model(100 + seed, max_chain=3 + (seed % 2) * 3, max_con=1 + (seed % 2) * 3, max_leaves=4)
    p = tmp_path / f"rand{seed}.xml"
    p.write_text(xml)
    m = load_mj_model_from_file(str(p))
    sc = ind.setconst_numeric(m)
    assert np.allclose(sc["dof_invweight0"], m.dof_invweight0, rtol=1e-6)
    assert np.allclose(sc["body_invweight0"], m.body_invweight0, rtol=1e-6, atol=1e-12)
    assert np.isclose(sc["meaninertia"], m.stat.meaninertia, rtol=1e-8)
    rng = np.random.default_rng(seed)
    q = np.array(home, dtype=np.float64)
    q[7:] += rng.uniform(-0.05, 0.05, m.nq - 7)
    quat = rng.normal(size=4)
    q[3:7] = quat / np.linalg.norm(quat)
    v = rng.normal(size=m.nv)
    f = Oracle(m, m.opt.replace(disableflags=1)).forward(q, v, rng.normal(size=m.nu) * 0.1)
    qa = ind.aba(m, q, v, f["qfrc_passive"] + f["qfrc_actuator"], m.opt.gravity)
    assert np.abs(qa - f["qacc_smooth"]).max() < 5e-6 * max(1.0, np.abs(f["qacc_smooth"]).max())


# ---------------------------------------------------------------------------------------------------------------------------------
# Convex collision: the oracle's separating-axis / clipping restatement against routes that share nothing with it
# (Minkowski-difference hull for the penetration of two polytopes, constrained least squares for point / segment distances).
CONVEX_PAIR = """<mujoco><worldbody><geom name="base" type="{base}" {basearg} pos="0 0 .1"/>
  <body pos="0 0 1"><freejoint/><inertial pos="0 0 0" mass="1" diaginertia=".01 .01 .01"/>{geom}</body></worldbody>
  <asset><mesh name="gem" vertex="0 0 .07  .06 0 0  -.03 .05 0  -.03 -.05 0  0 0 -.05  .04 .04 .03"/>
  <mesh name="slab" vertex="-.2 -.2 -.1  .2 -.2 -.1  .2 .2 -.1  -.2 .2 -.1  -.15 -.2 .1  .2 -.2 .1  .2 .15 .1  -.15 .15 .1"/></asset></mujoco>"""


def _convex_pair(tmp_path, geom, base="box"):
    from ambersim_b200.utils.io_utils import load_mj_model_from_file

    f = tmp_path / "cp.xml"
    f.write_text(CONVEX_PAIR.format(geom=geom, base=base, basearg='size=".2 .2 .1"' if base == "box" else 'mesh="slab"'))
    return load_mj_model_from_file(str(f))


@pytest.mark.parametrize("base", ["box", "mesh"])
@pytest.mark.parametrize("geom", ['<geom type="box" size=".05 .08 .03"/>', '<geom type="mesh" mesh="gem"/>'])
def test_convex_convex_against_the_minkowski_difference(tmp_path, geom, base):
    m = _convex_pair(tmp_path, geom, base)
    o = Oracle(m)
    rng = np.random.default_rng(11)
    g1, g2 = int(m.pair_geom1[0]), int(m.pair_geom2[0])
    touching = edge = exact = preferred = 0
    for _ in range(120):
        q = np.concatenate([rng.uniform(-0.22, 0.22, 2), [rng.uniform(0.2, 0.3)], rng.normal(size=4)])
        q[3:] /= np.linalg.norm(q[3:])
        xpos, xrot, _, _ = ind.kinematics(m, q)
        depth, normal = ind.polytope_penetration(ind.geom_world_vertices(m, g1, xpos, xrot), ind.geom_world_vertices(m, g2, xpos, xrot))
        f = o.forward(q, np.zeros(6))
        d, n = f["contact_dist"], f["contact_frame"][0, 0]
        if depth < 1e-4:
            assert not (d < -1e-4).any()  # separated (or grazing): no contact reports a penetration
            continue
        touching += 1
        # the contact normal is the direction of least penetration (or a FACE normal within the 1e-5 m by which the restatement
        # prefers face axes to edge - edge axes), and no contact point is deeper than the penetration along it
        VA, VB = ind.geom_world_vertices(m, g1, xpos, xrot), ind.geom_world_vertices(m, g2, xpos, xrot)
        along = float((VA @ n).max() - (VB @ n).min())
        if not np.allclose(n, normal, atol=2e-6):
            preferred += 1
            assert depth - 2e-7 <= along <= depth + 1.1e-5, (q, n, normal, along, depth)
        depth = along
        assert d.min() >= -depth - 2e-7  # (the model's vertices cross the ABI as float32)
        if (d == 1.0).sum() == 3:  # an edge - edge contact: its single point carries the whole penetration
            edge += 1
            assert np.isclose(d[0], -depth, atol=2e-7)
        exact += int(np.isclose(d.min(), -depth, atol=2e-7))
    assert touching >= 40 and edge >= 3 and preferred <= 0.1 * touching
    # face contacts: the deepest clipped point reaches the full depth unless the deepest vertex lies outside the reference face's prism
    assert exact >= 0.8 * touching, (exact, touching)


def test_sphere_convex_against_point_polytope_distance(tmp_path):
    for base in ("box", "mesh"):
        m = _convex_pair(tmp_path, '<geom type="sphere" size=".06"/>', base)
        o = Oracle(m)
        rng = np.random.default_rng(12)
        g2 = int(m.pair_geom2[0])
        active = 0
        for _ in range(60):
            q = np.concatenate([rng.uniform(-0.27, 0.27, 2), [rng.uniform(0.2, 0.27)], [1, 0, 0, 0]])
            xpos, xrot, _, _ = ind.kinematics(m, q)
            V = ind.geom_world_vertices(m, g2, xpos, xrot)
            dist, closest = ind.point_polytope_distance(q[:3], V)
            f = o.forward(q, np.zeros(6))
            if dist < 1e-6:
                continue  # centre inside the hull: outside what the reference's function is meant for
            if dist - 0.06 < 0:
                active += 1
                if base == "mesh":
                    # sphere_convex measures against ONE face (the least penetrated one the sphere reaches behind): on a hull with
                    # slanted sides the closest point of that polygon need not be the hull's closest point, so the reported distance
                    # is an upper bound of the true one, tight to a fraction of a millimetre here
                    assert dist - 0.06 - 1e-6 <= f["contact_dist"][0] <= dist - 0.06 + 5e-4, (q, f["contact_dist"], dist - 0.06)
                    continue
                assert np.isclose(f["contact_dist"][0], dist - 0.06, atol=1e-6), (q, f["contact_dist"], dist - 0.06)
                # (mjx's closest_segment_point carries a 1e-6 regulariser in its denominator: micrometres along an edge)
                assert np.allclose(f["contact_frame"][0, 0], (closest - q[:3]) / dist, atol=1e-3)
                assert np.allclose(f["contact_pos"][0], closest + 0.5 * (0.06 - dist) * (closest - q[:3]) / dist, atol=1e-5)
            else:
                assert f["contact_dist"][0] > 0
        assert active >= 15


def test_capsule_convex_against_segment_polytope_distance(tmp_path):
    m = _convex_pair(tmp_path, '<geom type="capsule" size=".03 .05"/>')
    o = Oracle(m)
    rng = np.random.default_rng(13)
    g1, g2 = int(m.pair_geom1[0]), int(m.pair_geom2[0])
    active = 0
    for _ in range(80):
        q = np.concatenate([rng.uniform(-0.12, 0.12, 2), [rng.uniform(0.2, 0.3)], rng.normal(size=4)])  # over the top face, away from its edges
        q[3:] /= np.linalg.norm(q[3:])
        xpos, xrot, _, _ = ind.kinematics(m, q)
        V = ind.geom_world_vertices(m, g2, xpos, xrot)
        R = xrot[int(m.geom_bodyid[g1])] @ ind._qmat(m.geom_quat[g1])
        c = xpos[int(m.geom_bodyid[g1])] + xrot[int(m.geom_bodyid[g1])] @ m.geom_pos[g1]
        a, b = c - 0.05 * R[:, 2], c + 0.05 * R[:, 2]
        dist = ind.segment_polytope_distance(a, b, V)
        f = o.forward(q, np.zeros(6))
        if dist < 1e-6:
            continue  # the axis itself is inside the box
        if dist - 0.03 < -1e-6:
            active += 1
            assert np.isclose(f["contact_dist"].min(), dist - 0.03, atol=1e-6), (q, f["contact_dist"], dist - 0.03)
            assert np.allclose(f["contact_frame"][int(np.argmin(f["contact_dist"])), 0], [0, 0, -1], atol=1e-6)  # capsule -> box, through the top face
        elif dist - 0.03 > 1e-6:
            assert not (f["contact_dist"] < 0).any()
    assert active >= 20
