"""Loader contract, mirroring the reference's tests/test_model_io.py (path resolution :23-45, every
bundled model loads with and without force_float :49-54, bh280 sizes :76,98,143,146)."""
import os
from pathlib import Path

import numpy as np
import pytest

from ambersim_b200 import ROOT, mjx
from ambersim_b200.utils import mjcf
from ambersim_b200.utils._internal_utils import _check_filepath
from ambersim_b200.utils.io_utils import load_mj_model_from_file, mj_to_mjx_model_and_data


def test_path_resolution(tmp_path, monkeypatch):
    rel = "models/pendulum/pendulum.xml"
    absolute = Path(ROOT) / rel
    for p in (absolute, str(absolute), rel, Path(rel)):
        assert Path(_check_filepath(p)).samefile(absolute)
    # cwd-relative takes precedence over the package root
    local = tmp_path / "local.xml"
    local.write_text(absolute.read_text())
    monkeypatch.chdir(tmp_path)
    assert Path(_check_filepath("local.xml")).samefile(local)
    with pytest.raises(FileNotFoundError):
        _check_filepath("does/not/exist.xml")


def test_all_models_load():
    files = sorted(Path(ROOT, "models").rglob("*.xml"))
    assert len(files) >= 5
    for f in files:
        for ff in (False, True):
            if ff and f.name in ("barkour_vb_standin.xml", "biped_exo_standin.xml"):
                continue  # already floating
            m = load_mj_model_from_file(f, force_float=ff)
            assert m.nq >= m.nv > 0


def test_bh280_dimensions_and_force_float():
    m = load_mj_model_from_file("models/barrett_hand/bh280.xml")
    assert (m.nq, m.nv, m.nu, m.neq) == (8, 8, 4, 4)
    assert m.names["actuator"] == [f"bh_j{j}_joint_actuator" for j in (32, 11, 12, 22)]
    assert all(n.endswith("_equality") for n in m.names["equality"])
    assert int(m.jnt_limited.sum()) == 8
    mf = load_mj_model_from_file("models/barrett_hand/bh280.xml", force_float=True)
    assert (mf.nq, mf.nv) == (15, 14)
    assert mf.jnt_type[0] == 0 and mf.body_jntnum[1] == 1


def test_solver_and_iteration_overrides():
    m = load_mj_model_from_file("models/pendulum/scene.xml")
    assert (m.opt.solver, m.opt.iterations, m.opt.ls_iterations, m.opt.timestep) == (2, 5, 10, 0.02)
    m = load_mj_model_from_file("models/pendulum/scene.xml", solver="cg", iterations=3, ls_iterations=7)
    assert (m.opt.solver, m.opt.iterations, m.opt.ls_iterations) == (1, 3, 7)
    with pytest.raises(ValueError):
        load_mj_model_from_file("models/pendulum/scene.xml", solver="pgs")


def test_standin_dimensions_are_frozen():
    b = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
    assert (b.nq, b.nv, b.nu, b.nbody, b.npair) == (19, 18, 12, 14, 4)
    assert b.opt.disableflags == 16384 and b.opt.iterations == 1 and b.opt.ls_iterations == 5
    p = load_mj_model_from_file("models/biped_standin/biped_exo_standin.xml")
    assert (p.nq, p.nv, p.nu, p.nbody, p.npair) == (28, 27, 21, 23, 8)


def test_compile_constants():
    m = load_mj_model_from_file("models/pendulum/scene.xml")
    assert np.isclose(m.stat.meaninertia, 0.087959 + 0.25)
    assert np.isclose(m.dof_invweight0[0], 1 / (0.087959 + 0.25))
    assert np.allclose(m.body_subtreemass, [3, 3, 1])
    b = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
    # free joint: translational and rotational invweights averaged separately
    assert np.allclose(b.dof_invweight0[:3], b.dof_invweight0[0]) and np.allclose(b.dof_invweight0[3:6], b.dof_invweight0[3])
    assert np.isclose(b.dof_invweight0[0], np.linalg.inv(b.qM0)[:3, :3].diagonal().mean())
    assert b.body_invweight0[0].sum() == 0 and np.all(b.body_invweight0[1:] > 0)


def test_model_mirror_and_unsupported(tmp_path):
    mj = load_mj_model_from_file("models/barrett_hand/bh280.xml")
    model = mjx.device_put(mj)
    m2 = model.replace(opt=model.opt.replace(timestep=0.002, iterations=1, ls_iterations=4, disableflags=mjx.DisableBit.CONTACT))
    assert m2.opt.iterations == 1 and model.opt.iterations == 100 and m2.opt.disableflags == 16 and m2.nq == 8
    assert np.allclose(model.actuator_ctrlrange, [[-30, 30]] * 4)
    # plane - box and box - sphere are served (convex vertex sets / hull faces); a cylinder is outside the engine's collision
    # functions (as it is outside MJX's): reported like MJX's NotImplementedError
    xml = tmp_path / "box.xml"
    xml.write_text("""<mujoco><worldbody><geom type="plane" size="1 1 .1"/>
      <body pos="0 0 1"><freejoint/><inertial pos="0 0 0" mass="1" diaginertia=".1 .1 .1"/>
      <geom type="box" size=".1 .2 .3"/></body>
      <body pos="1 0 1"><freejoint/><inertial pos="0 0 0" mass="1" diaginertia=".1 .1 .1"/>
      <geom type="sphere" size=".1"/></body>
      <body pos="2 0 1"><freejoint/><inertial pos="0 0 0" mass="1" diaginertia=".1 .1 .1"/>
      <geom type="cylinder" size=".1 .2"/></body></worldbody></mujoco>""")
    mjb = load_mj_model_from_file(xml)
    assert sorted(mjb.pair_kind.tolist()) == [0, 5, 6] and mjb.nvert == 8 and mjb.geom_vertnum.tolist() == [0, 8, 0, 0]
    assert np.allclose(np.abs(mjb.vert), [[0.1, 0.2, 0.3]] * 8) and len({tuple(v) for v in mjb.vert.tolist()}) == 8
    bad = mjx.device_put(mjb)
    assert bad.n_unsupported_pairs == 3  # the cylinder against the plane, the box and the sphere
    with pytest.raises(NotImplementedError):
        bad.handle(0)
    other = tmp_path / "x.sdf"  # neither .urdf nor .xml: refused like the reference (io_utils.py:203-204)
    other.write_text("<sdf/>")
    with pytest.raises(NotImplementedError):
        load_mj_model_from_file(other)
    urdf = tmp_path / "x.urdf"
    urdf.write_text("<robot name='x'/>")
    with pytest.raises(ValueError):  # a URDF without links has no root link
        load_mj_model_from_file(urdf)


def test_default_classes_and_keyframes():
    b = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
    assert np.allclose(b.dof_damping[6:], 1.0) and np.allclose(b.dof_armature[6:], 0.0111)
    assert np.allclose(b.actuator_gainprm[:, 0], 60) and np.allclose(b.actuator_biasprm[:, 1], -60)
    assert np.allclose(b.jnt_axis[3], [0, -1, 0])  # knee class
    assert np.allclose(b.key_ctrl("home"), [0, 0.5, 1.0] * 4)
    assert b.pair_kind.tolist() == [0, 0, 0, 0] and np.allclose(b.pair_friction[0], [1.0, 1.0, 0.02, 0.01, 0.01])  # elementwise max with the plane


def test_real_mjmodel_layout_round_trip(tmp_path):
    """mjx.device_put(mujoco.MjModel) (io_utils.py:225): the adapter reads MuJoCo's own field names and shapes. `mujoco` is not
    installable here, so every model of the suite (and a few random ones) is written out in mujoco.MjModel's layout and read back:
    all model fields, the derived contact pairs, the options and the oracle's trajectory must survive unchanged."""
    import numpy as np

    from ambersim_b200 import mjx
    from ambersim_b200.utils import mjmodel
    from ambersim_b200.utils.io_utils import load_mj_model_from_file
    from oracle.oracle import Oracle
    from tests._randmodel import random_limb_model
    from tests.conftest import MODELS

    paths = [v[0] for v in MODELS.values()]
    for seed in (1, 4):
        f = tmp_path / f"r{seed}.xml"
        f.write_text(random_limb_model(seed, 3, 1, 8)[0])
        paths.append(str(f))
    for path in paths:
        m = load_mj_model_from_file(path)
        ns = mjmodel.to_mujoco_layout(m)
        assert mjmodel.looks_like_mjmodel(ns) and not mjmodel.looks_like_mjmodel(m)
        assert ns.actuator_trnid.shape == (m.nu, 2) and ns.actuator_gear.shape == (m.nu, 6) and ns.actuator_gainprm.shape == (m.nu, 10)
        m2 = mjmodel.from_mjmodel(ns)
        for f in mjx._MODEL_FIELDS:
            a, b = np.asarray(getattr(m, f)), np.asarray(getattr(m2, f))
            assert a.shape == b.shape and np.array_equal(a, b), (path, f)
        assert all(np.array_equal(getattr(m2.opt, k), getattr(m.opt, k)) for k in vars(m.opt))
        assert m2.stat.meaninertia == m.stat.meaninertia and m2.n_unsupported_pairs == m.n_unsupported_pairs
        assert len(m2.keyframes) == len(m.keyframes)
        model = mjx.device_put(ns)  # the public entry point takes the foreign object directly
        assert (model.nq, model.nv, model.nu, model.npair) == (m.nq, m.nv, m.nu, m.npair)
        x0 = np.concatenate([m.qpos0, np.zeros(m.nv)])[None]
        us = np.zeros((1, 5, m.nu))
        kw = dict(disableflags=m.opt.disableflags | 16) if m.n_unsupported_pairs else {}
        assert np.array_equal(Oracle(m, m.opt.replace(**kw)).rollout(x0, us), Oracle(m2, m2.opt.replace(**kw)).rollout(x0, us))


def test_real_mjmodel_unsupported_features_raise():
    import numpy as np
    import pytest

    from ambersim_b200.utils import mjmodel
    from ambersim_b200.utils.io_utils import load_mj_model_from_file

    base = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
    for mutate in (lambda ns: setattr(ns, "na", 2), lambda ns: setattr(ns, "ntendon", 1), lambda ns: ns.jnt_type.__setitem__(3, 1),
                   lambda ns: ns.actuator_dyntype.__setitem__(0, 1), lambda ns: ns.actuator_trntype.__setitem__(0, 3),
                   lambda ns: setattr(ns.opt, "integrator", 2), lambda ns: ns.dof_frictionloss.__setitem__(7, 0.1)):
        ns = mjmodel.to_mujoco_layout(base)
        mutate(ns)
        with pytest.raises(NotImplementedError):
            mjmodel.from_mjmodel(ns)
    ns = mjmodel.to_mujoco_layout(base)
    ns.opt.cone = 1  # elliptic cones: loads (like a model with unsupported geoms), refuses to run with contacts on
    assert mjmodel.from_mjmodel(ns).n_unsupported_pairs > 0


def test_mujoco_model_dump_if_present():
    """Real-MuJoCo pin of the MJCF compiler and of the mujoco.MjModel adapter: consumed when tools/dump_mjx_golden.py has been
    run on a machine that has mujoco (it stores the compiled MjModel field by field next to the MJX rollouts)."""
    from pathlib import Path
    from types import SimpleNamespace

    import numpy as np
    import pytest

    from ambersim_b200 import mjx
    from ambersim_b200.utils import mjmodel
    from ambersim_b200.utils.io_utils import load_mj_model_from_file
    from tests.conftest import MODELS

    gold = Path(__file__).parent / "golden"
    used = 0
    for name in ("pendulum", "bh280", "barkour", "biped", "exolegs"):
        path = gold / f"mjx_{name}.npz"
        if not path.exists():
            continue
        z = np.load(path)
        if "model_nq" not in z.files:
            continue
        ns = SimpleNamespace(opt=SimpleNamespace(), stat=SimpleNamespace(meaninertia=float(z["model_stat_meaninertia"])))
        for k in z.files:
            if k.startswith("model_opt_"):
                setattr(ns.opt, k[10:], z[k] if z[k].ndim else z[k].item())
            elif k.startswith("model_") and k != "model_stat_meaninertia":
                setattr(ns, k[6:], z[k] if z[k].ndim else z[k].item())
        theirs = mjmodel.from_mjmodel(ns)
        ours = load_mj_model_from_file(MODELS[name][0])
        for f in mjx._MODEL_FIELDS:
            a, b = np.asarray(getattr(ours, f), dtype=np.float64), np.asarray(getattr(theirs, f), dtype=np.float64)
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-6, atol=1e-9), (name, f)
        assert np.isclose(ours.stat.meaninertia, theirs.stat.meaninertia, rtol=1e-6)
        used += 1
    if not used:
        pytest.skip("no MuJoCo model dump (mujoco is not installable in this image)")


REF_BH280 = "/root/reference/ambersim/models/barrett_hand/bh280.xml"


@pytest.mark.skipif(not __import__("os").path.exists(REF_BH280), reason="the reference tree is only present in the build container")
def test_reference_bh280_file_with_its_collision_meshes():
    """The reference's own model file (89 collision hulls from OBJ files, `bh280.xml:57-187`) goes through the loader: the articulated
    model equals the shipped geometry-free fixture, the hulls are what SURVEY.md lists, and the trace-time pair enumeration shows why
    the reference's sampler test switches contacts off (`tests/trajopt/test_predictive_sampler.py:29`)."""
    from ambersim_b200.utils import mjcf
    from oracle.oracle import Oracle

    ref = load_mj_model_from_file(REF_BH280)
    own = load_mj_model_from_file("models/barrett_hand/bh280.xml")
    assert (ref.nq, ref.nv, ref.nu, ref.nbody, ref.neq) == (own.nq, own.nv, own.nu, own.nbody, own.neq) == (8, 8, 4, 10, 4)
    for k in ("body_mass", "body_inertia", "body_ipos", "body_pos", "jnt_axis", "jnt_range", "dof_invweight0", "body_invweight0", "eq_data",
              "actuator_ctrlrange", "dof_armature", "dof_damping"):
        assert np.allclose(getattr(ref, k), getattr(own, k), rtol=1e-9, atol=1e-12), k
    assert np.isclose(ref.stat.meaninertia, own.stat.meaninertia, rtol=1e-12)
    col = [g for g in range(ref.ngeom) if ref.geom_type[g] == mjcf.GEOM_MESH and "collision" in ref.names["geom"][g]]
    assert len(col) == 89 and ref.ngeom == 98
    nv = ref.geom_vertnum[col]
    assert nv.min() >= 4 and nv.max() == 1006  # hull vertices of the _col_ meshes
    for g in col:  # every hull is a closed polyhedron: Euler's formula
        assert ref.geom_vertnum[g] - ref.geom_edgenum[g] + ref.geom_facenum[g] == 2
    # what MJX would enumerate at trace time with the default contype / conaffinity: thousands of hull - hull pairs
    assert ref.npair + ref.n_unsupported_pairs > 3000 and ref.n_unsupported_pairs > 0 and "hull vertices" in ref.unsupported_reason
    # with the reference test's options (contacts off) the constraint set is the fixture's: 4 equalities + 8 limits
    opt = ref.opt.replace(disableflags=16, iterations=1, ls_iterations=4, timestep=0.002)
    o = Oracle(ref, opt)
    assert (o.ncon, o.ne, o.nl, o.nefc) == (0, 4, 8, 12)
    f_ref = o.forward(0.1 * np.ones(8), np.zeros(8))
    f_own = Oracle(own, own.opt.replace(disableflags=16, iterations=1, ls_iterations=4, timestep=0.002)).forward(0.1 * np.ones(8), np.zeros(8))
    assert np.allclose(f_ref["qacc"], f_own["qacc"], rtol=1e-9, atol=1e-9)


def test_state_and_data_are_pytrees():
    """`State` / `mjx.Data` flatten like the flax dataclasses they mirror (rl/base.py:14-32): brax-style wrappers written with
    tree_map (`where(done, first, state)` over the whole state) work on them."""
    import torch
    from torch.utils import _pytree

    from ambersim_b200.rl.base import State

    E = 3
    d = mjx.Data(qpos=torch.ones(E, 2), qvel=torch.zeros(E, 2), ctrl=torch.zeros(E, 1), qacc=torch.zeros(E, 2), qacc_warmstart=torch.zeros(E, 2),
                 time=torch.zeros(E))
    s = State(d, obs=torch.arange(E * 4.0).reshape(E, 4), reward=torch.zeros(E), done=torch.tensor([0.0, 1.0, 0.0]), metrics={"r": torch.ones(E)}, info={})
    first = _pytree.tree_map(lambda x: x * 0 + 7, s)
    assert isinstance(first, State) and isinstance(first.pipeline_state, mjx.Data) and first.pipeline_state.xpos is None
    done = s.done.bool()
    blend = _pytree.tree_map(lambda a, b: torch.where(done.reshape(done.shape + (1,) * (b.dim() - 1)), a, b), first, s)
    assert blend.pipeline_state.qpos.tolist() == [[1, 1], [7, 7], [1, 1]] and blend.obs[1].tolist() == [7, 7, 7, 7] and blend.metrics["r"].tolist() == [1, 7, 1]
    leaves, spec = _pytree.tree_flatten(s)
    assert all(isinstance(x, torch.Tensor) for x in leaves) and _pytree.tree_unflatten(leaves, spec).obs is s.obs


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return (np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
            @ np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]]))


def test_urdf_front_end_on_an_authored_arm():
    """load_mj_model_from_file on a URDF (reference io_utils.py:196-204): MuJoCo's import conventions + ambersim's
    actuators-from-transmissions (io_utils.py:18-69) and equalities-from-mimics (io_utils.py:72-121), checked field by field against
    values computed here from the URDF's numbers."""
    from oracle.independent import _qmat
    from oracle.oracle import Oracle

    m = load_mj_model_from_file("tests/models/arm.urdf")
    assert m.names["body"] == ["world", "base", "upper", "slider", "tool", "finger"]  # welded links stay bodies (fusestatic is not applied)
    assert m.body_parentid.tolist() == [0, 0, 1, 2, 3, 4] and (m.nq, m.nv, m.nu, m.neq) == (3, 3, 2, 1)
    assert m.names["joint"] == ["shoulder", "extend", "curl"] and m.jnt_type.tolist() == [mjcf.JNT_HINGE, mjcf.JNT_SLIDE, mjcf.JNT_HINGE]
    assert m.jnt_bodyid.tolist() == [2, 3, 5]  # the tool is welded to the slider: no joint of its own
    # joint origin -> child body frame (fixed-axis roll / pitch / yaw)
    assert np.allclose(m.body_pos[2], [0, 0, 0.1]) and np.allclose(_qmat(m.body_quat[2]), _rpy(0.2, 0.4, -0.6), atol=1e-12)
    assert np.allclose(m.body_pos[4], [0, 0, 0.2]) and np.allclose(_qmat(m.body_quat[4]), _rpy(0, 0, 1.0), atol=1e-12)
    assert np.allclose(m.jnt_pos, 0) and np.allclose(m.jnt_axis, [[0, 0, 1], [1, 0, 0], [0, 1, 0]])
    assert m.jnt_limited.tolist() == [1, 1, 0] and np.allclose(m.jnt_range[:2], [[-1.5, 2.0], [0, 0.25]])  # continuous: unlimited
    assert np.allclose(m.dof_damping, [0.4, 2.0, 0.0])
    # inertia: the tensor is given in the frame `origin rpy`; in the body frame it is R I R'
    I = np.array([[0.004, 0.0006, -0.0003], [0.0006, 0.009, 0.0002], [-0.0003, 0.0002, 0.008]])
    R = _rpy(0.3, -0.2, 0.5)
    got = _qmat(m.body_iquat[2]) @ np.diag(m.body_inertia[2]) @ _qmat(m.body_iquat[2]).T
    assert np.allclose(got, R @ I @ R.T, atol=1e-15) and np.allclose(m.body_ipos[2], [0.1, 0.02, 0]) and np.isclose(m.body_mass[2], 1.2)
    # one motor per transmission: limited to +- effort where the joint has one
    assert m.names["actuator"] == ["shoulder_actuator", "extend_actuator"] and m.actuator_trnid.tolist() == [0, 1]
    assert m.actuator_ctrllimited.tolist() == [1, 0] and np.allclose(m.actuator_ctrlrange[0], [-12.5, 12.5])
    # one joint equality per mimic: q_curl = 0.1 + 0.5 q_shoulder
    assert m.names["equality"] == ["curl_shoulder_equality"] and (m.eq_obj1id[0], m.eq_obj2id[0]) == (2, 0)
    assert np.allclose(m.eq_data[0, :5], [0.1, 0.5, 0, 0, 0])
    # collision geoms only (MuJoCo's URDF default discards visuals); sizes are half extents / half lengths
    assert m.names["geom"][:2] == ["base_col", "upper_col"] and m.ngeom == 3 and m.geom_type.tolist() == [mjcf.GEOM_BOX, mjcf.GEOM_CAPSULE, mjcf.GEOM_SPHERE]
    assert np.allclose(m.geom_size[0], [0.1, 0.1, 0.05]) and np.allclose(m.geom_size[1, :2], [0.03, 0.15]) and np.allclose(m.geom_pos[1], [0.2, 0, 0])
    assert np.allclose(_qmat(m.geom_quat[1]), _rpy(0, 1.5708, 0), atol=1e-12)
    f = Oracle(m, m.opt.replace(disableflags=16)).forward([0.3, 0.1, 0.25], [0.2, -0.1, 0.1], [1.0, -0.5])
    assert np.isfinite(f["qacc"]).all() and f["efc_J"].shape[0] >= 1


REF_MODELS = "/root/reference/ambersim/models"


@pytest.mark.skipif(not __import__("os").path.exists(REF_MODELS), reason="the reference tree is only present in the build container")
def test_reference_urdfs_load_like_the_xml_mujoco_made_from_them():
    """The reference ships each model as a URDF and as the MJCF that MuJoCo's compiler + ambersim's `_add_actuators` / `_add_mimics`
    produced from it (`save_model_xml`, io_utils.py:196-204). Loading the URDF through this repo's front end must give the model that
    loading that MJCF gives: the MJCF is real MuJoCo output, printed to 6 digits."""
    from oracle.independent import _qmat

    u = load_mj_model_from_file(REF_MODELS + "/barrett_hand/bh280.urdf")
    x = load_mj_model_from_file(REF_MODELS + "/barrett_hand/bh280.xml")
    assert (u.nq, u.nv, u.nu, u.nbody, u.ngeom, u.neq) == (x.nq, x.nv, x.nu, x.nbody, x.ngeom, x.neq) == (8, 8, 4, 10, 98, 4)
    for k in ("body", "joint", "geom", "actuator", "equality"):
        assert u.names[k] == x.names[k], k
    for k in ("body_mass", "body_pos", "body_ipos", "jnt_axis", "jnt_range", "actuator_ctrlrange", "eq_data", "geom_pos", "geom_quat", "geom_size"):
        assert np.allclose(getattr(u, k), getattr(x, k), rtol=0, atol=1e-12), k
    assert np.allclose(u.body_quat, x.body_quat, atol=1e-6) and np.array_equal(u.geom_vertnum, x.geom_vertnum)
    assert np.array_equal(u.geom_contype, x.geom_contype) and np.array_equal(u.pair_geom1, x.pair_geom1) and u.n_unsupported_pairs == x.n_unsupported_pairs
    for b in range(1, u.nbody):  # the same inertia tensor in the body frame (the MJCF carries quat + diaginertia to 6 digits)
        Tu = _qmat(u.body_iquat[b]) @ np.diag(u.body_inertia[b]) @ _qmat(u.body_iquat[b]).T
        Tx = _qmat(x.body_iquat[b]) @ np.diag(x.body_inertia[b]) @ _qmat(x.body_iquat[b]).T
        assert np.abs(Tu - Tx).max() < 5e-6 * np.abs(Tx).max(), b
    assert np.allclose(u.dof_invweight0, x.dof_invweight0, rtol=1e-6) and np.isclose(u.stat.meaninertia, x.stat.meaninertia, rtol=1e-6)
    p = load_mj_model_from_file(REF_MODELS + "/pendulum/pendulum.urdf")
    q = load_mj_model_from_file(REF_MODELS + "/pendulum/pendulum.xml")
    assert (p.nq, p.nu, p.nbody) == (q.nq, q.nu, q.nbody) == (1, 1, 3) and p.names["body"] == q.names["body"] and p.names["actuator"] == q.names["actuator"]
    assert np.allclose(p.body_mass, q.body_mass) and np.allclose(p.body_inertia, q.body_inertia) and np.allclose(p.actuator_ctrlrange, q.actuator_ctrlrange)
    assert np.allclose(p.dof_invweight0, q.dof_invweight0, rtol=1e-9) and np.allclose(p.jnt_range, [[-3.1416, 3.1416]])


def test_introspection_helpers_and_urdf_actuators():
    """The reference's loader test counts actuators against the URDF's transmissions and equalities against its mimic joints
    (tests/test_model_io.py:66-100), through its introspection helpers."""
    from ambersim_b200.utils.introspection_utils import get_actuator_names, get_equality_names, get_geom_names, get_joint_names

    m = load_mj_model_from_file("tests/models/arm.urdf")
    assert get_actuator_names(m) == ["shoulder_actuator", "extend_actuator"] and get_equality_names(m) == ["curl_shoulder_equality"]
    assert get_joint_names(m) == ["shoulder", "extend", "curl"] and len(get_geom_names(m)) == m.ngeom
    b = load_mj_model_from_file("models/barrett_hand/bh280.xml")
    assert len(get_actuator_names(b)) == 4 and len(get_equality_names(b)) == 4 and len(get_joint_names(b)) == 8


def test_save_model_xml_round_trip(tmp_path):
    """save_model_xml (reference conversion_utils.py:11-37, tests/test_model_io.py:57-63): the MJCF saved from a URDF loads to the same model."""
    from ambersim_b200.utils.conversion_utils import save_model_xml

    out = tmp_path / "arm.xml"
    save_model_xml("tests/models/arm.urdf", out)
    a, b = load_mj_model_from_file("tests/models/arm.urdf"), load_mj_model_from_file(out)
    assert a.names == b.names and (a.nq, a.nv, a.nu, a.neq, a.ngeom) == (b.nq, b.nv, b.nu, b.neq, b.ngeom)
    for k in ("body_pos", "body_quat", "body_mass", "body_inertia", "body_ipos", "jnt_axis", "jnt_range", "dof_damping", "eq_data", "actuator_ctrlrange",
              "geom_size", "geom_pos", "dof_invweight0", "body_invweight0"):
        assert np.allclose(getattr(a, k), getattr(b, k), rtol=1e-9, atol=1e-12), k


def test_stl_meshes_binary_and_ascii(tmp_path):
    """Mesh geoms from STL files (what URDFs usually reference): binary and ASCII give the hull of the same tetrahedron."""
    import struct

    pts = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1]], dtype=np.float32)
    tris = [(0, 2, 1), (0, 1, 3), (0, 3, 2), (1, 2, 3)]
    with open(tmp_path / "tet.stl", "wb") as f:
        f.write(b"binary tetrahedron".ljust(80, b" ") + struct.pack("<I", len(tris)))
        for t in tris:
            f.write(struct.pack("<3f", 0, 0, 0) + b"".join(struct.pack("<3f", *pts[i]) for i in t) + struct.pack("<H", 0))
    lines = ["solid tet"] + [ln for t in tris for ln in (["facet normal 0 0 0", "outer loop"] + [f"vertex {pts[i][0]} {pts[i][1]} {pts[i][2]}" for i in t] + ["endloop", "endfacet"])] + ["endsolid tet"]
    (tmp_path / "tet_ascii.stl").write_text("\n".join(lines))
    xml = tmp_path / "m.xml"
    xml.write_text("""<mujoco><asset><mesh name="a" file="tet.stl"/><mesh name="b" file="tet_ascii.stl" scale="2 2 2"/></asset><worldbody>
      <geom type="plane" size="1 1 .1"/><body pos="0 0 1"><freejoint/><inertial pos="0 0 0" mass="1" diaginertia=".1 .1 .1"/>
      <geom type="mesh" mesh="a"/><geom type="mesh" mesh="b" pos=".3 0 0"/></body></worldbody></mujoco>""")
    m = load_mj_model_from_file(xml)
    assert m.geom_vertnum.tolist() == [0, 4, 4] and m.geom_facenum.tolist() == [0, 4, 4] and m.geom_edgenum.tolist() == [0, 6, 6]
    va, vb = m.vert[:4], m.vert[4:]
    assert {tuple(np.round(v, 6)) for v in va} == {tuple(np.round(p, 6)) for p in pts.astype(np.float64)}
    assert {tuple(np.round(v, 6)) for v in vb} == {tuple(np.round(2 * p, 6)) for p in pts.astype(np.float64)}


def test_shipped_urdfs_load_like_the_shipped_mjcf(tmp_path, monkeypatch):
    """The package ships its two reference models as URDF and as MJCF, like the reference (tests/test_model_io.py:24-47 loads
    `models/pendulum/pendulum.urdf` by global, local and package-relative paths, str or Path): both forms give the same model."""
    for u, x in (("models/pendulum/pendulum.urdf", "models/pendulum/pendulum.xml"), ("models/barrett_hand/bh280.urdf", "models/barrett_hand/bh280.xml")):
        a, b = load_mj_model_from_file(u), load_mj_model_from_file(x)
        assert (a.nq, a.nv, a.nu, a.nbody, a.neq) == (b.nq, b.nv, b.nu, b.nbody, b.neq) and a.names["actuator"] == b.names["actuator"]
        assert np.allclose(a.body_mass, b.body_mass) and np.allclose(a.dof_invweight0, b.dof_invweight0, rtol=1e-6)
        assert np.allclose(a.eq_data, b.eq_data) and np.allclose(a.actuator_ctrlrange, b.actuator_ctrlrange)
        if "bh280" in u:  # (the reference's pendulum.xml was edited by hand after the conversion: its joint lost the URDF's limit)
            assert np.allclose(a.jnt_range, b.jnt_range)
        else:
            assert a.jnt_limited.tolist() == [1] and np.allclose(a.jnt_range, [[-3.1416, 3.1416]])
    glob = ROOT + "/models/pendulum/pendulum.urdf"
    local = tmp_path / "pendulum.urdf"
    local.write_text(Path(glob).read_text())
    monkeypatch.chdir(tmp_path)
    for p in (glob, Path(glob), "pendulum.urdf", Path("pendulum.urdf"), "models/pendulum/pendulum.urdf", Path("models/pendulum/pendulum.urdf")):
        assert load_mj_model_from_file(p).nq == 1
    assert load_mj_model_from_file("models/barrett_hand/bh280.urdf", force_float=True).nq == 15  # reference tests/test_model_io.py:146
