"""Flatten a real `mujoco.MjModel` into the host model the engine uploads: the producer behind
`mjx.device_put(mj_model)` at the reference's call sites (ambersim/utils/io_utils.py:225, ambersim/rl/base.py:52)
when the caller has `mujoco` installed and hands over its own MjModel.

`mujoco` itself is never imported: the adapter is duck-typed on MuJoCo's public field names and shapes
(mjmodel.h, MuJoCo 3.x), so it also runs on any object exposing them. `to_mujoco_layout` writes one of this repo's
compiled models back out in that layout; the CPU tests use the pair as a round trip, since `mujoco` is not
installable in this image.

Unsupported features raise NotImplementedError where MJX's device_put would (io_utils.py:228-241).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from ambersim_b200.utils import mjcf

# mjtJoint / mjtGeom / mjtTrn / mjtDyn / mjtGain / mjtBias / mjtEq / mjtIntegrator / mjtCone / mjtSolver values
_JNT_FREE, _JNT_BALL, _JNT_SLIDE, _JNT_HINGE = 0, 1, 2, 3
_TRN_JOINT = 0
_DYN_NONE = 0
_GAIN_FIXED, _GAIN_AFFINE = 0, 1
_BIAS_NONE, _BIAS_AFFINE = 0, 1
_EQ_JOINT = 2
_INT_EULER, _INT_RK4 = 0, 1
_CONE_PYRAMIDAL = 0
_SOLVER_CG, _SOLVER_NEWTON = 1, 2


def looks_like_mjmodel(obj) -> bool:
    """True for a mujoco.MjModel (or anything laid out like one) that is not already this repo's host model."""
    return not isinstance(obj, mjcf.MjModel) and all(hasattr(obj, a) for a in ("nbody", "body_parentid", "geom_contype", "actuator_trnid", "opt"))


def _arr(x, dtype, shape=None):
    a = np.array(x, dtype=dtype)
    return a if shape is None else a.reshape(shape)


def _cols(x, rows: int, keep: int) -> np.ndarray:
    """[rows, >= keep] (or empty) -> the first `keep` columns as float64 [rows, keep]."""
    a = np.array(x, dtype=np.float64)
    if rows == 0 or a.size == 0:
        return np.zeros((rows, keep))
    return a.reshape(rows, -1)[:, :keep].copy()


def from_mjmodel(mj) -> mjcf.MjModel:
    """mujoco.MjModel -> flat host model (same field names as mjx.Model; derived pair tables as MJX builds at trace time)."""
    m = mjcf.MjModel()
    nq, nv, nu, nbody, njnt, ngeom, neq = (int(getattr(mj, k)) for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "neq"))
    # ---- features outside the engine (the same refusals MJX's device_put issues for what it lacks)
    if int(getattr(mj, "na", 0)) > 0:
        raise NotImplementedError("stateful actuators (na > 0) are not supported")
    if int(getattr(mj, "ntendon", 0)) > 0:
        raise NotImplementedError("tendons are not supported")
    if int(getattr(mj, "nmocap", 0)) > 0:
        raise NotImplementedError("mocap bodies are not supported")
    o = mj.opt
    if int(o.integrator) not in (_INT_EULER, _INT_RK4):
        raise NotImplementedError("integrator must be Euler or RK4")
    if int(o.solver) not in (_SOLVER_CG, _SOLVER_NEWTON):
        raise NotImplementedError("solver must be CG or Newton")
    jnt_type = _arr(mj.jnt_type, np.int32, (njnt,))
    if np.any(jnt_type == _JNT_BALL):
        raise NotImplementedError("ball joints are not supported")
    if nv and np.any(_arr(getattr(mj, "dof_frictionloss", np.zeros(nv)), np.float64) != 0):
        raise NotImplementedError("joint frictionloss is not supported")
    m.opt = mjcf.Option().replace(
        timestep=float(o.timestep), impratio=float(o.impratio), tolerance=float(o.tolerance), ls_tolerance=float(o.ls_tolerance),
        gravity=_arr(o.gravity, np.float64, (3,)), integrator=int(o.integrator), cone=int(o.cone), jacobian=int(o.jacobian),
        solver=int(o.solver), iterations=int(o.iterations), ls_iterations=int(o.ls_iterations), disableflags=int(o.disableflags))
    m.stat = mjcf.Statistic()
    m.stat.meaninertia = float(mj.stat.meaninertia)
    m.nq, m.nv, m.nu, m.na, m.nbody, m.njnt, m.ngeom, m.neq = nq, nv, nu, 0, nbody, njnt, ngeom, neq
    # ---- bodies, joints, dofs: same names, same shapes
    for k in ("body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr"):
        setattr(m, k, _arr(getattr(mj, k), np.int32, (nbody,)))
    for k, w in (("body_pos", 3), ("body_quat", 4), ("body_ipos", 3), ("body_iquat", 4), ("body_inertia", 3), ("body_invweight0", 2)):
        setattr(m, k, _arr(getattr(mj, k), np.float64, (nbody, w)))
    m.body_mass = _arr(mj.body_mass, np.float64, (nbody,))
    m.body_subtreemass = _arr(mj.body_subtreemass, np.float64, (nbody,))
    m.jnt_type = jnt_type
    for k in ("jnt_qposadr", "jnt_dofadr", "jnt_bodyid"):
        setattr(m, k, _arr(getattr(mj, k), np.int32, (njnt,)))
    m.jnt_limited = _arr(mj.jnt_limited, np.int32, (njnt,))
    for k, w in (("jnt_solref", 2), ("jnt_solimp", 5), ("jnt_pos", 3), ("jnt_axis", 3), ("jnt_range", 2)):
        setattr(m, k, _arr(getattr(mj, k), np.float64, (njnt, w)))
    m.jnt_stiffness = _arr(mj.jnt_stiffness, np.float64, (njnt,))
    m.jnt_margin = _arr(mj.jnt_margin, np.float64, (njnt,))
    for k in ("dof_bodyid", "dof_jntid", "dof_parentid"):
        setattr(m, k, _arr(getattr(mj, k), np.int32, (nv,)))
    for k in ("dof_armature", "dof_damping", "dof_invweight0"):
        setattr(m, k, _arr(getattr(mj, k), np.float64, (nv,)))
    m.body_lastdof = np.array([m.body_dofadr[b] + m.body_dofnum[b] - 1 if m.body_dofnum[b] else -1 for b in range(nbody)], dtype=np.int32)
    m.qpos0 = _arr(mj.qpos0, np.float64, (nq,))
    m.qpos_spring = _arr(mj.qpos_spring, np.float64, (nq,))
    # ---- geoms and the static contact pairs MJX enumerates at trace time
    m.geom_type = _arr(mj.geom_type, np.int32, (ngeom,))
    m.geom_bodyid = _arr(mj.geom_bodyid, np.int32, (ngeom,))
    for k, w in (("geom_size", 3), ("geom_pos", 3), ("geom_quat", 4), ("geom_friction", 3), ("geom_solref", 2), ("geom_solimp", 5)):
        setattr(m, k, _arr(getattr(mj, k), np.float64, (ngeom, w)))
    for k in ("geom_contype", "geom_conaffinity", "geom_condim", "geom_priority"):
        setattr(m, k, _arr(getattr(mj, k), np.int32, (ngeom,)))
    for k in ("geom_solmix", "geom_margin", "geom_gap"):
        setattr(m, k, _arr(getattr(mj, k), np.float64, (ngeom,)))
    # convex vertex sets: the 8 corners of a box, the hull vertices of a mesh (mj.mesh_vert / mesh_vertadr / mesh_vertnum via geom_dataid)
    adr, num, pool = [], [], []
    dataid = _arr(getattr(mj, "geom_dataid", np.full(ngeom, -1)), np.int32, (ngeom,))
    for g in range(ngeom):
        v = np.zeros((0, 3))
        if int(m.geom_type[g]) == mjcf.GEOM_BOX:
            v = mjcf.box_vertices(m.geom_size[g])
        elif int(m.geom_type[g]) == mjcf.GEOM_MESH and dataid[g] >= 0 and hasattr(mj, "mesh_vert"):
            a, n = int(np.ravel(mj.mesh_vertadr)[dataid[g]]), int(np.ravel(mj.mesh_vertnum)[dataid[g]])
            v = mjcf.convex_vertices(_arr(mj.mesh_vert, np.float64).reshape(-1, 3)[a:a + n])
        adr.append(sum(num)); num.append(len(v)); pool.extend(np.asarray(v).tolist())
    m.geom_vertadr, m.geom_vertnum = np.array(adr, dtype=np.int32), np.array(num, dtype=np.int32)
    m.nvert = int(sum(num))
    m.vert = np.array(pool, dtype=np.float64).reshape(m.nvert, 3)
    mjcf.attach_convex_topology(m)
    geoms = [dict(name=f"geom{g}", type=int(m.geom_type[g]), body=int(m.geom_bodyid[g]), contype=int(m.geom_contype[g]),
                  conaffinity=int(m.geom_conaffinity[g]), condim=int(m.geom_condim[g]), priority=int(m.geom_priority[g]),
                  friction=m.geom_friction[g], solmix=float(m.geom_solmix[g]), solref=m.geom_solref[g], solimp=m.geom_solimp[g],
                  margin=float(m.geom_margin[g]), gap=float(m.geom_gap[g])) for g in range(ngeom)]
    excludes = set()
    for sig in _arr(getattr(mj, "exclude_signature", []), np.int64).reshape(-1):  # (body1 << 16) + body2
        b1, b2 = int(sig) >> 16, int(sig) & 0xFFFF
        excludes.add((min(b1, b2), max(b1, b2)))
    explicit = []
    for p in range(int(getattr(mj, "npair", 0))):  # <contact><pair>: parameters already mixed by the MuJoCo compiler
        explicit.append(dict(g1=int(mj.pair_geom1[p]), g2=int(mj.pair_geom2[p]), condim=int(mj.pair_dim[p]),
                             friction=_arr(mj.pair_friction, np.float64).reshape(-1, 5)[p], solref=_arr(mj.pair_solref, np.float64).reshape(-1, 2)[p],
                             solimp=_arr(mj.pair_solimp, np.float64).reshape(-1, 5)[p],
                             includemargin=float(np.ravel(mj.pair_margin)[p]) - float(np.ravel(mj.pair_gap)[p])))
    m.names = dict(body=[f"body{i}" for i in range(nbody)], joint=[f"joint{i}" for i in range(njnt)], geom=[g["name"] for g in geoms],
                   actuator=[f"actuator{i}" for i in range(nu)], equality=[f"eq{i}" for i in range(neq)])
    mjcf.enumerate_pairs(m, geoms, excludes, explicit)
    # ---- equalities (joint couplings only, like the loader)
    eq_type = _arr(mj.eq_type, np.int32, (neq,))
    if np.any(eq_type != _EQ_JOINT):
        raise NotImplementedError("only joint equality constraints are supported")
    m.eq_type = eq_type
    m.eq_obj1id = _arr(mj.eq_obj1id, np.int32, (neq,))
    m.eq_obj2id = _arr(mj.eq_obj2id, np.int32, (neq,))
    m.eq_active = _arr(getattr(mj, "eq_active0", getattr(mj, "eq_active", np.ones(neq))), np.int32, (neq,))
    m.eq_solref = _arr(mj.eq_solref, np.float64, (neq, 2))
    m.eq_solimp = _arr(mj.eq_solimp, np.float64, (neq, 5))
    m.eq_data = _cols(mj.eq_data, neq, 11)
    # ---- actuators: joint transmissions, no activation state; gain/bias parameter vectors cut to the three the engine uses
    if nu:
        if np.any(_arr(mj.actuator_trntype, np.int32, (nu,)) != _TRN_JOINT):
            raise NotImplementedError("only joint transmissions are supported")
        if np.any(_arr(mj.actuator_dyntype, np.int32, (nu,)) != _DYN_NONE):
            raise NotImplementedError("stateful actuators (dyntype) are not supported")
        gaintype, biastype = _arr(mj.actuator_gaintype, np.int32, (nu,)), _arr(mj.actuator_biastype, np.int32, (nu,))
        if np.any((gaintype != _GAIN_FIXED) & (gaintype != _GAIN_AFFINE)) or np.any((biastype != _BIAS_NONE) & (biastype != _BIAS_AFFINE)):
            raise NotImplementedError("actuator gain / bias types must be fixed / affine / none")
        trn = _cols(mj.actuator_trnid, nu, 1)[:, 0].astype(np.int32)
        if np.any((m.jnt_type[trn] != _JNT_HINGE) & (m.jnt_type[trn] != _JNT_SLIDE)):
            raise NotImplementedError("actuators on free joints are not supported")
    else:
        gaintype = biastype = trn = np.zeros(0, dtype=np.int32)
    m.actuator_trnid, m.actuator_gaintype, m.actuator_biastype = trn, gaintype, biastype
    m.actuator_ctrllimited = _arr(mj.actuator_ctrllimited, np.int32, (nu,))
    m.actuator_forcelimited = _arr(mj.actuator_forcelimited, np.int32, (nu,))
    m.actuator_ctrlrange = _arr(mj.actuator_ctrlrange, np.float64, (nu, 2))
    m.actuator_forcerange = _arr(mj.actuator_forcerange, np.float64, (nu, 2))
    m.actuator_gainprm = _cols(mj.actuator_gainprm, nu, 3)
    m.actuator_biasprm = _cols(mj.actuator_biasprm, nu, 3)
    m.actuator_gear = _cols(mj.actuator_gear, nu, 1)[:, 0]
    # ---- keyframes (by index; MuJoCo keeps the names in its name buffer)
    nkey = int(getattr(mj, "nkey", 0))
    for k in range(nkey):
        m.keyframes[f"key{k}"] = dict(qpos=_arr(mj.key_qpos, np.float64).reshape(nkey, nq)[k], qvel=_arr(mj.key_qvel, np.float64).reshape(nkey, nv)[k],
                                      ctrl=_arr(mj.key_ctrl, np.float64).reshape(nkey, nu)[k])
    m.source = "mujoco.MjModel"
    return m


def to_mujoco_layout(m: mjcf.MjModel) -> SimpleNamespace:
    """One of this repo's compiled models written out with mujoco.MjModel's field names and shapes (MuJoCo 3.x:
    actuator_trnid [nu,2], actuator_gear [nu,6], gain / bias parameter vectors of 10, eq_active0, key_* arrays)."""
    o = m.opt
    ns = SimpleNamespace(
        nq=m.nq, nv=m.nv, nu=m.nu, na=0, nbody=m.nbody, njnt=m.njnt, ngeom=m.ngeom, neq=m.neq, ntendon=0, nmocap=0, npair=0,
        opt=SimpleNamespace(timestep=o.timestep, impratio=o.impratio, tolerance=o.tolerance, ls_tolerance=o.ls_tolerance, gravity=np.array(o.gravity),
                            integrator=o.integrator, cone=o.cone, jacobian=o.jacobian, solver=o.solver, iterations=o.iterations,
                            ls_iterations=o.ls_iterations, disableflags=o.disableflags),
        stat=SimpleNamespace(meaninertia=m.stat.meaninertia), exclude_signature=np.array(getattr(m, "exclude_signature", np.zeros(0)), dtype=np.int32))
    same = ("body_parentid body_rootid body_weldid body_jntnum body_jntadr body_dofnum body_dofadr body_pos body_quat body_ipos body_iquat "
            "body_inertia body_invweight0 body_mass body_subtreemass jnt_type jnt_qposadr jnt_dofadr jnt_bodyid jnt_limited jnt_solref jnt_solimp "
            "jnt_pos jnt_axis jnt_range jnt_stiffness jnt_margin dof_bodyid dof_jntid dof_parentid dof_armature dof_damping dof_invweight0 qpos0 "
            "qpos_spring geom_type geom_bodyid geom_size geom_pos geom_quat geom_friction geom_solref geom_solimp geom_contype geom_conaffinity "
            "geom_condim geom_priority geom_solmix geom_margin geom_gap eq_type eq_obj1id eq_obj2id eq_solref eq_solimp eq_data "
            "actuator_gaintype actuator_biastype actuator_ctrllimited actuator_forcelimited actuator_ctrlrange actuator_forcerange").split()
    for k in same:
        setattr(ns, k, np.array(getattr(m, k)))
    # meshes: one mjModel mesh per mesh geom (geom_dataid -> mesh_vertadr / mesh_vertnum / mesh_vert), boxes carry no data
    dataid, madr, mnum, mv = [], [], [], []
    for g in range(m.ngeom):
        if int(m.geom_type[g]) == mjcf.GEOM_MESH:
            dataid.append(len(madr)); madr.append(len(mv)); mnum.append(int(m.geom_vertnum[g]))
            mv.extend(np.asarray(m.vert)[m.geom_vertadr[g]:m.geom_vertadr[g] + m.geom_vertnum[g]].tolist())
        else:
            dataid.append(-1)
    ns.geom_dataid = np.array(dataid, dtype=np.int32)
    ns.mesh_vertadr, ns.mesh_vertnum = np.array(madr, dtype=np.int32), np.array(mnum, dtype=np.int32)
    ns.mesh_vert = np.array(mv, dtype=np.float64).reshape(-1, 3)
    ns.nmesh = len(madr)
    ns.dof_frictionloss = np.zeros(m.nv)
    ns.eq_active0 = np.array(m.eq_active, dtype=np.uint8)
    nu = m.nu
    ns.actuator_trntype = np.zeros(nu, dtype=np.int32)
    ns.actuator_dyntype = np.zeros(nu, dtype=np.int32)
    ns.actuator_trnid = np.stack((np.array(m.actuator_trnid, dtype=np.int32), -np.ones(nu, dtype=np.int32)), axis=1).reshape(nu, 2)
    gear = np.zeros((nu, 6))
    gear[:, 0] = m.actuator_gear
    ns.actuator_gear = gear
    gp, bp = np.zeros((nu, 10)), np.zeros((nu, 10))
    gp[:, :3], bp[:, :3] = np.array(m.actuator_gainprm).reshape(nu, 3), np.array(m.actuator_biasprm).reshape(nu, 3)
    ns.actuator_gainprm, ns.actuator_biasprm = gp, bp
    keys = list(m.keyframes.values())
    ns.nkey = len(keys)
    ns.key_qpos = np.array([k["qpos"] for k in keys]).reshape(len(keys), m.nq)
    ns.key_qvel = np.array([k["qvel"] for k in keys]).reshape(len(keys), m.nv)
    ns.key_ctrl = np.array([k["ctrl"] for k in keys]).reshape(len(keys), m.nu)
    return ns
