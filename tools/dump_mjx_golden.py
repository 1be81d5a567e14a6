"""Run on a machine that HAS mujoco + mujoco-mjx + jax (not this image): dumps, for each of the repo's five models and three convex-collision fixtures,
(1) a reference rollout through the reference's own `shoot`, (2) every per-stage `mjx.Data` field of one `mjx.forward`
at a seeded state (controls and warm start included) and (3) the compiled `MjModel` field by field (mj_setConst
constants, defaults, frames) into tests/golden/mjx_<model>.npz. tests/test_oracle_physics.py
(test_mjx_golden_if_present, test_mjx_golden_stages_if_present) and tests/test_model_io.py
(test_mujoco_model_dump_if_present) consume them and PIN the oracle and the loader against real MJX / MuJoCo.

One command, anywhere with network:   pip install "mujoco>=3.0.1,<3.2" "mujoco-mjx>=3.0.1,<3.2" "jax[cpu]" && python tools/dump_mjx_golden.py
then commit tests/golden/mjx_*.npz. Until such a file lands, parity with MJX is UNPINNED (README).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
try:
    import jax
    import jax.numpy as jnp
    import mujoco
    from mujoco import mjx
except ImportError as e:  # pragma: no cover
    sys.exit(f"needs jax + mujoco + mujoco-mjx: {e}")

MODELS = {"pendulum": ("pendulum/scene.xml", None, {}), "bh280": ("barrett_hand/bh280.xml", None, dict(timestep=0.002, iterations=1, ls_iterations=4, disableflags=16)),
          "barkour": ("barkour_standin/barkour_vb_standin.xml", "home", {}), "biped": ("biped_standin/biped_exo_standin.xml", "stand", {}),
          "exolegs": ("biped_standin/exo_legs_standin.xml", "stand", {}),
          # convex collision fixtures (paths from the repo root): plane / sphere / capsule / convex - convex pairs of collision_convex.py
          "boxbot": ("tests/models/boxbot.xml", "home", {}), "blocks": ("tests/models/blocks.xml", "home", {}),
          "bh280_hulls": ("tests/models/bh280_hulls.xml", None, dict(timestep=0.002, iterations=2, ls_iterations=6))}
STAGES = ("xpos xquat xmat xipos ximat xanchor xaxis geom_xpos geom_xmat subtree_com cinert cdof crb qM qLD actuator_length actuator_velocity "
          "actuator_force qfrc_actuator cvel cdof_dot qfrc_passive qfrc_bias qfrc_smooth qacc_smooth efc_J efc_D efc_aref efc_pos efc_force "
          "qfrc_constraint qacc qacc_warmstart").split()
for name, (rel, key, kw) in MODELS.items():
    mj_model = mujoco.MjModel.from_xml_path(str(ROOT / rel if rel.startswith("tests/") else ROOT / "ambersim_b200/models" / rel))
    for k, v in kw.items():
        setattr(mj_model.opt, k, v)
    m = mjx.device_put(mj_model)
    rng = np.random.default_rng(42)
    q = mj_model.key_qpos[0] if key else mj_model.qpos0 + rng.uniform(0, 0.4, mj_model.nq)
    x0 = np.concatenate([q, 0.1 * rng.normal(size=mj_model.nv)])
    c0 = mj_model.key_ctrl[0] if key else np.zeros(mj_model.nu)
    us = np.clip(c0 + 0.2 * rng.normal(size=(20, mj_model.nu)), mj_model.actuator_ctrlrange[:, 0], mj_model.actuator_ctrlrange[:, 1])

    def shoot(x0, us):  # ambersim/trajopt/shooting.py:22-48
        d = mjx.make_data(m).replace(qpos=x0[: m.nq], qvel=x0[m.nq:])
        d = mjx.forward(m, d)

        def f(d, u):
            d = mjx.step(m, d.replace(ctrl=u))
            return d, jnp.concatenate((d.qpos, d.qvel))

        _, xs = jax.lax.scan(f, d, us)
        return jnp.concatenate((x0[None], xs))

    xs = np.asarray(jax.jit(shoot)(jnp.asarray(x0, jnp.float32), jnp.asarray(us, jnp.float32)))
    # one mjx.forward at a seeded state, every intermediate field (the stage-by-stage pin of the oracle)
    sq = q.copy()
    if key:
        sq[7:] += rng.uniform(-0.1, 0.1, mj_model.nq - 7)
        sq[2] -= 0.004  # feet pressed into the floor: contact rows active
    if name == "boxbot":
        sq[2] = 0.06  # the chassis box and the mesh foot on the tilted floor
    if name == "blocks":
        sq[2] = 0.31  # settled on the pedestal: capsule foot, sphere hand, mesh tail and (tilted) the chassis box
        sq[3:7] = np.array([0.99, -0.13, -0.04, 0.0]) / np.linalg.norm([0.99, -0.13, -0.04, 0.0])
    if name == "bh280_hulls":
        sq = np.array([2.34, 0.7, 1.5, 2.35, 0.7, 1.5, 1.4, 0.5])  # finger tips closed on each other
    sv, sc, sw = 0.3 * rng.normal(size=mj_model.nv), us[0], rng.normal(size=mj_model.nv)
    d = mjx.make_data(m).replace(qpos=jnp.asarray(sq, jnp.float32), qvel=jnp.asarray(sv, jnp.float32), ctrl=jnp.asarray(sc, jnp.float32),
                                 qacc_warmstart=jnp.asarray(sw, jnp.float32))
    d = jax.jit(mjx.forward)(m, d)
    stage = {f"stage_{f}": np.asarray(getattr(d, f)) for f in STAGES if hasattr(d, f)}
    if hasattr(d, "contact"):
        stage.update(stage_contact_dist=np.asarray(d.contact.dist), stage_contact_pos=np.asarray(d.contact.pos), stage_contact_frame=np.asarray(d.contact.frame))
    stage.update(stage_in_qpos=sq, stage_in_qvel=sv, stage_in_ctrl=sc, stage_in_qacc_warmstart=sw)
    # the compiled MjModel itself, field by field (MuJoCo's own layout): pins this repo's MJCF compiler (mj_setConst constants,
    # defaults, frames) and its mujoco.MjModel adapter through tests/test_model_io.py::test_mujoco_model_dump_if_present
    fields = ("body_parentid body_rootid body_weldid body_jntnum body_jntadr body_dofnum body_dofadr body_pos body_quat body_ipos body_iquat "
              "body_inertia body_invweight0 body_mass body_subtreemass jnt_type jnt_qposadr jnt_dofadr jnt_bodyid jnt_limited jnt_solref jnt_solimp "
              "jnt_pos jnt_axis jnt_range jnt_stiffness jnt_margin dof_bodyid dof_jntid dof_parentid dof_armature dof_damping dof_invweight0 "
              "dof_frictionloss qpos0 qpos_spring geom_type geom_bodyid geom_size geom_pos geom_quat geom_friction geom_solref geom_solimp "
              "geom_contype geom_conaffinity geom_condim geom_priority geom_solmix geom_margin geom_gap eq_type eq_obj1id eq_obj2id eq_active0 "
              "eq_solref eq_solimp eq_data actuator_trntype actuator_dyntype actuator_trnid actuator_gaintype actuator_biastype "
              "actuator_ctrllimited actuator_forcelimited actuator_ctrlrange actuator_forcerange actuator_gainprm actuator_biasprm actuator_gear "
              "exclude_signature key_qpos key_qvel key_ctrl").split()
    model = {f"model_{f}": np.asarray(getattr(mj_model, f)) for f in fields if hasattr(mj_model, f)}
    for f in ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "neq", "ntendon", "nmocap", "npair", "nkey"):
        model[f"model_{f}"] = int(getattr(mj_model, f))
    for f in ("timestep", "impratio", "tolerance", "ls_tolerance", "gravity", "integrator", "cone", "jacobian", "solver", "iterations", "ls_iterations", "disableflags"):
        model[f"model_opt_{f}"] = np.asarray(getattr(mj_model.opt, f))
    model["model_stat_meaninertia"] = float(mj_model.stat.meaninertia)
    np.savez_compressed(ROOT / "tests/golden" / f"mjx_{name}.npz", x0=x0, us=us, xs=xs, opt_timestep=mj_model.opt.timestep,
                        opt_iterations=mj_model.opt.iterations, opt_ls_iterations=mj_model.opt.ls_iterations,
                        opt_disableflags=mj_model.opt.disableflags, mujoco_version=mujoco.__version__, **model, **stage)
    print("wrote", name, xs.shape)
