"""Run on a machine that HAS mujoco + mujoco-mjx + jax (not this image): dumps reference rollouts of
the repo's models into tests/golden/mjx_<model>.npz so that tests/test_oracle_physics.py
(test_mjx_golden_if_present) can pin the oracle against real MJX outputs.

    python tools/dump_mjx_golden.py
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
          "barkour": ("barkour_standin/barkour_vb_standin.xml", "home", {}), "biped": ("biped_standin/biped_exo_standin.xml", "stand", {})}
for name, (rel, key, kw) in MODELS.items():
    mj_model = mujoco.MjModel.from_xml_path(str(ROOT / "ambersim_b200/models" / rel))
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
    np.savez_compressed(ROOT / "tests/golden" / f"mjx_{name}.npz", x0=x0, us=us, xs=xs, opt_timestep=mj_model.opt.timestep,
                        opt_iterations=mj_model.opt.iterations, opt_ls_iterations=mj_model.opt.ls_iterations,
                        opt_disableflags=mj_model.opt.disableflags, mujoco_version=mujoco.__version__)
    print("wrote", name, xs.shape)
