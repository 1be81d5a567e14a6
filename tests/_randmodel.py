"""Random floating-base limb models (MJCF text) for the parity tests: a trunk on a free joint with 2-8 leaf paths,
chains of random length, optional forks (bodies shared by several paths), hinge / slide joints with random axes,
anchors and body-frame rotations, random inertias with tilted principal axes, damping, armature, springs, limits, position / motor actuators
with control and force ranges (some joints unactuated), sphere feet with condim 1 or 3 and random friction on
random bodies of a path, and a tilted floor. Everything a model of the limb-kernel class can contain
(abr_limb.cuh eligibility); the generic kernels and the oracle take the same file."""
from __future__ import annotations

import numpy as np


def _f(a):
    return " ".join(f"{x:.5g}" for x in np.atleast_1d(a))


def random_limb_model(seed: int, max_chain: int = 3, max_con: int = 1, max_leaves: int = 4, iterations: int = 1, forks: bool = True, quad: bool = False,
                      nlimb: int = 0, leaf_contacts_only: bool = False):
    """Returns (xml, home_qpos, home_ctrl). Chains have <= max_chain joints between trunk and leaf and at most
    max_con foot spheres per root-to-leaf path (the capacity of the compiled limb kernels: (3,1) and (6,4)).
    quad: the flat four-limb class with the common options (hinges only, no forks, eulerdamp disabled), i.e. a model the
    compile-time fast variants of the limb kernels serve. nlimb > 0 with quad=True gives the other flat classes (2 = legs only, 6 = hexapod);
    leaf_contacts_only keeps every foot on a leaf body (the contact-body form of the long-chain kernels)."""
    rng = np.random.default_rng(seed)
    acts, qpos, ctrl = [], [], []
    counter, leaves = [0], [0]

    def body(pos, remaining, budget_con, prefix, may_fork):
        """One chain body at `pos` in its parent's frame, with `remaining` more bodies below it on this path."""
        k = counter[0]
        counter[0] += 1
        name = f"{prefix}{k}"
        slide = (not quad) and rng.random() < 0.2
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        limited = rng.random() < 0.7
        rng_lo, rng_hi = (-0.06, 0.06) if slide else (-float(rng.uniform(0.5, 1.4)), float(rng.uniform(0.5, 1.4)))
        q0 = float(rng.uniform(0.5 * rng_lo, 0.5 * rng_hi))
        jattr = (f'name="{name}" axis="{_f(axis)}" pos="{_f(rng.uniform(-0.02, 0.02, 3))}" damping="{rng.uniform(0.1, 0.8):.4g}" '
                 f'armature="{rng.uniform(0.005, 0.03):.4g}"')
        if slide:
            jattr += ' type="slide"'
        if limited:
            jattr += f' range="{rng_lo:.4g} {rng_hi:.4g}"'
        if rng.random() < 0.3:
            jattr += f' stiffness="{rng.uniform(5, 40):.4g}" springref="{rng.uniform(-0.1, 0.1):.3g}"'
        mass = float(rng.uniform(0.15, 0.7))
        inertia = mass * rng.uniform(0.002, 0.012, 3)
        length = float(rng.uniform(0.08, 0.16))
        xml = [f'<body name="{name}" pos="{_f(pos)}" euler="{_f(rng.uniform(-0.3, 0.3, 3))}">',
               f'<inertial pos="{_f(rng.uniform(-0.02, 0.02, 3))}" mass="{mass:.4g}" diaginertia="{_f(inertia)}"'
               + (f' euler="{_f(rng.uniform(-0.8, 0.8, 3))}"' if rng.random() < 0.5 else "") + "/>",  # tilted principal axes on half of the bodies
               f"<joint {jattr}/>"]
        qpos.append(q0)
        r = rng.random()
        if r < 0.6:
            kp = float(rng.uniform(150, 400) if slide else rng.uniform(15, 45))
            a = f'<position name="{name}" joint="{name}" kp="{kp:.4g}"'
            if limited:
                a += f' ctrlrange="{rng_lo:.4g} {rng_hi:.4g}"'
            if rng.random() < 0.4:
                fr = float(rng.uniform(4, 12))
                a += f' forcerange="{-fr:.4g} {fr:.4g}"'
            acts.append(a + "/>")
            ctrl.append(q0)
        elif r < 0.8:
            acts.append(f'<motor name="{name}" joint="{name}" ctrlrange="-3 3" gear="{rng.uniform(0.5, 2):.3g}"/>')
            ctrl.append(0.0)
        # feet: a leaf gets them while the path still has budget; inner bodies sometimes
        ncon_here = 0
        if budget_con > 0 and (remaining == 0 or (not leaf_contacts_only and rng.random() < 0.25)):
            ncon_here = int(rng.integers(1, (min(budget_con, 2) if remaining else budget_con) + 1))
            for c in range(ncon_here):
                cd = 3 if rng.random() < 0.75 else 1
                xml.append(f'<geom name="{name}_f{c}" class="foot" size="{rng.uniform(0.02, 0.035):.3g}" '
                           f'pos="{_f(rng.uniform(-0.03, 0.03, 2))} {-0.5 * length:.4g}" condim="{cd}" friction="{rng.uniform(0.4, 1.2):.3g} 0.02 0.01"/>')
        if remaining > 0:
            nchild = 2 if (may_fork and leaves[0] < max_leaves and rng.random() < 0.3) else 1
            leaves[0] += nchild - 1
            for ch in range(nchild):
                xml.append(body(np.concatenate([rng.uniform(-0.03, 0.03, 2) + (0.05 * (2 * ch - 1) if nchild > 1 else 0), [-length]]),
                                remaining - 1, budget_con - ncon_here, prefix, may_fork))
        xml.append("</body>")
        return "\n".join(xml)

    limbs = []
    nlimb = nlimb if nlimb > 0 else (4 if quad else int(rng.integers(2, min(4, max_leaves) + 1)))
    leaves[0] = nlimb
    for limb in range(nlimb):
        n = int(rng.integers(1, max_chain + 1))
        ang = 2 * np.pi * limb / nlimb + rng.uniform(-0.2, 0.2) / max(1, nlimb // 4)
        limbs.append(body(np.array([0.18 * np.cos(ang), 0.18 * np.sin(ang), -0.03]), n - 1, max_con, f"l{limb}_", forks and not quad))
    height = 0.45
    trunk_mass = float(rng.uniform(2.0, 5.0))
    tilt = rng.uniform(-0.05, 0.05, 2)
    belly = '<geom name="belly" class="foot" size="0.05" pos="0 0 -0.1" condim="1"/>' if rng.random() < 0.3 else ""
    xml = f"""<mujoco model="random_limb_{seed}">
  <compiler angle="radian" autolimits="true"/>
  <option timestep="0.003" iterations="{iterations}" ls_iterations="6" integrator="Euler" solver="Newton"><flag eulerdamp="{'enable' if (not quad and rng.random() < 0.5) else 'disable'}"/></option>
  <default>
    <geom contype="0" conaffinity="0"/>
    <default class="foot"><geom type="sphere" contype="1" conaffinity="0"/></default>
  </default>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 0.05" euler="{_f(tilt)} 0" contype="0" conaffinity="1" condim="3"/>
    <body name="base" pos="0 0 {height}">
      <inertial pos="{_f(rng.uniform(-0.02, 0.02, 3))}" mass="{trunk_mass:.4g}" diaginertia="{_f(trunk_mass * rng.uniform(0.01, 0.03, 3))}" euler="{_f(rng.uniform(-0.5, 0.5, 3))}"/>
      <freejoint name="base"/>
      {belly}
      {chr(10).join(limbs)}
    </body>
  </worldbody>
  <actuator>
    {chr(10).join(acts)}
  </actuator>
</mujoco>
"""
    home_q = np.concatenate([[0, 0, height, 1, 0, 0, 0], qpos])
    return xml, home_q, np.asarray(ctrl, dtype=np.float64)
