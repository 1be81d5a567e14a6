"""Randomised models of the limb-kernel class (tests/_randmodel.py): host-side plan invariants on the CPU, and on the
GPU the limb kernels and the generic kernels against the float64 oracle, teacher-forced, from settled contact states.
Tolerances are the single-step bounds of tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from ambersim_b200 import mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file
from oracle.oracle import Oracle
from tests._randmodel import random_limb_model

CONFIGS = {"quad": (3, 1, 4), "long": (6, 4, 4), "wide": (3, 1, 8), "fastquad": (3, 1, 4), "hexa": (3, 1, 8), "legs2": (6, 4, 2), "flat4long": (6, 4, 4)}
# flat classes with the common options, served by compile-time sharing patterns and the fast variants of the limb kernels
FLAT = {"fastquad": dict(quad=True), "hexa": dict(quad=True, nlimb=6), "legs2": dict(quad=True, nlimb=2, leaf_contacts_only=True),
        "flat4long": dict(quad=True, nlimb=4, leaf_contacts_only=True)}


def _load(tmp_path, seed, cfg, **kw):
    if cfg in FLAT:
        kw = {**kw, **FLAT[cfg], "iterations": 1}
    xml, q, c = random_limb_model(seed, *CONFIGS[cfg], **kw)
    f = tmp_path / f"rand_{cfg}_{seed}.xml"
    f.write_text(xml)
    return load_mj_model_from_file(str(f)), q, c


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_random_model_plans_cover_every_body_once(tmp_path, cfg):
    eligible = 0
    for seed in range(12):
        mj, q, c = _load(tmp_path, seed, cfg)
        assert len(q) == mj.nq and len(c) == mj.nu
        p = mjx.limb_plan(mj)
        if not p["eligible"]:
            continue
        eligible += 1
        owned = {}
        for g, path in enumerate(p["paths"]):
            assert path[0] == 1 or path[0] == -1  # every carried path starts at the trunk (dummy lanes are all -1)
            for k, b in enumerate(path):
                if b > 0 and p["own"][g][k]:
                    assert b not in owned, f"body {b} owned twice (seed {seed})"
                    owned[b] = g
                if b > 0 and k > 0:
                    assert mj.body_parentid[b] == path[k - 1]
        assert sorted(owned) == list(range(1, mj.nbody)), f"seed {seed}: bodies without an owner lane"
        # lanes sharing a body agree on its sharing level, and the level is log2 of the number of sharing lanes
        for k in range(p["NL"] + 1):
            groups = {}
            for g, path in enumerate(p["paths"]):
                if path[k] > 0:
                    groups.setdefault(path[k], []).append(g)
            for b, lanes in groups.items():
                assert {p["level"][g][k] for g in lanes} == {int(np.ceil(np.log2(len(lanes))))} or len(lanes) == 1
    assert eligible >= 6
    if cfg in FLAT:
        want = {"fastquad": (2, 4), "hexa": (3, 8), "legs2": (1, 2), "flat4long": (2, 4)}[cfg]
        plans = [p for p in (mjx.limb_plan(_load(tmp_path, s, cfg)[0]) for s in range(12)) if p["eligible"]]
        assert eligible >= 10 and all((p["pattern"], p["lanes"]) == want for p in plans)  # a trunk contact on top of four feet overflows a lane


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_random_model_oracle_float32_tracks_float64(tmp_path, cfg):
    mj, q, c = _load(tmp_path, 3, cfg)
    o = Oracle(mj)
    x0 = np.concatenate([q, np.zeros(mj.nv)])[None]
    us = np.tile(c, (1, 60, 1))
    a, b = o.rollout(x0, us), o.rollout(x0, us, prec=1)
    assert np.isfinite(a).all() and np.abs(a - b).max() < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("seed", range(int(os.environ.get("ABR_SOAK_SEEDS", "10"))))  # a soak run sets ABR_SOAK_SEEDS higher
def test_random_model_step_parity(tmp_path, cfg, seed):
    import torch

    mj, q, c = _load(tmp_path, seed, cfg, iterations=1 + seed % 2)
    o = Oracle(mj)
    rng = np.random.default_rng(100 + seed)
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    clip = lambda u: np.where(mj.actuator_ctrllimited > 0, np.clip(u, lo, hi), u)
    # settle onto the floor under the home controls, then a teacher-forced stretch with noisy controls
    settle = o.rollout(np.concatenate([q, np.zeros(mj.nv)])[None], np.tile(clip(c), (1, 150, 1)))[0, -1]
    N = 24
    us = clip(c + 0.2 * rng.normal(size=(N, mj.nu)))
    qq, vv = settle[: mj.nq].copy(), settle[mj.nq:].copy()
    ww = o.forward(qq, vv)["qacc_warmstart"]
    qs, vs, ws, q1, v1, w1, q32, v32 = [], [], [], [], [], [], [], []
    for t in range(N):
        qs.append(qq); vs.append(vv); ws.append(ww)
        a32 = o.step(qq, vv, us[t], ww, prec=1)
        qq, vv, ww, _ = o.step(qq, vv, us[t], ww)
        q1.append(qq); v1.append(vv); w1.append(ww); q32.append(a32[0]); v32.append(a32[1])
    q1, v1, q32, v32 = (np.stack(a) for a in (q1, v1, q32, v32))
    t32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device="cuda")
    plan = mjx.limb_plan(mj)
    ran = []
    for lanes in ((1, 8) if plan["eligible"] else (8, 32)):
        m = mjx.device_put(mj)
        m.set_lanes(lanes)
        d = mjx.Data(qpos=t32(np.stack(qs)), qvel=t32(np.stack(vs)), ctrl=t32(us), qacc=torch.zeros(N, mj.nv, device="cuda"),
                     qacc_warmstart=t32(np.stack(ws)), time=torch.zeros(N, device="cuda"))
        d1 = mjx.step(m, d)
        gq, gv = d1.qpos.cpu().numpy(), d1.qvel.cpu().numpy()
        assert np.all(np.abs(gq - q1) <= 1e-5 + 1e-4 * np.abs(q1) + 3 * np.abs(q32 - q1)), f"qpos, lanes {lanes}"
        vmax = np.maximum(1.0, np.abs(v1).max(axis=1))
        assert np.all(np.abs(gv - v1).max(axis=1) <= 1e-4 * vmax + 3 * np.abs(v32 - v1).max(axis=1)), f"qvel, lanes {lanes}"
        ran.append(gq)
    if plan["eligible"]:  # limb vs generic kernels: different arithmetic, so different roundings (two group sizes of the generic kernels may well agree bit for bit)
        assert not np.array_equal(ran[0], ran[1])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("seed", range(max(2, int(os.environ.get("ABR_SOAK_SEEDS", "10")) // 5)))
def test_random_model_sampler_and_env_parity(tmp_path, cfg, seed):
    """The other two entry points on random models: a predictive-sampling solve (per-sample costs against oracle rollouts, first-minimum
    argmin on the device's own costs, winner trajectory = rollout of the winner controls) and env steps with substeps and the
    auto-reset blend against the oracle."""
    import torch

    from ambersim_b200.rl.base import VectorEnvStepper
    from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
    from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams, shoot
    from oracle.oracle import quad_cost

    mj, q, c = _load(tmp_path, 1000 + seed, cfg, iterations=1)
    if mj.nu == 0:
        pytest.skip("a model without actuators has nothing to sample")
    o = Oracle(mj)
    m = mjx.device_put(mj)
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    clip = lambda u: np.where(mj.actuator_ctrllimited > 0, np.clip(u, lo, hi), u)
    settle = o.rollout(np.concatenate([q, np.zeros(mj.nv)])[None], np.tile(clip(c), (1, 150, 1)))[0, -1]
    nx, S, N = mj.nq + mj.nv, 24, 6
    t32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device="cuda")
    Q, Qf, R = np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu)
    ps = VanillaPredictiveSampler(model=m, cost_function=StaticGoalQuadraticCost(Q, Qf, R, settle), nsamples=S, stdev=0.1)
    rng = np.random.default_rng(500 + seed)
    noise = rng.normal(size=(S - 1, N, mj.nu)).astype(np.float32)
    ug = np.tile(clip(c), (N, 1)).astype(np.float32)
    x0 = settle.astype(np.float32)
    xs, us, info = ps.optimize(VanillaPredictiveSamplerParams(key=0, x0=t32(x0), us_guess=t32(ug), noise=t32(noise)), return_info=True)
    costs = info["costs"].reshape(-1).cpu().numpy()
    us_all = np.concatenate([ug[None], ug[None] + np.float32(0.1) * noise]).astype(np.float64)
    if (mj.actuator_ctrllimited > 0).all():
        us_all = np.clip(us_all, lo, hi)
        ref = quad_cost(o.rollout(x0.astype(np.float64), us_all), us_all, Q, Qf, R, settle)
        assert np.allclose(costs, ref, rtol=1e-2, atol=1e-3), np.abs(costs / ref - 1).max()
    k = int(info["best_idx"])
    assert np.isfinite(costs).all() and k == int(np.argmin(costs))
    # the winner's trajectory is the rollout of the winner's controls (another compile-time variant of the kernel ran it: equal up to rounding)
    assert (xs - shoot(m, t32(x0), us)).abs().max() < 1e-4
    # env steps: two substeps, half of the envs reset to their first state before stepping
    E = 6
    q0s = np.tile(settle[: mj.nq], (E, 1))
    st = VectorEnvStepper(m, t32(q0s), torch.zeros(E, mj.nv, device="cuda"), nsubsteps=2)
    ctrl = clip(c + 0.2 * rng.normal(size=(E, mj.nu)))
    st.step(t32(ctrl))
    mid_q, mid_v, mid_w = st.qpos.cpu().numpy().copy(), st.qvel.cpu().numpy().copy(), st.warm.cpu().numpy().copy()
    done = torch.tensor([1, 0, 1, 0, 1, 0], device="cuda")
    st.step(t32(ctrl), done)
    fq, fv, fw = st.first_qpos.cpu().numpy(), st.first_qvel.cpu().numpy(), st.first_warm.cpu().numpy()
    for e in range(E):
        a = (fq[e], fv[e], fw[e]) if int(done[e]) else (mid_q[e], mid_v[e], mid_w[e])
        qr, vr, _, _ = o.step(a[0], a[1], ctrl[e], a[2], nsteps=2)
        q32, v32, _, _ = o.step(a[0], a[1], ctrl[e], a[2], nsteps=2, prec=1)
        assert np.all(np.abs(st.qpos[e].cpu().numpy() - qr) <= 2e-5 + 2e-4 * np.abs(qr) + 3 * np.abs(q32 - qr)), e
        assert np.abs(st.qvel[e].cpu().numpy() - vr).max() <= 2e-4 * max(1.0, np.abs(vr).max()) + 3 * np.abs(v32 - vr).max(), e


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("seed", range(max(2, int(os.environ.get("ABR_SOAK_SEEDS", "10")) // 5)))
def test_random_model_derived_fields_and_host_calls(tmp_path, cfg, seed):
    """Batched derived mjx.Data fields (xpos, xquat, cvel, qfrc_*, efc_*, contact_*) of one forward pass on random models against the
    oracle's stage dump, and the host-pointer entry points (numpy in, numpy out) against the device-pointer ones."""
    import torch

    from ambersim_b200.trajopt.shooting import shoot

    mj, q, c = _load(tmp_path, 2000 + seed, cfg, iterations=2)
    o = Oracle(mj)
    m = mjx.device_put(mj)
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    clip = lambda u: np.where(mj.actuator_ctrllimited > 0, np.clip(u, lo, hi), u) if mj.nu else u
    settle = o.rollout(np.concatenate([q, np.zeros(mj.nv)])[None], np.tile(clip(c), (1, 120, 1)))[0, -1]
    rng = np.random.default_rng(700 + seed)
    E = 3
    qs = np.tile(settle[: mj.nq], (E, 1))
    qs[:, 7:] += rng.uniform(-0.05, 0.05, (E, mj.nq - 7))
    vs = 0.3 * rng.normal(size=(E, mj.nv))
    cs = clip(c + 0.2 * rng.normal(size=(E, mj.nu)))
    ws = rng.normal(size=(E, mj.nv))
    t32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device="cuda")
    names = ("xpos", "xquat", "cvel", "qfrc_smooth", "qacc_smooth", "qfrc_constraint", "efc_force", "efc_D", "efc_aref", "contact_dist", "contact_pos")
    d = mjx.forward(m, mjx.Data(qpos=t32(qs), qvel=t32(vs), ctrl=t32(cs), qacc=torch.zeros(E, mj.nv, device="cuda"), qacc_warmstart=t32(ws),
                                time=torch.zeros(E, device="cuda")), fields=names)
    for e in range(E):
        ref = o.forward(d.qpos[e].cpu().numpy(), vs[e], cs[e], ws[e])
        for n in names:
            r, g = ref[n].ravel(), getattr(d, n)[e].cpu().numpy().ravel()
            tol = 5e-3 if n in ("efc_force", "qfrc_constraint") else 5e-4  # two Newton iterations, not converged: the forces carry the solver's float32 path
            assert np.abs(r - g).max() <= tol * max(1e-6, np.abs(r).max()) + 1e-5, (n, e, np.abs(r - g).max(), np.abs(r).max())
    # host-pointer rollout = device-pointer rollout, bit for bit
    N = 9
    x0 = np.concatenate([qs[0], vs[0]]).astype(np.float32)
    us = clip(c + 0.1 * rng.normal(size=(4, N, mj.nu))).astype(np.float32)
    xs_h = shoot(m, x0, us)
    xs_d = shoot(m, t32(x0), t32(us)).cpu().numpy()
    assert isinstance(xs_h, np.ndarray) and np.array_equal(xs_h, xs_d)
