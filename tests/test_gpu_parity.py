"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on the same
seeded inputs (sizes the oracle finishes in seconds), plus size-independent properties at
BASELINE.json's full sizes. Tolerances (float32 kernel vs float64 oracle):
  * one forward/step from an identical state: |d qpos| <= 1e-5 + 1e-4 |ref| elementwise,
    |d qvel| <= 1e-4 max(1, max|qvel_ref|) (velocities are dt x accelerations of magnitude 1e2-1e3, so
    the bound is vector-relative), 2e-4 relative on internal stages;
  * contact-free rollouts (N <= 32): 1e-3 absolute on qpos/qvel, cost rtol 1e-3; where the
    float32 build of the oracle itself drifts further than that from float64 (the Barrett hand's
    1e-5 kg m^2 finger inertias under unit torques), the bound is 3x that float32-vs-float64 drift;
  * rollouts with contacts: teacher-forced per-step comparison, cost rtol 1e-2;
  * argmin: bit-exact against numpy.argmin of the device's own costs, and equal to the oracle's
    index whenever the oracle's best-vs-second gap exceeds the cost tolerance.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from ambersim_b200 import _lib, mjx
from ambersim_b200.rl.base import VectorEnvStepper
from ambersim_b200.rl.pendulum.swingup import PendulumSwingupEnv
from ambersim_b200.trajopt.base import CostFunction, CostFunctionParams
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost
from ambersim_b200.trajopt.shooting import VanillaPredictiveSampler, VanillaPredictiveSamplerParams, _seed_of, shoot, shoot_cost
from oracle.oracle import Oracle, quad_cost
from tests import _philox

pytestmark = pytest.mark.gpu
DEV = "cuda"
STAGES = ["xpos", "xquat", "xipos", "cinert", "cdof", "qM", "cvel", "cdof_dot", "contact_dist", "contact_pos", "contact_frame",
          "qfrc_smooth", "qacc_smooth", "efc_J", "efc_D", "efc_aref", "qacc", "efc_force", "qfrc_constraint"]
MODEL_KEY = {"barkour": "home", "biped": "stand", "exolegs": "stand", "tripod": "home", "tripod3": "home"}
BH_OPT = dict(timestep=0.002, iterations=1, ls_iterations=4, integrator=0, solver=2, disableflags=16)  # the reference test's options


def t32(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32, device=DEV)


def sample_state(mj, name, rng):
    key = MODEL_KEY.get(name)
    q = mj.key_qpos(key) if key else mj.qpos0.copy()
    if name == "bh280":
        q = q + rng.uniform(0.0, 0.5, mj.nq)
    elif name == "pendulum":
        q = q + rng.uniform(-2, 2, 1)
    else:
        q[7:] += rng.uniform(-0.1, 0.1, mj.nq - 7)
        q[2] -= 0.004
    v = rng.normal(size=mj.nv) * 0.3
    c = (mj.key_ctrl(key) if key else np.zeros(mj.nu)) + rng.normal(size=mj.nu) * 0.1
    return q, v, c


def model_with(load_model, name, **opt):
    mj = load_model(name)
    if name == "bh280":
        opt = {**BH_OPT, **opt}
    m = mjx.device_put(mj)
    m = m.replace(opt=m.opt.replace(**opt))
    return mj, m, Oracle(mj, m.opt)


@pytest.mark.parametrize("name", ["pendulum", "bh280", "barkour", "biped"])
def test_forward_stage_parity(load_model, name):
    mj, m, o = model_with(load_model, name)
    rng = np.random.default_rng(0)
    q, v, c = sample_state(mj, name, rng)
    w = rng.normal(size=mj.nv)
    ref = o.forward(q, v, c, w)
    got = mjx.debug_forward(m, q, v, c, w, names=STAGES)
    for f in STAGES:
        r, g = ref[f].ravel(), got[f].ravel()
        assert r.size == g.size, f
        if r.size:
            assert np.abs(r - g).max() <= 2e-4 * max(1e-6, np.abs(r).max()), f


@pytest.mark.parametrize("name", ["pendulum", "bh280", "barkour", "biped", "exolegs", "tripod", "tripod3"])
@pytest.mark.parametrize("variant", ["default", "rk4", "cg", "eulerdamp", "converged", "nowarm"])
def test_single_step_parity_option_variants(load_model, name, variant):
    opt = dict(default={}, rk4=dict(integrator=1), cg=dict(solver=1, iterations=8, ls_iterations=10),
               eulerdamp=dict(disableflags=0), converged=dict(iterations=30, ls_iterations=30), nowarm=dict(disableflags=16384 | 256))[variant]
    if name == "bh280" and "disableflags" in opt:
        opt = dict(disableflags=opt["disableflags"] | 16)
    mj, m, o = model_with(load_model, name, **opt)
    rng = np.random.default_rng(11)
    E = 6
    states = [sample_state(mj, name, rng) for _ in range(E)]
    q = np.stack([s[0] for s in states]); v = np.stack([s[1] for s in states]); c = np.stack([s[2] for s in states])
    w = rng.normal(size=(E, mj.nv))
    d = mjx.Data(qpos=t32(q), qvel=t32(v), ctrl=t32(c), qacc=torch.zeros(E, mj.nv, device=DEV), qacc_warmstart=t32(w),
                 time=torch.zeros(E, device=DEV))
    d1 = mjx.step(m, d)
    for e in range(E):
        qr, vr, wr, tr = o.step(q[e], v[e], c[e], w[e])
        q32, v32, w32, _ = o.step(q[e], v[e], c[e], w[e], prec=1)  # the same algorithm in float32 on the CPU
        assert np.all(np.abs(d1.qpos[e].cpu().numpy() - qr) <= 1e-5 + 1e-4 * np.abs(qr) + 3 * np.abs(q32 - qr))
        assert np.abs(d1.qvel[e].cpu().numpy() - vr).max() <= 1e-4 * max(1.0, np.abs(vr).max()) + 3 * np.abs(v32 - vr).max()
        assert np.abs(d1.qacc_warmstart[e].cpu().numpy() - wr).max() <= 1e-3 * max(1.0, np.abs(wr).max()) + 3 * np.abs(w32 - wr).max()
        assert abs(float(d1.time[e]) - tr) < 1e-6


@pytest.mark.parametrize("name", ["pendulum", "bh280"])
@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_contact_free_rollout_parity(load_model, name, lanes):
    mj, m, o = model_with(load_model, name)
    m.set_lanes(lanes)
    rng = np.random.default_rng(2)
    W, N = 5, 32
    x0 = np.stack([np.concatenate(sample_state(mj, name, rng)[:2]) for _ in range(W)])
    us = rng.normal(size=(W, N, mj.nu)) * 0.5
    xs = shoot(m, t32(x0), t32(us)).cpu().numpy()
    ref = o.rollout(x0, us)
    assert xs.shape == (W, N + 1, mj.nq + mj.nv)
    assert np.array_equal(xs[:, 0], x0.astype(np.float32))  # row 0 is the caller's x0 verbatim
    drift32 = np.abs(o.rollout(x0, us, prec=1) - ref).max()
    assert np.abs(xs - ref).max() < max(1e-3, 3 * drift32)
    eye = np.eye(mj.nq + mj.nv)
    cf = StaticGoalQuadraticCost(eye, 10 * eye, 0.01 * np.eye(mj.nu), np.zeros(mj.nq + mj.nv))
    costs = shoot_cost(m, t32(x0), t32(us), cf).cpu().numpy()
    assert np.allclose(costs, quad_cost(ref, us, eye, 10 * eye, 0.01 * np.eye(mj.nu), 0.0), rtol=1e-3)


@pytest.mark.parametrize("name", ["barkour", "biped", "exolegs", "tripod", "tripod3"])
@pytest.mark.parametrize("lanes", [1, 4, 8, 16, 32])  # 1 = the limb (path-decomposed) kernels, 4..32 = generic group sizes
def test_contact_rollout_teacher_forced(load_model, name, lanes):
    mj, m, o = model_with(load_model, name)
    m.set_lanes(lanes)
    rng = np.random.default_rng(4)
    N = 40
    q, v, c = sample_state(mj, name, rng)
    x0 = np.concatenate([q, v])
    us = np.clip(c + 0.1 * rng.normal(size=(N, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
    ref = o.rollout(x0, us)
    # free-running comparison (chaotic once contacts switch, so a loose bound) ...
    xs = shoot(m, t32(x0), t32(us)).cpu().numpy()
    assert np.abs(xs - ref).max() < 5e-3
    # ... and teacher-forced: every step restarted from the oracle's state, with the oracle's warm start
    qs, vs, ws = [], [], []
    qq, vv, ww = x0[: mj.nq].copy(), x0[mj.nq:].copy(), o.forward(x0[: mj.nq], x0[mj.nq:])["qacc_warmstart"]
    for t in range(N):
        qs.append(qq); vs.append(vv); ws.append(ww)
        qq, vv, ww, _ = o.step(qq, vv, us[t], ww)
    d = mjx.Data(qpos=t32(np.stack(qs)), qvel=t32(np.stack(vs)), ctrl=t32(us), qacc=torch.zeros(N, mj.nv, device=DEV),
                 qacc_warmstart=t32(np.stack(ws)), time=torch.zeros(N, device=DEV))
    d1 = mjx.step(m, d)
    assert np.all(np.abs(d1.qpos.cpu().numpy() - ref[1:, : mj.nq]) <= 1e-5 + 1e-4 * np.abs(ref[1:, : mj.nq]))
    assert np.all(np.abs(d1.qvel.cpu().numpy() - ref[1:, mj.nq:]).max(axis=1) <= 1e-4 * np.maximum(1.0, np.abs(ref[1:, mj.nq:]).max(axis=1)))
    eye = np.eye(mj.nq + mj.nv)
    cf = StaticGoalQuadraticCost(eye, 10 * eye, 0.01 * np.eye(mj.nu), x0)
    cost = float(shoot_cost(m, t32(x0), t32(us), cf).cpu())
    assert np.isclose(cost, quad_cost(ref, us, eye, 10 * eye, 0.01 * np.eye(mj.nu), x0), rtol=1e-2)


def test_shoot_shapes_and_host_api(load_model):
    """x0 (nx,) with us (S,N,nu) = vmap(shoot,(None,None,0)); x0 (B,nx) with us (B,N,nu) =
    vmap(shoot,(None,0,0)) (tests/trajopt/test_predictive_sampler.py:85); numpy in -> numpy out."""
    mj, m, o = model_with(load_model, "bh280")
    rng = np.random.default_rng(5)
    x0 = rng.normal(size=16) * 0.2
    us = rng.normal(size=(7, 10, 4))
    a = shoot(m, t32(x0), t32(us))
    b = shoot(m, t32(np.tile(x0, (7, 1))), t32(us))
    assert a.shape == (7, 11, 16) and torch.equal(a, b)
    single = shoot(m, t32(x0), t32(us[3]))
    assert single.shape == (11, 16) and torch.equal(single, a[3])
    host = shoot(m, x0.astype(np.float32), us.astype(np.float32))
    assert isinstance(host, np.ndarray) and np.array_equal(host, a.cpu().numpy())
    empty = shoot(m, t32(x0), t32(np.zeros((0, 4))))  # N = 0: just x0
    assert empty.shape == (1, 16) and np.array_equal(empty.cpu().numpy()[0], x0.astype(np.float32))


# ------------------------------------------------------------------ predictive sampler
@pytest.fixture
def vps_data(load_model):
    """The reference fixture (tests/trajopt/test_predictive_sampler.py:17-41): bh280, contacts off,
    Q = I, Qf = 10 I, R = 0.01 I, xg = 0, 100 samples, stdev 0.01."""
    mj, model, o = model_with(load_model, "bh280")
    nx = model.nq + model.nv
    cost_function = StaticGoalQuadraticCost(Q=torch.eye(nx), Qf=10.0 * torch.eye(nx), R=0.01 * torch.eye(model.nu), xg=torch.zeros(nx))
    ps = VanillaPredictiveSampler(model=model, cost_function=cost_function, nsamples=100, stdev=0.01)
    return ps, model, cost_function, o


def test_smoke_VPS(vps_data):
    ps, model, _, _ = vps_data
    params = VanillaPredictiveSamplerParams(key=0, x0=torch.zeros(model.nq + model.nv, device=DEV),
                                            us_guess=torch.zeros((10, model.nu), device=DEV))
    xs_star, us_star = ps.optimize(params)
    assert xs_star.shape == (11, 16) and us_star.shape == (10, 4)
    assert torch.isfinite(xs_star).all() and torch.isfinite(us_star).all()


def test_VPS_cost_decrease(vps_data):
    """The reference's test verbatim (tests/trajopt/test_predictive_sampler.py:63-76): x0 ~ N(0,1), us_guess ~ N(0,1), the
    winner is no worse than the guess. torch's stream for this seed holds one hand pose with joints 1.2 - 2.1 rad outside their
    limits; under the fixture's single Newton iteration that start DIVERGES in any precision (the float64 oracle passes
    |x| = 1e20 within the 10 steps, the float32 oracle - MJX's own precision - overflows to NaN on the same samples:
    profiles/r2_vps_n01_divergence.txt), and a NaN cost is the argmin by numpy's semantics (shooting.py:154). So: wherever
    every sample's float32 cost is finite the reference's inequality holds as written; a problem with non-finite sample costs
    must be one whose offending sample diverges in the float64 oracle too (algorithmic, not a robustness gap of the kernel)."""
    ps, model, cost_function, o = vps_data
    g = torch.Generator(device=DEV).manual_seed(0)
    B, N = 10, 10
    x0 = torch.randn((B, model.nq + model.nv), generator=g, device=DEV)
    us_guess = torch.randn((B, N, model.nu), generator=g, device=DEV)
    key = torch.tensor([0, 7])
    xs_stars, us_stars, info = ps.optimize(VanillaPredictiveSamplerParams(key=key, x0=x0, us_guess=us_guess), return_info=True)
    costs_star, _ = cost_function.cost(xs_stars, us_stars, CostFunctionParams())
    xs_guess = shoot(model, x0, us_guess)
    costs_guess, _ = cost_function.cost(xs_guess, us_guess, CostFunctionParams())
    ok = torch.isfinite(info["costs"]).all(dim=1)
    assert int(ok.sum()) >= B - 2
    assert torch.all(costs_star[ok] <= costs_guess[ok]), (costs_star, costs_guess)
    from ambersim_b200.trajopt.shooting import _seed_of

    lim = model.actuator_ctrlrange.astype(np.float32)
    for b in torch.nonzero(~ok).flatten().tolist():
        bad = int(torch.nonzero(~torch.isfinite(info["costs"][b])).flatten()[0])
        z = _philox.normals(_seed_of(key), bad, b, np.arange(N * model.nu)).reshape(N, model.nu)
        u = np.clip(us_guess[b].cpu().numpy() + z * np.float32(ps.stdev), lim[:, 0], lim[:, 1])
        xs64 = o.rollout(x0[b].cpu().numpy().astype(np.float64), u.astype(np.float64))
        assert (not np.isfinite(xs64).all()) or np.abs(xs64).max() > 1e12, "float32 cost not finite where the float64 oracle does not diverge"


def test_VPS_parity_mode_against_oracle(vps_data):
    """Caller-supplied normals: the engine's costs match the oracle's, the device argmin is the numpy
    argmin of its own costs bit-exactly, and equals the oracle's wherever the gap is resolvable."""
    ps, model, cf, o = vps_data
    rng = np.random.default_rng(7)
    B, S, N = 3, 100, 10
    x0 = rng.normal(size=(B, 16)) * 0.3
    ug = rng.normal(size=(B, N, 4))
    noise = rng.normal(size=(B, S - 1, N, 4))
    xs, us, info = ps.optimize(VanillaPredictiveSamplerParams(key=0, x0=t32(x0), us_guess=t32(ug), noise=t32(noise)), return_info=True)
    costs = info["costs"].cpu().numpy()
    assert np.array_equal(info["best_idx"].cpu().numpy(), np.argmin(costs, axis=1))
    lim = model.actuator_ctrlrange
    for b in range(B):
        us_all = np.clip(ug[b][None] + np.concatenate([np.zeros((1, N, 4)), noise[b] * 0.01]), lim[:, 0], lim[:, 1])
        xs_all = o.rollout(np.tile(x0[b], (S, 1)), us_all)
        ref = quad_cost(xs_all, us_all, np.eye(16), 10 * np.eye(16), 0.01 * np.eye(4), 0.0)
        assert np.allclose(costs[b], ref, rtol=1e-3)
        order = np.sort(ref)
        if order[1] - order[0] > 2e-3 * abs(order[0]):
            assert int(info["best_idx"][b]) == int(np.argmin(ref))
        k = int(info["best_idx"][b])
        assert np.allclose(us[b].cpu().numpy(), us_all[k], atol=1e-6)
        drift32 = np.abs(o.rollout(x0[b], us_all[k], prec=1) - xs_all[k]).max()
        assert np.abs(xs[b].cpu().numpy() - xs_all[k]).max() < max(1e-3, 3 * drift32)
        assert np.isclose(float(info["best_cost"][b]), costs[b, k])


def test_VPS_device_noise_matches_numpy_philox_and_sharding(vps_data):
    ps, model, cf, o = vps_data
    rng = np.random.default_rng(8)
    B, S, N, seed = 2, 64, 6, 0x1234ABCD5678
    x0, ug = t32(rng.normal(size=(B, 16)) * 0.2), t32(rng.normal(size=(B, N, 4)))
    ps = VanillaPredictiveSampler(model=model, cost_function=cf, nsamples=S, stdev=0.5)
    p = VanillaPredictiveSamplerParams(key=seed, x0=x0, us_guess=ug)
    xs, us, info = ps.optimize(p, return_info=True)
    idx = info["best_idx"].cpu().numpy()
    lim = model.actuator_ctrlrange.astype(np.float32)
    for b in range(B):
        z = _philox.normals(seed, idx[b], b, np.arange(N * 4)).reshape(N, 4) if idx[b] > 0 else np.zeros((N, 4), np.float32)
        expect = np.clip(ug[b].cpu().numpy() + z * np.float32(0.5), lim[:, 0], lim[:, 1])
        assert np.allclose(us[b].cpu().numpy(), expect, atol=2e-6)
    # sample 0 is the un-noised guess
    ps1 = VanillaPredictiveSampler(model=model, cost_function=cf, nsamples=1, stdev=0.5)
    _, us1 = ps1.optimize(p)
    assert torch.equal(us1, torch.clamp(ug, -30, 30))
    # sharding by global sample id: the union over shards reproduces the single call bit-exactly
    parts = []
    for lo, hi in ((0, 20), (20, 48), (48, 64)):
        psk = VanillaPredictiveSampler(model=model, cost_function=cf, nsamples=hi - lo, stdev=0.5)
        parts.append(psk.optimize(p, sample_offset=lo, nsamples_total=S, return_info=True)[2])
    assert torch.equal(torch.cat([q["costs"] for q in parts], dim=1), info["costs"])
    best = torch.stack([q["best_cost"] for q in parts])
    gid = torch.stack([q["best_idx"] for q in parts])
    pick = best.argmin(dim=0)
    assert torch.equal(gid[pick, torch.arange(B, device=DEV)], info["best_idx"])
    # host-pointer entry point gives the same answer
    xs_h, us_h = ps.optimize(VanillaPredictiveSamplerParams(key=seed, x0=x0.cpu().numpy(), us_guess=ug.cpu().numpy()))
    assert np.array_equal(xs_h, xs.cpu().numpy()) and np.array_equal(us_h, us.cpu().numpy())


def test_argmin_nan_counts_as_minimum(vps_data):
    ps, model, cf, _ = vps_data
    x0 = torch.zeros((2, 16), device=DEV)
    x0[1, 3] = float("nan")
    _, _, info = ps.optimize(VanillaPredictiveSamplerParams(key=1, x0=x0, us_guess=torch.zeros((2, 5, 4), device=DEV)), return_info=True)
    assert int(info["best_idx"][1]) == 0 and torch.isnan(info["best_cost"][1])  # all NaN -> first index
    assert torch.isfinite(info["best_cost"][0])


def test_generic_cost_function_path(vps_data):
    ps, model, cf, _ = vps_data

    class Mine(CostFunction):
        def cost(self, xs, us, params):
            return cf.cost(xs, us, params)

    g = torch.Generator(device=DEV).manual_seed(3)
    x0, ug = torch.randn(16, generator=g, device=DEV) * 0.2, torch.randn((8, 4), generator=g, device=DEV)
    nz = torch.randn((99, 8, 4), generator=g, device=DEV)
    a = ps.optimize(VanillaPredictiveSamplerParams(key=0, x0=x0, us_guess=ug, noise=nz))
    b = VanillaPredictiveSampler(model=model, cost_function=Mine(), nsamples=100, stdev=0.01).optimize(
        VanillaPredictiveSamplerParams(key=0, x0=x0, us_guess=ug, noise=nz))
    # the fused path forms guess + stdev * noise with one FMA on the device, torch with a multiply and an add: the controls agree
    # to an ulp, and the stiff joint couplings of the hand amplify that over the horizon
    assert torch.allclose(a[1], b[1]) and torch.allclose(a[0], b[0], rtol=2e-4, atol=2e-4)


# ------------------------------------------------------------------ env step
def test_pipeline_step_and_env(load_model):
    env = PendulumSwingupEnv(num_envs=16)
    state = env.reset(0)
    assert state.obs.shape == (16, 3) and env.action_size == 1 and env.observation_size == 3
    o = Oracle(env.model)
    q, v = state.pipeline_state.qpos.cpu().numpy(), state.pipeline_state.qvel.cpu().numpy()
    act = torch.linspace(-3, 3, 16, device=DEV)[:, None]  # beyond the +-2 ctrl limit: clamped in fwd_actuation
    s1 = env.step(state, act)
    for e in range(16):
        qr, vr, _, _ = o.step(q[e], v[e], [float(act[e])])
        assert np.allclose(s1.pipeline_state.qpos[e].cpu().numpy(), qr, atol=1e-5)
        assert np.allclose(s1.pipeline_state.qvel[e].cpu().numpy(), vr, atol=1e-4)
    th, thd = s1.pipeline_state.qpos[:, 0], s1.pipeline_state.qvel[:, 0]
    assert torch.allclose(s1.obs, torch.stack((torch.cos(th), torch.sin(th), thd), -1))
    assert torch.all(s1.reward <= 0) and torch.all(s1.done == 0) and s1.info["step"] == 1


def test_substeps_and_auto_reset(load_model):
    mj, m, o = model_with(load_model, "barkour")
    rng = np.random.default_rng(9)
    E = 8
    st = [sample_state(mj, "barkour", rng) for _ in range(E)]
    q0, v0, c = (np.stack([s[i] for s in st]) for i in range(3))
    env = VectorEnvStepper(m, t32(q0), t32(v0), nsubsteps=3)
    first_q = env.first_qpos.clone()
    env.step(t32(c))
    for e in range(0, E, 3):
        w0 = o.forward(q0[e], v0[e])["qacc_warmstart"]
        qr, vr, _, tr = o.step(q0[e], v0[e], c[e], w0, nsteps=3)
        assert np.all(np.abs(env.qpos[e].cpu().numpy() - qr) <= 2e-5 + 2e-4 * np.abs(qr))
        assert np.isclose(float(env.time[e]), tr, atol=1e-6)
    after_one = env.qpos.clone()
    done = torch.zeros(E, dtype=torch.bool, device=DEV)
    done[2] = done[5] = True
    env.step(t32(c), done)
    # reset envs restart from the cached first state: they land where the first step landed
    assert torch.allclose(env.qpos[done], after_one[done], atol=1e-6)
    assert torch.allclose(env.time[done], torch.full((2,), 0.012, device=DEV))
    assert not torch.allclose(env.qpos[~done], after_one[~done], atol=1e-6)
    assert torch.equal(env.first_qpos, first_q)


def test_unsupported_option_raises(load_model):
    mj = load_model("pendulum")
    m = mjx.device_put(mj)
    with pytest.raises(NotImplementedError):
        m.replace(opt=m.opt.replace(cone=1)).handle(0)
    with pytest.raises(NotImplementedError):
        m.replace(opt=m.opt.replace(integrator=3)).handle(0)


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties_barkour(load_model):
    """BASELINE config C2 shape (4096 worlds x 1000 steps): determinism, batch independence,
    finiteness, and physical sanity. The oracle cannot cover this size in seconds; these properties do."""
    mj, m, o = model_with(load_model, "barkour")
    W, N = 4096, 1000
    g = torch.Generator(device=DEV).manual_seed(2)
    c0, q0 = t32(mj.key_ctrl("home")), mj.key_qpos("home")
    lim = t32(mj.actuator_ctrlrange)
    us = torch.clamp(c0 + 0.1 * torch.randn((W, N, mj.nu), generator=g, device=DEV), lim[:, 0], lim[:, 1])
    x0 = t32(np.concatenate([q0, np.zeros(mj.nv)])).repeat(W, 1)
    x0[:, 7:19] += (torch.rand((W, 12), generator=g, device=DEV) - 0.5) * 0.1
    eye = np.eye(37)
    cf = StaticGoalQuadraticCost(eye, 10 * eye, 0.01 * np.eye(12), np.concatenate([q0, np.zeros(mj.nv)]))
    c1 = shoot_cost(m, x0, us, cf)
    c2 = shoot_cost(m, x0, us, cf)
    assert torch.equal(c1, c2)  # bit-deterministic
    assert torch.isfinite(c1).all()
    sub = torch.tensor([0, 1, 777, 4095], device=DEV)
    xs_sub = shoot(m, x0[sub], us[sub])
    assert torch.isfinite(xs_sub).all()
    c_sub, _ = cf.cost(xs_sub, us[sub], None)
    assert torch.allclose(c_sub, c1[sub], rtol=1e-4)  # a world's result does not depend on its batch
    quat = xs_sub[:, :, 3:7]
    assert torch.allclose(quat.norm(dim=-1), torch.ones_like(quat[..., 0]), atol=1e-5)  # integrator keeps unit quaternions
    # first 20 steps of one world against the oracle
    ref = o.rollout(x0[777].cpu().numpy().astype(np.float64), us[777, :20].cpu().numpy().astype(np.float64))
    assert np.abs(xs_sub[2, :21].cpu().numpy() - ref).max() < 2e-3


def test_full_size_properties_biped(load_model):
    """BASELINE config C3 shape (16384 worlds x 1000 steps of the biped / exoskeleton-class stand-in): determinism, finiteness,
    batch independence, unit quaternions, and the first steps of one world against the oracle."""
    mj, m, o = model_with(load_model, "biped")
    W, N = 16384, 1000
    g = torch.Generator(device=DEV).manual_seed(3)
    c0, q0 = t32(mj.key_ctrl("stand")), mj.key_qpos("stand")
    lim = t32(mj.actuator_ctrlrange)
    us = torch.clamp(c0 + 0.05 * torch.randn((W, N, mj.nu), generator=g, device=DEV), lim[:, 0], lim[:, 1])
    x0 = t32(np.concatenate([q0, np.zeros(mj.nv)])).repeat(W, 1)
    x0[:, 7:mj.nq] += (torch.rand((W, mj.nq - 7), generator=g, device=DEV) - 0.5) * 0.04
    nx = mj.nq + mj.nv
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), np.concatenate([q0, np.zeros(mj.nv)]))
    c1 = shoot_cost(m, x0, us, cf)
    assert torch.equal(c1, shoot_cost(m, x0, us, cf))  # bit-deterministic
    assert torch.isfinite(c1).all()
    sub = torch.tensor([0, 5, 8191, 16383], device=DEV)
    xs_sub = shoot(m, x0[sub], us[sub])
    c_sub, _ = cf.cost(xs_sub, us[sub], None)
    assert torch.allclose(c_sub, c1[sub], rtol=1e-4)  # a world's result does not depend on its batch
    assert torch.allclose(xs_sub[:, :, 3:7].norm(dim=-1), torch.ones_like(xs_sub[:, :, 3]), atol=1e-5)
    ref = o.rollout(x0[8191].cpu().numpy().astype(np.float64), us[8191, :15].cpu().numpy().astype(np.float64))
    assert np.abs(xs_sub[2, :16].cpu().numpy() - ref).max() < 2e-3


def test_full_size_properties_sampler_sweep(load_model):
    """BASELINE config C4 at its largest size (1 048 576 samples x horizon 32, one problem): the winner is the first minimum of the
    returned costs (numpy, NaN as minimum), its controls are the guess plus its own Philox noise, its trajectory is the rollout of
    those controls, and four shards by global sample id reproduce the costs and the winner bit for bit."""
    mj, m, o = model_with(load_model, "barkour")
    nx, nu, S, N, seed = mj.nq + mj.nv, mj.nu, 1 << 20, 32, 0xC4C4
    x0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    xg = x0.copy()
    xg[0] += 0.2
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(nu), xg)
    ug = t32(np.tile(mj.key_ctrl("home"), (N, 1)))
    p = VanillaPredictiveSamplerParams(key=seed, x0=t32(x0), us_guess=ug)
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=S, stdev=0.1)
    xs, us, info = ps.optimize(p, return_info=True)
    costs = info["costs"].reshape(-1).cpu().numpy()
    assert costs.shape == (S,) and np.isfinite(costs).all()
    k = int(info["best_idx"])
    assert k == int(np.argmin(costs)) and float(info["best_cost"]) == costs[k]  # first minimum, bit-exact on the device's own costs
    lim = mj.actuator_ctrlrange.astype(np.float32)
    z = _philox.normals(seed, k, 0, np.arange(N * nu)).reshape(N, nu) if k > 0 else np.zeros((N, nu), np.float32)
    assert np.allclose(us.cpu().numpy(), np.clip(ug.cpu().numpy() + z * np.float32(0.1), lim[:, 0], lim[:, 1]), atol=2e-6)
    assert torch.equal(xs, shoot(m, t32(x0), us))  # the returned trajectory is the rollout of the returned controls
    assert costs[0] == float(shoot_cost(m, t32(x0), ug[None], cf)[0])  # sample 0 is the un-noised guess
    parts = []
    for q in range(4):
        psk = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=S // 4, stdev=0.1)
        parts.append(psk.optimize(p, sample_offset=q * (S // 4), nsamples_total=S, return_info=True)[2])
    assert torch.equal(torch.cat([q["costs"].reshape(-1) for q in parts]), info["costs"].reshape(-1))
    best = torch.stack([q["best_cost"].reshape(()) for q in parts])
    assert int(parts[int(best.argmin())]["best_idx"]) == k


def test_full_size_properties_env_step(load_model):
    """BASELINE config C5 shape (8192 auto-resetting envs): 300 fused env steps under random actions. Episode counters, done /
    truncation flags and the reset blend stay consistent for every env, rewards are finite and negative, and every env that
    finished holds its first state again."""
    from ambersim_b200.rl.wrappers import FusedQuadraticTaskEnv, QuadraticTaskEnv

    mj = load_model("barkour")
    nx, E, T, L = mj.nq + mj.nv, 8192, 300, 120
    q0 = mj.key_qpos("home")
    rew = StaticGoalQuadraticCost(np.eye(nx), np.eye(nx), 0.01 * np.eye(mj.nu), np.concatenate([q0, np.zeros(mj.nv)]))
    env = FusedQuadraticTaskEnv(QuadraticTaskEnv(mj, rew, q0, E, z_min=0.12, jitter=0.05), episode_length=L)
    st = env.reset(5)
    first_q = st.pipeline_state.qpos.clone()
    g = torch.Generator(device=DEV).manual_seed(9)
    lim = t32(mj.actuator_ctrlrange)
    steps_prev = st.info["steps"].clone()
    n_done = n_trunc = 0
    for t in range(T):
        a = torch.clamp(t32(mj.key_ctrl("home")) + 0.6 * torch.randn((E, mj.nu), generator=g, device=DEV), lim[:, 0], lim[:, 1])
        st = env.step(st, a)
        done, trunc, steps = st.done.bool(), st.info["truncation"].bool(), st.info["steps"]
        assert torch.isfinite(st.reward).all() and (st.reward <= 0).all()
        assert torch.equal(steps, steps_prev + 1) and (steps <= L).all()  # (the counter of a finished env restarts on its next step)
        assert torch.equal(done & (steps == L) | trunc, done & (steps == L)) and not (trunc & ~done).any()
        fin = done.nonzero().flatten()
        assert torch.equal(st.pipeline_state.qpos[fin], first_q[fin])  # where(done, first_state, state)
        assert torch.isfinite(st.pipeline_state.qpos).all() and torch.isfinite(st.obs).all()
        n_done += int(done.sum()); n_trunc += int(trunc.sum())
        steps_prev = torch.where(done, torch.zeros_like(steps), steps)
    assert n_done > E and n_trunc > 0 and n_done > n_trunc  # both kinds of episode ends happened


def test_urdf_loaded_model_steps_like_the_oracle():
    """A URDF goes through the loader's URDF front end (transmissions -> motors, mimic -> joint equality, welded tool link, slide joint,
    box / capsule / sphere collision geoms) and then steps on the engine like any MJCF model."""
    from ambersim_b200.utils.io_utils import load_mj_model_from_file

    mj = load_mj_model_from_file("tests/models/arm.urdf")
    m = mjx.device_put(mj)
    m = m.replace(opt=m.opt.replace(timestep=0.002, iterations=3, ls_iterations=8))
    o = Oracle(mj, m.opt)
    assert o.ne == 1 and o.nl == 2 and o.ncon >= 1
    rng = np.random.default_rng(31)
    for k in range(6):
        q = np.array([rng.uniform(-1.4, 1.9), rng.uniform(0.0, 0.25), 0.0])
        q[2] = 0.1 + 0.5 * q[0] + rng.uniform(-0.05, 0.05)
        v, c, w = rng.normal(size=3), rng.uniform(-2, 2, 2), rng.normal(size=3)
        qr, vr, wr, _ = o.step(q, v, c, w)
        d = mjx.step(m, mjx.Data(qpos=t32(q), qvel=t32(v), ctrl=t32(c), qacc=t32(v), qacc_warmstart=t32(w), time=torch.zeros((), device=DEV)))
        assert np.abs(d.qpos.cpu().numpy() - qr).max() < 2e-5 + 2e-4 * np.abs(qr).max(), k
        assert np.abs(d.qvel.cpu().numpy() - vr).max() < 2e-3 * max(1.0, np.abs(vr).max()), k


def test_ffma_peak_is_plausible():
    import ctypes as C

    tf, ms = C.c_double(), C.c_double()
    _lib.check(_lib.lib().abr_ffma_peak(0, C.byref(tf), C.byref(ms)))
    assert 20.0 < tf.value < 90.0  # nominal 74.4 TFLOP/s at 1965 MHz


# ------------------------------------------------------------------ limb (path-decomposed) kernels
@pytest.mark.parametrize("name", ["barkour", "biped", "exolegs", "tripod", "tripod3"])
def test_limb_path_is_default_and_matches_generic(load_model, name, monkeypatch):
    """Eligible models run on the limb kernels by default (lanes = 0 or 1); pinning a generic group size
    gives the same trajectory and costs up to float32 rounding, and the general-sharing build of the limb
    kernel (ABR_LIMB_GENERAL) agrees with the flat-pattern build."""
    mj, m, o = model_with(load_model, name)
    rng = np.random.default_rng(21)
    W, N = 24, 12
    key = MODEL_KEY[name]
    x0 = np.tile(np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)]), (W, 1))
    x0[:, 7:mj.nq] += rng.uniform(-0.05, 0.05, (W, mj.nq - 7))
    x0[:, mj.nq:] = 0.2 * rng.normal(size=(W, mj.nv))
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    us = mj.key_ctrl(key) + 0.1 * rng.normal(size=(W, N, mj.nu))
    us = np.where(mj.actuator_ctrllimited[None, None, :] > 0, np.clip(us, lo, hi), us)
    nx = mj.nq + mj.nv
    qd = rng.uniform(0.5, 2.0, nx)
    cf = StaticGoalQuadraticCost(np.diag(qd), 10 * np.diag(qd), 0.01 * np.eye(mj.nu), x0[0])
    out = {}
    for lanes in (0, 1, 32):
        mm = mjx.device_put(mj).replace(opt=m.opt)
        mm.set_lanes(lanes)
        out[lanes] = (shoot(mm, t32(x0), t32(us)).cpu().numpy(), shoot_cost(mm, t32(x0), t32(us), cf).cpu().numpy())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])  # default == pinned limb
    assert not np.array_equal(out[1][0], out[32][0])  # a different kernel really ran
    assert np.abs(out[1][0] - out[32][0]).max() < 2e-3
    assert np.allclose(out[1][1], out[32][1], rtol=2e-3)
    ref = o.rollout(x0, us)
    assert np.abs(out[1][0] - ref).max() < 5e-3
    assert np.allclose(out[1][1], quad_cost(ref, us, np.diag(qd), 10 * np.diag(qd), 0.01 * np.eye(mj.nu), x0[0]), rtol=1e-2)
    monkeypatch.setenv("ABR_LIMB_GENERAL", "1")
    mm = mjx.device_put(mj).replace(opt=m.opt)
    mm.set_lanes(1)
    xs_g = shoot(mm, t32(x0), t32(us)).cpu().numpy()
    assert np.abs(xs_g - out[1][0]).max() < 2e-3


def test_model_describe_names_the_serving_kernels(load_model):
    mj, m, _ = model_with(load_model, "barkour")
    assert "limb kernels <NL=3, NC=1, flat 4-lane pattern>" in m.describe() and "fast (eulerdamp off)" in m.describe()
    assert "fast (eulerdamp on)" in m.replace(opt=m.opt.replace(disableflags=0)).describe()
    assert "general variant" in m.replace(opt=m.opt.replace(iterations=3)).describe()
    assert "biped pattern, contact-body form" in model_with(load_model, "biped")[1].describe()
    assert "sharing pattern from the table" in model_with(load_model, "tripod")[1].describe()
    assert m.describe().count("\n") == 0 and "hand kernels <NL=3>" in model_with(load_model, "bh280")[1].describe()
    assert "generic kernels" in model_with(load_model, "bh280", solver=1)[1].describe() and "generic kernels" in model_with(load_model, "boxbot")[1].describe()
    m.set_lanes(16)
    assert "generic kernels, 16 lanes" in m.describe()


def test_limb_path_eligibility(load_model):
    """CG and RK4 stay on the generic kernels, and so does a fixed-base model with contacts on: pinning the limb / hand path
    there is an error, not a silent fallback. Fixed-base chains with joint equalities and no contacts (pendulum, bh280 under
    the reference fixture's options) are served by the hand kernels."""
    for name in ("pendulum", "bh280"):
        mj, m, _ = model_with(load_model, name)
        m.set_lanes(1)
        assert "hand kernels" in m.describe()
        assert shoot(m, t32(np.zeros(mj.nq + mj.nv)), t32(np.zeros((2, mj.nu)))).shape == (3, mj.nq + mj.nv)
    for name, opt in (("bh280", dict(solver=1)), ("pendulum", dict(integrator=1)), ("barkour", dict(solver=1)), ("barkour", dict(integrator=1))):
        mj, m, _ = model_with(load_model, name, **opt)
        with pytest.raises(Exception):
            m.set_lanes(1)
            shoot(m, t32(np.zeros(mj.nq + mj.nv)), t32(np.zeros((2, mj.nu))))
    mj, m, _ = model_with(load_model, "barkour")
    m.set_lanes(1)
    assert shoot(m, t32(np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])), t32(np.zeros((2, mj.nu)))).shape == (3, mj.nq + mj.nv)


@pytest.mark.parametrize("name", ["barkour", "bh280"])
def test_vps_kept_trajectories_equal_rerolled_winner(load_model, name, monkeypatch):
    """The sampler either keeps every sample's trajectory in scratch and gathers the winner (small solves) or
    re-rolls the winner (large sweeps): both give the same (xs*, us*) bit for bit."""
    mj, m, _ = model_with(load_model, name)
    nx = mj.nq + mj.nv
    x0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]) if name == "barkour" else np.zeros(nx)
    ug = np.tile(mj.key_ctrl("home"), (8, 1)) if name == "barkour" else np.zeros((8, mj.nu))
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), x0)
    res = []
    for mb in ("256", "0"):
        monkeypatch.setenv("ABR_KEEP_TRAJ_MB", mb)
        ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=64, stdev=0.05)
        xs, us, info = ps.optimize(VanillaPredictiveSamplerParams(key=5, x0=t32(x0), us_guess=t32(ug)), return_info=True)
        res.append((xs.cpu().numpy(), us.cpu().numpy(), int(info["best_idx"])))
    assert res[0][2] == res[1][2]
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_host_rollout_pipelined_slices_equal_one_launch(load_model, monkeypatch):
    """abr_rollout_host sends long control sequences up in horizon slices that overlap the previous slice's
    steps (state and running cost carried between the slice launches): same bits as the single launch."""
    mj, m, _ = model_with(load_model, "barkour")
    rng = np.random.default_rng(8)
    W, N = 40, 70
    nx = mj.nq + mj.nv
    x0 = np.tile(np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]), (W, 1)).astype(np.float32)
    x0[:, 7:mj.nq] += rng.uniform(-0.05, 0.05, (W, mj.nq - 7)).astype(np.float32)
    us = np.clip(mj.key_ctrl("home") + 0.1 * rng.normal(size=(W, N, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]).astype(np.float32)
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), x0[0])
    res = {}
    for mode, env in (("sliced", {"ABR_SLICES": "5"}), ("graded", {}), ("single", {"ABR_NO_PIPELINE": "1"})):
        for k in ("ABR_SLICES", "ABR_NO_PIPELINE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        res[mode] = (shoot(m, x0, us), shoot_cost(m, x0, us, cf))
    assert isinstance(res["sliced"][0], np.ndarray)
    assert np.array_equal(res["sliced"][0], res["single"][0]) and np.array_equal(res["graded"][0], res["single"][0])
    assert np.array_equal(np.asarray(res["sliced"][1]), np.asarray(res["single"][1])) and np.array_equal(np.asarray(res["graded"][1]), np.asarray(res["single"][1]))
    assert np.isfinite(res["sliced"][0]).all()


@pytest.mark.parametrize("name", ["barkour", "bh280"])
def test_mpc_loop_on_device_equals_python_loop(load_model, name):
    """abr_mpc_dev (solve, step the plant under us*[0], shift the guess; no host round trips) reproduces the
    Python loop over `optimize` bit for bit."""
    mj, m, _ = model_with(load_model, name)
    nx = mj.nq + mj.nv
    x0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]) if name == "barkour" else 0.1 * np.ones(nx)
    ug = np.tile(mj.key_ctrl("home"), (8, 1)) if name == "barkour" else np.zeros((8, mj.nu))
    xg = x0.copy()
    xg[0] += 0.3  # ask for some motion
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), xg)
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=48, stdev=0.1)
    T = 6
    xs_d, us_d, info = ps.mpc(VanillaPredictiveSamplerParams(key=11, x0=t32(x0), us_guess=t32(ug)), T)
    x, g = t32(x0), t32(ug)
    xs_p, us_p, idx_p = [x.clone()], [], []
    for t in range(T):
        xs, us, inf = ps.optimize(VanillaPredictiveSamplerParams(key=11 + t, x0=x, us_guess=g), return_info=True)
        x = xs[1].clone()
        g = torch.cat((us[1:], us[-1:]), dim=0)
        xs_p.append(x.clone()); us_p.append(us[0].clone()); idx_p.append(int(inf["best_idx"]))
    assert torch.equal(xs_d, torch.stack(xs_p)) and torch.equal(us_d, torch.stack(us_p))
    assert info["best_idx"].cpu().tolist() == idx_p
    assert torch.equal(info["us_guess"], g) and torch.equal(info["x"], x)


@pytest.mark.parametrize("name", ["barkour", "bh280"])
def test_mpc_closed_loop_against_the_oracle(load_model, name):
    """The receding-horizon loop end to end against the float64 oracle: every tick the oracle draws the same noise (numpy Philox),
    rolls all samples, evaluates the quadratic cost, takes the first minimum, steps the plant under us*[0] and shifts the guess.
    The device loop must pick the same winners (wherever the runner-up is resolvably worse) and visit the same states."""
    mj, m, o = model_with(load_model, name)
    nx, nu = mj.nq + mj.nv, mj.nu
    x0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]) if name == "barkour" else 0.1 * np.ones(nx)
    ug = np.tile(mj.key_ctrl("home"), (8, 1)) if name == "barkour" else np.zeros((8, nu))
    xg = x0.copy()
    xg[0] += 0.3
    Q, Qf, R = np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(nu)
    S, N, T, stdev = 48, 8, 6, 0.1
    ps = VanillaPredictiveSampler(model=m, cost_function=StaticGoalQuadraticCost(Q, Qf, R, xg), nsamples=S, stdev=stdev)
    xs_d, us_d, info = ps.mpc(VanillaPredictiveSamplerParams(key=11, x0=t32(x0), us_guess=t32(ug)), T)
    xs_d, us_d, idx_d = xs_d.cpu().numpy(), us_d.cpu().numpy(), info["best_idx"].cpu().numpy()
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    x, g = x0.astype(np.float32).astype(np.float64), ug.astype(np.float32).astype(np.float64)
    agreed = 0
    for t in range(T):
        z = np.zeros((S, N, nu), np.float32)
        z[1:] = _philox.normals(_seed_of(11 + t), np.arange(1, S)[:, None], 0, np.arange(N * nu)[None, :]).reshape(S - 1, N, nu)
        us = np.clip((g.astype(np.float32) + z * np.float32(stdev)).astype(np.float64), lo, hi)
        xs = o.rollout(x, us)
        costs = quad_cost(xs, us, Q, Qf, R, xg)
        order = np.argsort(costs, kind="stable")
        best = int(order[0])
        gap = (costs[order[1]] - costs[best]) / max(1.0, abs(costs[best]))
        if int(idx_d[t]) != best:
            assert gap < 1e-4, (t, best, int(idx_d[t]), gap)  # only an unresolvable runner-up may win instead
            break
        agreed += 1
        assert np.abs(us_d[t] - us[best, 0]).max() < 1e-5
        assert np.abs(xs_d[t + 1] - xs[best, 1]).max() < 2e-3 * max(1.0, np.abs(xs[best, 1]).max()), t
        # continue from the DEVICE's state so that rounding does not accumulate into a different problem
        x, g = xs_d[t + 1].astype(np.float64), np.concatenate([us[best, 1:], us[best, -1:]])
    assert agreed >= 4


def test_limb_path_edge_cases(load_model):
    """Empty batches, N = 0, a shared x0 row, ragged batch sizes (not a multiple of the worlds per warp) and a
    diverging world (NaN stays in its own world and sorts first in the argmin) on the limb kernels."""
    mj, m, o = model_with(load_model, "barkour")
    m.set_lanes(1)
    nx = mj.nq + mj.nv
    q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
    assert shoot(m, t32(np.zeros((0, nx))), t32(np.zeros((0, 5, mj.nu)))).shape == (0, 6, nx)  # no worlds
    x = shoot(m, t32(q0), t32(np.zeros((0, mj.nu))))  # N = 0: just x0
    assert x.shape == (1, nx) and np.array_equal(x.cpu().numpy()[0], q0.astype(np.float32))
    rng = np.random.default_rng(3)
    for W in (1, 7, 9, 33):  # 8 worlds per warp on this model
        us = np.clip(mj.key_ctrl("home") + 0.1 * rng.normal(size=(W, 6, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
        a = shoot(m, t32(q0), t32(us)).cpu().numpy()  # shared x0 row (x0_stride = 0)
        b = shoot(m, t32(np.tile(q0, (W, 1))), t32(us)).cpu().numpy()
        assert np.array_equal(a, b)
        assert np.abs(a - o.rollout(np.tile(q0, (W, 1)), us)).max() < 2e-3
        c1 = shoot_cost(m, t32(q0), t32(us), cf).cpu().numpy()
        assert np.array_equal(c1[:1], shoot_cost(m, t32(q0), t32(us[:1]), cf).cpu().numpy())  # batch-independent bits
    # one world starts from a non-finite state: its neighbours in the warp are untouched, its cost is NaN
    W = 16
    us = np.tile(mj.key_ctrl("home"), (W, 6, 1))
    x0 = np.tile(q0, (W, 1))
    x0[5, 2] = np.nan
    xs = shoot(m, t32(x0), t32(us)).cpu().numpy()
    costs = shoot_cost(m, t32(x0), t32(us), cf).cpu().numpy()
    ok = [w for w in range(W) if w != 5]
    assert np.isfinite(xs[ok]).all() and np.isfinite(costs[ok]).all() and np.isnan(costs[5])
    assert np.array_equal(xs[ok[0]], xs[ok[-1]])


@pytest.mark.parametrize("name,lanes", [("barkour", 0), ("barkour", 8), ("biped", 0), ("tripod", 0)])
def test_fused_task_env_equals_wrapper_chain(load_model, name, lanes):
    """abr_env_task_step_dev (physics + obs + reward + done + episode counter + auto-reset in ONE launch) against
    AutoResetWrapper(EpisodeWrapper(QuadraticTaskEnv)) built from torch ops around pipeline_step: identical physics
    state, done / truncation flags and step counters, rewards to float rounding; and the first step's reward and
    state against the oracle. The episode length (7) and the height threshold make both kinds of episode end occur."""
    from ambersim_b200.rl.wrappers import AutoResetWrapper, EpisodeWrapper, FusedQuadraticTaskEnv, QuadraticTaskEnv

    mj = load_model(name)
    key = MODEL_KEY[name]
    nx = mj.nq + mj.nv
    rng = np.random.default_rng(11)
    qd, rd = rng.uniform(0.5, 2.0, nx), rng.uniform(0.01, 0.1, mj.nu)
    xg = np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)])
    cf = StaticGoalQuadraticCost(np.diag(qd), np.eye(nx), np.diag(rd), xg)
    E, T, LEN = 37, 40, 7
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    acts = t32(np.clip(mj.key_ctrl(key) + 0.6 * rng.standard_normal((T, E, mj.nu)), lo, hi))

    def make(z_min):
        env = QuadraticTaskEnv(mj, cf, mj.key_qpos(key), num_envs=E, z_min=z_min, jitter=0.1, physics_steps_per_control_step=2)
        if lanes:
            env.sys.set_lanes(lanes)
        return env

    # height threshold = the 30 % quantile of the lowest base height over the first episode: some envs terminate before the length does
    pre = make(-np.inf)
    sp = pre.reset(5)
    zlow = sp.pipeline_state.qpos[:, 2].clone()
    for t in range(LEN - 1):
        sp = pre.step(sp, acts[t])
        zlow = torch.minimum(zlow, sp.pipeline_state.qpos[:, 2])
    z_min = float(torch.quantile(zlow, 0.3))
    make = (lambda f: (lambda: f(z_min)))(make)
    ref_env = AutoResetWrapper(EpisodeWrapper(make(), LEN))
    fused = FusedQuadraticTaskEnv(make(), LEN)
    s_ref, s_f = ref_env.reset(5), fused.reset(5)
    assert torch.equal(s_ref.obs, s_f.obs)
    # first step against the oracle (float64): state within the single-step bound, reward to 1e-4 relative
    o = Oracle(mj)
    d0 = s_ref.pipeline_state
    q0, v0, w0 = (t.cpu().numpy() for t in (d0.qpos, d0.qvel, d0.qacc_warmstart))
    n_term = n_trunc = 0
    for t in range(T):
        s_ref = ref_env.step(s_ref, acts[t])
        s_f = fused.step(s_f, acts[t])
        dr, df = s_ref.pipeline_state, s_f.pipeline_state
        for a, b in ((dr.qpos, df.qpos), (dr.qvel, df.qvel), (dr.qacc_warmstart, df.qacc_warmstart), (s_ref.obs, s_f.obs)):
            assert torch.equal(a, b), f"step {t}"
        assert torch.allclose(dr.time, df.time, atol=1e-6)
        assert torch.equal(s_ref.done.bool(), s_f.done.bool()) and torch.equal(s_ref.info["steps"], s_f.info["steps"])
        assert torch.equal(s_ref.info["truncation"].bool(), s_f.info["truncation"].bool())
        assert torch.allclose(s_ref.reward, s_f.reward, rtol=1e-5, atol=1e-5)
        n_trunc += int(s_f.info["truncation"].sum())
        n_term += int((s_f.done.bool() & ~s_f.info["truncation"].bool()).sum())
        if t == 0:
            for e in range(0, E, 6):
                qr, vr, _, _ = o.step(q0[e], v0[e], acts[0, e].cpu().numpy(), w0[e], nsteps=2)
                xr = np.concatenate([qr, vr])
                rr = -0.5 * (np.sum(qd * (xr - xg) ** 2) + np.sum(rd * acts[0, e].cpu().numpy().astype(np.float64) ** 2))
                assert np.isclose(float(s_f.reward[e]), rr, rtol=1e-3, atol=1e-4)
                if not bool(s_f.done[e]):
                    assert np.all(np.abs(df.qpos[e].cpu().numpy() - qr) <= 2e-5 + 2e-4 * np.abs(qr))
    assert n_trunc > 0 and n_term > 0


@pytest.mark.parametrize("name,lanes", [("barkour", 0), ("barkour", 8), ("biped", 0), ("tripod", 0), ("tripod", 16)])
def test_env_domain_randomization(load_model, name, lanes):
    """abr_env_set_randomization: every env steps its own variant of the model (contact friction scale, actuator strength
    scale). Each env is compared with an oracle built from a model whose pair_friction / gainprm / biasprm were scaled
    BEFORE loading; scale (1, 1) reproduces the unrandomised step bit for bit."""
    import copy

    mj, m, _ = model_with(load_model, name)
    if lanes:
        m.set_lanes(lanes)
    rng = np.random.default_rng(17)
    dr = np.array([[1.0, 1.0], [0.5, 1.0], [1.6, 0.8], [0.7, 1.3], [1.0, 0.6], [0.3, 1.15]])
    E = len(dr)
    q, v, c = sample_state(mj, name, rng)
    v[:2] += 0.5  # sliding feet: the friction rows are active
    w = rng.normal(size=mj.nv)
    tile = lambda a: t32(np.tile(a, (E, 1)))
    d = mjx.Data(qpos=tile(q), qvel=tile(v), ctrl=tile(c), qacc=torch.zeros(E, mj.nv, device=DEV), qacc_warmstart=tile(w),
                 time=torch.zeros(E, device=DEV))
    plain = mjx.step(m, d)
    mjx.set_randomization(m, t32(dr))
    out = mjx.step(m, d)
    with pytest.raises(_lib.AbrError):  # a different batch size than the randomisation table is an error, not a guess
        mjx.step(m, mjx.Data(qpos=t32(q), qvel=t32(v), ctrl=t32(c), qacc=t32(v), qacc_warmstart=t32(w), time=torch.zeros((), device=DEV)))
    mjx.set_randomization(m, None)
    assert torch.equal(mjx.step(m, d).qpos, plain.qpos)
    assert torch.equal(out.qpos[0], plain.qpos[0]) and torch.equal(out.qvel[0], plain.qvel[0])
    assert not torch.equal(out.qvel[1], plain.qvel[1]) and not torch.equal(out.qvel[4], plain.qvel[4])
    for e in range(E):
        mj_e = copy.deepcopy(mj)
        mj_e.pair_friction = np.array(mj.pair_friction, dtype=np.float64).copy()
        mj_e.pair_friction[:, :2] *= dr[e, 0]
        mj_e.actuator_gainprm = np.array(mj.actuator_gainprm, dtype=np.float64) * dr[e, 1]
        mj_e.actuator_biasprm = np.array(mj.actuator_biasprm, dtype=np.float64) * dr[e, 1]
        o = Oracle(mj_e, m.opt)
        qr, vr, wr, _ = o.step(q, v, c, w)
        q32, v32, _, _ = o.step(q, v, c, w, prec=1)
        assert np.all(np.abs(out.qpos[e].cpu().numpy() - qr) <= 1e-5 + 1e-4 * np.abs(qr) + 3 * np.abs(q32 - qr)), f"env {e}"
        assert np.abs(out.qvel[e].cpu().numpy() - vr).max() <= 1e-4 * max(1.0, np.abs(vr).max()) + 3 * np.abs(v32 - vr).max(), f"env {e}"
    # the randomised oracles really differ from each other by far more than the tolerance
    assert np.abs(out.qvel[1].cpu().numpy() - out.qvel[2].cpu().numpy()).max() > 1e-3


@pytest.mark.parametrize("name,lanes,disable", [("barkour", 0, 0), ("barkour", 0, 16384), ("barkour", 8, 0), ("biped", 0, 0), ("tripod", 0, 0)])
def test_env_domain_randomization_four_parameters(load_model, name, lanes, disable):
    """abr_env_set_randomization_ex with four scales per env: contact friction, actuator strength, joint damping (the passive force AND
    the implicit-damping Euler term: both settings of eulerdamp) and joint armature. Each env is compared with an oracle whose host
    model carries the scaled dof_damping / dof_armature with the base model's compiled constants, which is what
    `model.replace(dof_damping=..., dof_armature=...)` gives an MJX user."""
    import copy

    mj, m, _ = model_with(load_model, name, disableflags=disable)
    if lanes:
        m.set_lanes(lanes)
    rng = np.random.default_rng(19)
    dr = np.array([[1.0, 1.0, 1.0, 1.0], [1.0, 1.0, 3.0, 1.0], [1.0, 1.0, 1.0, 4.0], [0.6, 1.2, 0.3, 2.5], [1.4, 0.7, 2.0, 0.5]])
    E = len(dr)
    q, v, c = sample_state(mj, name, rng)
    v[:2] += 0.5
    w = rng.normal(size=mj.nv)
    tile = lambda a: t32(np.tile(a, (E, 1)))
    d = mjx.Data(qpos=tile(q), qvel=tile(v), ctrl=tile(c), qacc=torch.zeros(E, mj.nv, device=DEV), qacc_warmstart=tile(w),
                 time=torch.zeros(E, device=DEV))
    plain = mjx.step(m, d)
    mjx.set_randomization(m, t32(dr))
    out = mjx.step(m, d)
    mjx.set_randomization(m, None)
    assert torch.equal(out.qpos[0], plain.qpos[0]) and torch.equal(out.qvel[0], plain.qvel[0])  # unit scales: the same numbers
    assert not torch.equal(out.qvel[1], plain.qvel[1]) and not torch.equal(out.qvel[2], plain.qvel[2])
    for e in range(E):
        mj_e = copy.deepcopy(mj)
        mj_e.pair_friction = np.array(mj.pair_friction, dtype=np.float64).copy()
        mj_e.pair_friction[:, :2] *= dr[e, 0]
        mj_e.actuator_gainprm = np.array(mj.actuator_gainprm, dtype=np.float64) * dr[e, 1]
        mj_e.actuator_biasprm = np.array(mj.actuator_biasprm, dtype=np.float64) * dr[e, 1]
        mj_e.dof_damping = np.array(mj.dof_damping, dtype=np.float64) * dr[e, 2]
        mj_e.dof_armature = np.array(mj.dof_armature, dtype=np.float64) * dr[e, 3]
        o = Oracle(mj_e, m.opt)
        qr, vr, wr, _ = o.step(q, v, c, w)
        q32, v32, _, _ = o.step(q, v, c, w, prec=1)
        assert np.all(np.abs(out.qpos[e].cpu().numpy() - qr) <= 1e-5 + 1e-4 * np.abs(qr) + 3 * np.abs(q32 - qr)), f"env {e}"
        assert np.abs(out.qvel[e].cpu().numpy() - vr).max() <= 1e-4 * max(1.0, np.abs(vr).max()) + 3 * np.abs(v32 - vr).max(), f"env {e}"
    assert np.abs(out.qvel[1].cpu().numpy() - out.qvel[2].cpu().numpy()).max() > 1e-3


@pytest.mark.parametrize("name", ["bh280", "barkour"])
def test_finite_difference_shooting(load_model, name):
    """Gradient-based shooting on the engine: the central-difference gradient (2 N nu + 1 rollouts in one launch) equals the
    float64 oracle's on the contact-free hand, and on both models the descent never raises the cost and ends below the
    start, with the returned trajectory being the rollout of the returned controls."""
    from ambersim_b200.trajopt.shooting import FiniteDifferenceShooting, ShootingParams

    mj, m, o = model_with(load_model, name)
    rng = np.random.default_rng(23)
    nx, N = mj.nq + mj.nv, 6
    if name == "bh280":
        x0 = np.zeros(nx)
        xg = np.concatenate([rng.uniform(0.1, 0.6, mj.nq), np.zeros(mj.nv)])
        us = rng.uniform(0.0, 0.3, (N, mj.nu))
    else:
        x0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
        xg = x0.copy()
        xg[7:mj.nq] += rng.uniform(-0.2, 0.2, mj.nq - 7)
        us = np.tile(mj.key_ctrl("home"), (N, 1))
    Q, Qf, R = np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu)
    cf = StaticGoalQuadraticCost(Q, Qf, R, xg)
    fd = FiniteDifferenceShooting(model=m, cost_function=cf, iterations=6, eps=1e-2)
    c0, g = fd.gradient(t32(x0), t32(us))
    if name == "bh280":
        ref = np.zeros((N, mj.nu))
        for t in range(N):
            for u in range(mj.nu):
                up, dn = us.copy(), us.copy()
                up[t, u] += 1e-2
                dn[t, u] -= 1e-2
                cp, cm = (quad_cost(o.rollout(x0[None], a[None]), a[None], Q, Qf, R, xg)[0] for a in (up, dn))
                ref[t, u] = (cp - cm) / 2e-2
        assert np.abs(g.cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()
    xs, us_star, info = fd.optimize(ShootingParams(x0=t32(x0), us_guess=t32(us)), return_info=True)
    costs = info["costs"].cpu().numpy()
    assert np.isclose(costs[0], float(c0), rtol=1e-6)
    assert np.all(np.diff(costs) <= 0) and costs[-1] < 0.98 * costs[0]
    assert torch.equal(xs, shoot(m, t32(x0), us_star))
    lo, hi = mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1]
    lim = mj.actuator_ctrllimited > 0
    assert np.all((us_star.cpu().numpy() >= lo - 1e-6)[:, lim]) and np.all((us_star.cpu().numpy() <= hi + 1e-6)[:, lim])


def test_model_reserve_presizes_scratch(load_model, vps_data=None):
    """abr_model_reserve sizes the handle-owned scratch up front; results are unchanged by it."""
    mj, m, _ = model_with(load_model, "barkour")
    nx = mj.nq + mj.nv
    q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=256, stdev=0.1)
    prm = VanillaPredictiveSamplerParams(key=1, x0=t32(q0), us_guess=t32(np.tile(mj.key_ctrl("home"), (16, 1))))
    xs_a, us_a = ps.optimize(prm)
    m2 = mjx.device_put(mj).replace(opt=m.opt)
    h = m2.handle()
    _lib.check(_lib.lib().abr_model_reserve(h.ptr, 512, 64, 1, 256))
    xs_b, us_b = VanillaPredictiveSampler(model=m2, cost_function=cf, nsamples=256, stdev=0.1).optimize(prm)
    assert torch.equal(xs_a, xs_b) and torch.equal(us_a, us_b)


@pytest.mark.parametrize("name", ["pendulum", "bh280", "barkour", "biped"])
def test_committed_golden_fixtures(load_model, name):
    """The CUDA path against the fixtures committed under tests/golden/ (float64 oracle rollouts written by
    tools/make_oracle_fixtures.py; an MJX dump, mjx_<model>.npz from tools/dump_mjx_golden.py, is used too when present):
    same inputs, same options, free-running 20-step rollouts within the rollout bounds of this file's header."""
    from pathlib import Path

    gold = Path(__file__).parent / "golden"
    used = 0
    for path, tol_free, tol_contact in ((gold / f"oracle_{name}.npz", 1e-3, 5e-3), (gold / f"mjx_{name}.npz", 2e-3, 1e-2)):
        if not path.exists():
            continue
        z = np.load(path)
        opt = {k[4:]: (float(z[k]) if k == "opt_timestep" else int(z[k])) for k in z.files if k.startswith("opt_")}
        mj, m, o = model_with(load_model, name, **opt)
        xs = shoot(m, t32(z["x0"]), t32(z["us"])).cpu().numpy()
        assert xs.shape == z["xs"].shape
        drift32 = np.abs(o.rollout(z["x0"], z["us"], prec=1) - z["xs"]).max() if "oracle" in path.name else 0.0
        tol = tol_free if name in ("pendulum", "bh280") else tol_contact
        assert np.abs(xs - z["xs"]).max() < max(tol, 3 * drift32), path.name
        used += 1
    assert used >= 1


def test_solve_is_cuda_graph_capturable(load_model):
    """The *_dev entry points are stream-ordered, never synchronise and (after a first call has sized the scratch) never
    allocate: a predictive-sampling solve captures into a CUDA graph and replays with bit-identical results on new inputs."""
    mj, m, _ = model_with(load_model, "barkour")
    nx = mj.nq + mj.nv
    q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=512, stdev=0.2)
    rng = np.random.default_rng(5)
    x_in = t32(q0)
    g_in = t32(np.tile(mj.key_ctrl("home"), (8, 1)))
    prm = VanillaPredictiveSamplerParams(key=9, x0=x_in, us_guess=g_in)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            ps.optimize(prm)  # sizes the handle's scratch outside the capture
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        xs_g, us_g = ps.optimize(prm)
    for trial in range(3):
        x_new = q0.copy()
        x_new[7:19] += rng.uniform(-0.1, 0.1, 12)
        x_in.copy_(t32(x_new))
        g_in.copy_(t32(np.clip(mj.key_ctrl("home") + 0.2 * rng.normal(size=(8, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])))
        graph.replay()
        torch.cuda.synchronize()
        xs_r, us_r = xs_g.clone(), us_g.clone()
        xs_d, us_d = ps.optimize(prm)
        assert torch.equal(xs_r, xs_d) and torch.equal(us_r, us_d), f"trial {trial}"


@pytest.mark.parametrize("eulerdamp", [False, True])
@pytest.mark.parametrize("name", ["barkour", "biped", "exolegs"])
def test_limb_fast_variants_match_the_general_variant(load_model, name, eulerdamp, monkeypatch):
    """The compile-time variants of the limb kernels (no eulerdamp / multi-iteration / output / other-mode code in the loop) against the
    general variant of the same kernel (ABR_LIMB_NOSPEC): rollouts with and without trajectories, sampler costs, env steps; with the
    models' own options (eulerdamp disabled) and with MuJoCo's default (implicit joint damping in the Euler step)."""
    mj, m, o = model_with(load_model, name, disableflags=0 if eulerdamp else 16384)
    rng = np.random.default_rng(31)
    key = MODEL_KEY[name]
    W, N = 16, 10
    nx = mj.nq + mj.nv
    x0 = np.tile(np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)]), (W, 1))
    x0[:, 7:mj.nq] += rng.uniform(-0.05, 0.05, (W, mj.nq - 7))
    us = np.clip(mj.key_ctrl(key) + 0.1 * rng.normal(size=(W, N, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), x0[0])

    def run():
        mm = mjx.device_put(mj).replace(opt=m.opt)
        xs = shoot(mm, t32(x0), t32(us)).cpu().numpy()
        costs = shoot_cost(mm, t32(x0), t32(us), cf).cpu().numpy()
        ps = VanillaPredictiveSampler(model=mm, cost_function=cf, nsamples=64, stdev=0.1)
        _, _, info = ps.optimize(VanillaPredictiveSamplerParams(key=2, x0=t32(x0[0]), us_guess=t32(us[0])), return_info=True)
        d = mjx.Data(qpos=t32(x0[:, :mj.nq]), qvel=t32(x0[:, mj.nq:]), ctrl=t32(us[:, 0]), qacc=torch.zeros(W, mj.nv, device=DEV),
                     qacc_warmstart=torch.zeros(W, mj.nv, device=DEV), time=torch.zeros(W, device=DEV))
        return xs, costs, info["costs"].cpu().numpy(), mjx.step(mm, d).qvel.cpu().numpy()

    fast = run()
    monkeypatch.setenv("ABR_LIMB_NOSPEC", "1")
    general = run()
    assert np.abs(fast[0] - general[0]).max() < 2e-3 and np.allclose(fast[1], general[1], rtol=2e-3)
    assert np.allclose(fast[2], general[2], rtol=2e-3) and np.abs(fast[3] - general[3]).max() < 1e-3
    ref = o.rollout(x0, us)
    assert np.abs(general[0] - ref).max() < 5e-3 and np.abs(fast[0] - ref).max() < 5e-3


def test_peer_exchange_single_rank_and_argument_checks():
    """abr_xchg_* with one rank (the degenerate collective: the record goes through the rank's own exchange buffer, flags and
    selection included) returns the local winners unchanged, call after call (epoch parity alternates); capacity and size errors
    are refused. The multi-GPU equality is checked by tools/mgpu_check.py under torchrun (profiles/r1_mgpu_check_*)."""
    import ctypes as C

    L = _lib.lib()
    B, nxs, nus = 3, 33 * 37, 32 * 12
    x, handle = C.c_void_p(), C.create_string_buffer(_lib.xchg_handle_bytes())
    _lib.check(L.abr_xchg_create(0, 1, 0, B * (2 + nxs + nus), C.byref(x), handle))
    _lib.check(L.abr_xchg_connect(x, handle.raw))
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=DEV).manual_seed(0)
    for call in range(5):
        cost = torch.rand(B, generator=g, device=DEV)
        cost[1] = float("nan")
        idx = torch.randint(0, 10000, (B,), generator=g, device=DEV, dtype=torch.int32)
        xs, us = torch.randn((B, nxs), generator=g, device=DEV), torch.randn((B, nus), generator=g, device=DEV)
        xs_o, us_o, idx_o, cost_o = torch.empty_like(xs), torch.empty_like(us), torch.empty_like(idx), torch.empty_like(cost)
        _lib.check(L.abr_xchg_merge_best_dev(x, p(cost), p(idx), p(xs), p(us), B, nxs, nus, p(xs_o), p(us_o), p(idx_o), p(cost_o), stream))
        torch.cuda.synchronize()
        assert torch.equal(xs_o, xs) and torch.equal(us_o, us) and torch.equal(idx_o, idx)
        assert torch.equal(torch.isnan(cost_o), torch.isnan(cost)) and torch.equal(cost_o[[0, 2]], cost[[0, 2]])
    flag = C.c_int(-1)
    _lib.check(L.abr_xchg_timed_out(x, C.byref(flag)))
    assert flag.value == 0
    assert L.abr_xchg_merge_best_dev(x, p(cost), p(idx), p(xs), p(us), B + 1, nxs, nus, p(xs_o), p(us_o), p(idx_o), p(cost_o), stream) == _lib.ABR_ECAPACITY
    assert L.abr_xchg_merge_best_dev(x, p(cost), p(idx), p(xs), p(us), 0, nxs, nus, p(xs_o), p(us_o), p(idx_o), p(cost_o), stream) == _lib.ABR_EINVAL
    _lib.check(L.abr_xchg_destroy(x))


# ------------------------------------------------------------------ derived mjx.Data fields, batched
@pytest.mark.parametrize("name", ["barkour", "bh280", "biped"])
def test_batched_derived_fields_match_oracle(load_model, name):
    """forward(..., fields=...) / step(..., fields=...) fill the mjx.Data fields an env's compute_obs reads (rl/base.py:98-125)
    for a whole batch in the launch that does the physics; every world is compared with the float64 oracle. After a step the
    fields are those of the last forward pass (the state BEFORE the final integration), as mjx.step leaves them."""
    mj, model, o = model_with(load_model, name)
    rng = np.random.default_rng(21)
    E = 24
    key = {"barkour": "home", "biped": "stand"}.get(name)
    q0 = mj.key_qpos(key) if key else mj.qpos0.copy()
    qs = np.tile(q0, (E, 1))
    if key:
        qs[:, 7:] += rng.uniform(-0.1, 0.1, (E, mj.nq - 7))
        qs[:, 2] -= rng.uniform(0.0, 0.01, E)
    else:
        qs += rng.uniform(0.0, 0.5, (E, mj.nq))
    vs, cs, ws = 0.3 * rng.normal(size=(E, mj.nv)), 0.2 * rng.normal(size=(E, mj.nu)), rng.normal(size=(E, mj.nv))
    if key:
        cs += mj.key_ctrl(key)
    names = list(mjx.DERIVED_FIELDS)
    d = mjx.Data(qpos=t32(qs), qvel=t32(vs), ctrl=t32(cs), qacc=torch.zeros((E, mj.nv), device=DEV), qacc_warmstart=t32(ws),
                 time=torch.zeros(E, device=DEV))
    f = mjx.forward(model, d, fields=names)
    plain = mjx.forward(model, d)
    assert float((f.qacc - plain.qacc).abs().max()) < 5e-4 * float(plain.qacc.abs().max())  # limb kernels (plain) vs generic kernels (fields)
    for e in range(E):
        ref = o.forward(qs[e], vs[e], cs[e], ws[e])
        for n in names:
            got = getattr(f, n)[e].cpu().numpy().astype(np.float64).ravel()
            r = ref[n].ravel()
            if n in ("efc_force", "efc_D", "efc_aref") and r.size != got.size:
                continue
            assert r.size == got.size, n
            if r.size:
                assert np.abs(r - got).max() <= 3e-4 * max(1e-6, np.abs(r).max()) + 1e-6, (n, e)
    # step: the state advances like the plain step; the fields are the pre-integration ones of the same launch
    s = mjx.step(model, f.replace(qacc_warmstart=plain.qacc_warmstart), fields=("xpos", "cvel", "contact_dist"))
    s_plain = mjx.step(model, plain)
    assert torch.allclose(s.qpos, s_plain.qpos, atol=2e-5) and torch.allclose(s.qvel, s_plain.qvel, rtol=1e-3, atol=2e-4)
    assert torch.allclose(s.xpos, f.xpos, atol=1e-6) and torch.allclose(s.cvel, f.cvel, atol=1e-5)


# ------------------------------------------------------------------ hand kernels: fixed-base chains + joint equalities
HAND_VARIANTS = [dict(), dict(disableflags=16 | 16384), dict(iterations=3, ls_iterations=8), dict(disableflags=16 | 256), dict(disableflags=16 | 2),
                 dict(disableflags=16 | 8), dict(disableflags=16 | 32 | 64)]


@pytest.mark.parametrize("name", ["bh280", "fixedbase"])
def test_hand_kernels_serve_fixed_base_models_and_match_oracle(load_model, name):
    """The reference's own model (bh280: palm welded to the world, three finger chains, four joint equalities of which one ties
    two fingers together) runs on the register-resident hand kernels (abr_hand.cuh): rollouts against the float64 oracle and
    against the generic kernels, under the reference fixture's options and six option variants."""
    for var in HAND_VARIANTS:
        opt = dict(timestep=0.002, iterations=1, ls_iterations=4, integrator=0, solver=2, disableflags=16)
        opt.update(var)
        mj = load_model(name)
        m = mjx.device_put(mj)
        m = m.replace(opt=m.opt.replace(**opt))
        o = Oracle(mj, m.opt)
        assert "hand kernels" in m.describe(), m.describe()
        rng = np.random.default_rng(31)
        W, N = 12, 25
        lo, hi = mj.jnt_range[:, 0], mj.jnt_range[:, 1]
        span = np.where(hi > lo, hi - lo, 1.0)
        x0 = np.concatenate([np.where(hi > lo, lo, -0.5) + span * rng.uniform(-0.15, 1.15, (W, mj.nq)), 0.5 * rng.normal(size=(W, mj.nv))], axis=1)
        us = rng.normal(size=(W, N, mj.nu)) * 1.5
        xs = shoot(m, t32(x0), t32(us)).cpu().numpy()
        ref = o.rollout(x0, us)
        ref32 = o.rollout(x0, us, prec=1)
        drift = np.abs(ref32 - ref).max(axis=(0, 2))
        err = np.abs(xs - ref).max(axis=(0, 2))
        scale = max(1.0, np.abs(ref).max())
        assert err[1] < 2e-5 * scale, (var, err[1])
        assert np.all(err < np.maximum(2e-3 * scale, 10 * drift)), (var, err, drift)
        m.set_lanes(8)
        gen = shoot(m, t32(x0), t32(us)).cpu().numpy()
        m.set_lanes(0)
        assert np.abs(gen - xs).max() < np.maximum(2e-3 * scale, 10 * drift.max())
        assert np.array_equal(xs[:, 0], x0.astype(np.float32))  # row 0 = the caller's x0 verbatim (shooting.py:47)


def test_hand_kernels_sampler_equals_generic_winner(vps_data):
    """The reference fixture (bh280, 100 samples x horizon 10) through the hand kernels: same winner as the generic kernels."""
    ps, model, cf, o = vps_data
    rng = np.random.default_rng(5)
    x0, ug = t32(rng.uniform(0.0, 0.4, (4, 16))), t32(rng.normal(size=(4, 10, 4)))
    p = VanillaPredictiveSamplerParams(key=9, x0=x0, us_guess=ug)
    assert "hand kernels" in model.describe()
    xs, us, info = ps.optimize(p, return_info=True)
    model.set_lanes(8)
    xs_g, us_g, info_g = ps.optimize(p, return_info=True)
    model.set_lanes(0)
    assert torch.allclose(info["costs"], info_g["costs"], rtol=2e-4)
    assert torch.equal(us, us_g) or torch.allclose(info["best_cost"], info_g["best_cost"], rtol=2e-4)
    assert np.array_equal(info["best_idx"].cpu().numpy(), np.argmin(info["costs"].cpu().numpy(), axis=1))


def test_user_env_reads_derived_data_fields(load_model):
    """A user-defined MjxEnv whose observation and reward read derived mjx.Data fields (body positions, body twists, contact
    distances: the pattern rl/base.py:98-125 is designed for), wrapped in the Episode / AutoReset wrappers: the fields arrive
    batched from the physics launch and match the oracle's; auto-reset blends them like every other leaf of the state."""
    from ambersim_b200.rl.base import MjxEnv, State
    from ambersim_b200.rl.wrappers import AutoResetWrapper, EpisodeWrapper

    mj = load_model("barkour")

    class FootEnv(MjxEnv):
        data_fields = ("xpos", "cvel", "contact_dist")

        def reset(self, rng):
            g = torch.Generator(device=DEV).manual_seed(int(rng))
            q = t32(np.tile(mj.key_qpos("home"), (16, 1)))
            q[:, 7:] += (torch.rand((16, mj.nq - 7), generator=g, device=DEV) - 0.5) * 0.1
            d = self.pipeline_init(q, torch.zeros((16, mj.nv), device=DEV))
            z = torch.zeros(16, device=DEV)
            return State(d, self.compute_obs(d, {}), z, z.clone())

        def compute_obs(self, data, info):
            return torch.cat((data.xpos[:, 1], data.cvel[:, 1], data.contact_dist), dim=-1)  # trunk position, trunk twist, foot clearances

        def compute_reward(self, data, info):
            return -data.contact_dist.clamp(min=0).sum(-1)

        def step(self, state, action):
            d = self.pipeline_step(state.pipeline_state, action)
            done = (d.qpos[:, 2] < 0.15).float()
            return state.replace(pipeline_state=d, obs=self.compute_obs(d, state.info), reward=self.compute_reward(d, state.info), done=done)

    env = AutoResetWrapper(EpisodeWrapper(FootEnv(mj), episode_length=5))
    s = env.reset(3)
    assert s.obs.shape == (16, 3 + 6 + 4)
    o = Oracle(mj)
    ref = o.forward(s.pipeline_state.qpos[2].cpu().numpy(), np.zeros(mj.nv))
    assert np.allclose(s.obs[2, :3].cpu().numpy(), ref["xpos"][1], atol=1e-5) and np.allclose(s.obs[2, 9:].cpu().numpy(), ref["contact_dist"], atol=1e-5)
    acts = t32(np.tile(mj.key_ctrl("home"), (16, 1)))
    prev = s
    for t in range(6):
        s = env.step(s, acts)
        assert torch.isfinite(s.obs).all() and torch.isfinite(s.reward).all()
        if t == 0:  # mjx.step leaves the fields of the state BEFORE the integration: the trunk position of the previous qpos
            assert torch.allclose(s.pipeline_state.xpos[:, 1], prev.pipeline_state.qpos[:, :3], atol=1e-6)
    assert bool((s.info["steps"] <= 5).all())


def test_spline_predictive_sampler(load_model):
    """Knot-parameterised predictive sampling (the alternate parameterisation ShootingParams.N anticipates, shooting.py:66-73):
    the interpolation weights are a partition of unity that reproduces the knots, sample 0 is the un-noised spline, the winner's
    knots regenerate the winner's controls, and the winner is no worse than the guess."""
    from ambersim_b200.trajopt.shooting import SplinePredictiveSampler, SplineShootingParams

    mj, m, o = model_with(load_model, "barkour")
    nx = mj.nq + mj.nv
    q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
    rng = np.random.default_rng(2)
    x0 = q0.copy()
    x0[7:19] += rng.uniform(-0.1, 0.1, 12)
    knots = np.clip(mj.key_ctrl("home") + 0.2 * rng.standard_normal((5, mj.nu)), mj.actuator_ctrlrange[:, 0], mj.actuator_ctrlrange[:, 1])
    lim = t32(mj.actuator_ctrlrange)
    for interp in ("zoh", "linear", "cubic"):
        ps = SplinePredictiveSampler(model=m, cost_function=cf, nsamples=256, stdev=0.1, interp=interp)
        W = ps.weights(5, 33)
        assert torch.allclose(W.sum(1), torch.ones(33)) and torch.allclose(W[::8], torch.eye(5), atol=1e-6)
        xs, ks, info = ps.optimize(SplineShootingParams(key=4, x0=t32(x0), us_guess=t32(knots), horizon=33), return_info=True)
        assert xs.shape == (34, nx) and ks.shape == (5, mj.nu)
        us_from_knots = torch.minimum(torch.maximum(W.to(DEV) @ ks, lim[:, 0]), lim[:, 1])
        assert torch.allclose(info["us_star"], us_from_knots, atol=1e-5)
        assert torch.allclose(shoot(m, t32(x0), info["us_star"]), xs, atol=1e-4)
        guess_cost = shoot_cost(m, t32(x0), torch.minimum(torch.maximum(W.to(DEV) @ t32(knots), lim[:, 0]), lim[:, 1])[None], cf)[0]
        assert float(info["best_cost"]) <= float(guess_cost) * (1 + 1e-5)
        one = SplinePredictiveSampler(model=m, cost_function=cf, nsamples=1, stdev=0.1, interp=interp)
        _, k1 = one.optimize(SplineShootingParams(key=4, x0=t32(x0), us_guess=t32(knots), horizon=33))
        assert torch.equal(k1, t32(knots))


def test_caller_provided_workspace(load_model):
    """`abr_workspace_bytes` + a caller-owned workspace, after which the stream-ordered solve calls never allocate: same
    winners bit for bit as with handle-owned scratch, ABR_ECAPACITY (not a cudaMalloc) for a solve the workspace was not sized for,
    and a sweep whose kept trajectories do not fit the workspace re-rolls its winners instead."""
    mj, m, _ = model_with(load_model, "barkour")
    nx = mj.nq + mj.nv
    q0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)])
    cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), q0)
    prm = VanillaPredictiveSamplerParams(key=2, x0=t32(q0), us_guess=t32(np.tile(mj.key_ctrl("home"), (16, 1))))
    ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=512, stdev=0.1)
    ref = ps.optimize(prm, return_info=True)
    L, h = _lib.lib(), m.handle(0)
    need = C.c_size_t()
    _lib.check(L.abr_workspace_bytes(h.ptr, 16, 1, 512, C.byref(need)))
    assert need.value >= 4 * 512 * (17 * nx + 16 * mj.nu)
    ws = torch.empty(need.value, dtype=torch.uint8, device=DEV)
    assert L.abr_model_set_workspace(h.ptr, C.c_void_p(ws.data_ptr()), need.value - 1024, 16, 1, 512) == _lib.ABR_ECAPACITY
    _lib.check(L.abr_model_set_workspace(h.ptr, C.c_void_p(ws.data_ptr()), need.value, 16, 1, 512))
    got = ps.optimize(prm, return_info=True)
    assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1]) and int(ref[2]["best_idx"]) == int(got[2]["best_idx"])
    big = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=4096, stdev=0.1)
    with pytest.raises(_lib.AbrError) as e:
        big.optimize(prm)  # per-sample costs of 4096 samples do not fit a workspace sized for 512
    assert e.value.code == _lib.ABR_ECAPACITY
    xs_m, us_m, _ = ps.mpc(prm, 3)
    _lib.check(L.abr_model_set_workspace(h.ptr, None, 0, 0, 0, 0))  # back to handle-owned scratch
    xs_o, us_o, _ = ps.mpc(prm, 3)
    assert torch.equal(xs_m, xs_o) and torch.equal(us_m, us_o)
    assert torch.isfinite(big.optimize(prm)[0]).all()


def test_limb_kernels_are_bit_identical_across_cta_sizes(load_model, monkeypatch):
    """The CTA size of a limb launch (limb::pick_tpb: 8 warps per CTA share one SM's instruction stream) is a scheduling choice:
    the same worlds give the same bits whether a warp runs alone in its CTA or with seven others, for shoot, the fused cost and
    the sampler."""
    for name, key in (("barkour", "home"), ("biped", "stand")):
        mj, m, _ = model_with(load_model, name)
        nx = mj.nq + mj.nv
        rng = np.random.default_rng(9)
        W, N = 300, 12  # 38 warps: the policy itself (tpb None) picks 6-warp CTAs for the biped class, 7-warp CTAs for Barkour (single wave)
        x0 = np.tile(np.concatenate([mj.key_qpos(key), np.zeros(mj.nv)]), (W, 1))
        x0[:, 7:mj.nq] += rng.uniform(-0.05, 0.05, (W, mj.nq - 7))
        us = mj.key_ctrl(key) + 0.1 * rng.standard_normal((W, N, mj.nu))
        cf = StaticGoalQuadraticCost(np.eye(nx), 10 * np.eye(nx), 0.01 * np.eye(mj.nu), x0[0])
        ps = VanillaPredictiveSampler(model=m, cost_function=cf, nsamples=300, stdev=0.1)
        prm = VanillaPredictiveSamplerParams(key=1, x0=t32(x0[0]), us_guess=t32(us[0]))
        out = []
        for tpb in ("32", "96", "192", "224", "256", None):
            if tpb is None:
                monkeypatch.delenv("ABR_LIMB_TPB", raising=False)
            else:
                monkeypatch.setenv("ABR_LIMB_TPB", tpb)
            out.append((shoot(m, t32(x0), t32(us)), shoot_cost(m, t32(x0), t32(us), cf), *ps.optimize(prm)))
        for o in out[1:]:
            assert all(torch.equal(a, b) for a, b in zip(out[0], o))
