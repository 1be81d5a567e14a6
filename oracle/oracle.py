"""Python binding of the CPU oracle (oracle/abr_oracle.cc). TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by the product package. Parity unpinned: see the header of abr_oracle.cc.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from ambersim_b200 import _abi

_DIR = Path(__file__).resolve().parent
_lib = None


def build(force: bool = False) -> Path:
    so = _DIR / "libabr_oracle.so"
    src = _DIR / "abr_oracle.cc"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, _abi.HEADER.stat().st_mtime):
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.orc_count_flops_step.restype = C.c_longlong
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


class Oracle:
    """CPU restatement of mjx.forward / mjx.step / shoot for one flattened model."""

    def __init__(self, mjmodel, opt=None):
        self.m = mjmodel
        self.h, self._keep = _abi.pack_model(mjmodel, opt)
        ncon, ne, nl, nefc = (C.c_int() for _ in range(4))
        lib().orc_sizes(C.byref(self.h), C.byref(ncon), C.byref(ne), C.byref(nl), C.byref(nefc))
        self.ncon, self.ne, self.nl, self.nefc = ncon.value, ne.value, nl.value, nefc.value
        self.nq, self.nv, self.nu = mjmodel.nq, mjmodel.nv, mjmodel.nu
        self.nx = self.nq + self.nv

    def forward(self, qpos, qvel, ctrl=None, qacc_warmstart=None, prec=0):
        """mjx.forward on one world; returns a dict of every intermediate field (float64 arrays)."""
        qpos, qvel = _d(qpos, self.nq), _d(qvel, self.nv)
        ctrl = _d(np.zeros(self.nu) if ctrl is None else ctrl, self.nu)
        warm = _d(np.zeros(self.nv) if qacc_warmstart is None else qacc_warmstart, self.nv)
        lib().orc_forward(C.byref(self.h), prec, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm))
        out = {}
        m, nv, nb = self.m, self.nv, self.m.nbody
        shapes = dict(
            qpos=(self.nq,), xpos=(nb, 3), xquat=(nb, 4), xmat=(nb, 3, 3), xipos=(nb, 3), ximat=(nb, 3, 3),
            xanchor=(m.njnt, 3), xaxis=(m.njnt, 3), geom_xpos=(m.ngeom, 3), geom_xmat=(m.ngeom, 3, 3),
            subtree_com=(nb, 3), cinert=(nb, 10), cdof=(nv, 6), crb=(nb, 10), qM=(nv, nv), qLD=(nv, nv),
            contact_dist=(self.ncon,), contact_pos=(self.ncon, 3), contact_frame=(self.ncon, 3, 3),
            efc_J=(self.nefc, nv), efc_D=(self.nefc,), efc_aref=(self.nefc,), efc_pos=(self.nefc,),
            actuator_length=(self.nu,), actuator_velocity=(self.nu,), actuator_force=(self.nu,),
            qfrc_actuator=(nv,), cvel=(nb, 6), cdof_dot=(nv, 6), qfrc_passive=(nv,), qfrc_bias=(nv,),
            qfrc_smooth=(nv,), qacc_smooth=(nv,), qacc=(nv,), qfrc_constraint=(nv,), efc_force=(self.nefc,),
            qacc_warmstart=(nv,), solver_niter=(1,),
        )
        for name, shape in shapes.items():
            size = int(np.prod(shape))
            buf = np.zeros(max(size, 1))
            n = C.c_int()
            rc = lib().orc_get(name.encode(), _dp(buf), buf.size, C.byref(n))
            if rc != 0:
                raise RuntimeError(f"oracle field {name}: rc={rc}")
            out[name] = buf[:size].reshape(shape).copy()
        return out

    def step(self, qpos, qvel, ctrl, qacc_warmstart=None, time=0.0, nsteps=1, prec=0):
        """nsteps x mjx.step with ctrl held. Returns (qpos, qvel, qacc_warmstart, time)."""
        qpos, qvel = _d(qpos, self.nq).copy(), _d(qvel, self.nv).copy()
        warm = _d(np.zeros(self.nv) if qacc_warmstart is None else qacc_warmstart, self.nv).copy()
        ctrl = _d(ctrl, self.nu)
        t = C.c_double(time)
        lib().orc_step(C.byref(self.h), prec, _dp(qpos), _dp(qvel), _dp(warm), C.byref(t), _dp(ctrl), int(nsteps))
        return qpos, qvel, warm, t.value

    def rollout(self, x0, us, prec=0, nthreads=1, return_xs=True):
        """`shoot` (shooting.py:22-48) for a batch: x0 (nx,) or (W,nx); us (W,N,nu) or (N,nu)."""
        us = _d(us)
        single = us.ndim == 2
        if single:
            us = us[None]
        W, N, _ = us.shape
        x0 = _d(x0)
        x0_stride = 0 if x0.ndim == 1 else self.nx
        xs = np.zeros((W, N + 1, self.nx)) if return_xs else None
        xf = np.zeros((W, self.nx))
        lib().orc_rollout_batch(C.byref(self.h), prec, _dp(x0), x0_stride, _dp(us), N * self.nu, W, N,
                                _dp(xs) if return_xs else None, _dp(xf), int(nthreads))
        if not return_xs:
            return xf[0] if single else xf
        return xs[0] if single else xs

    def count_flops_step(self, qpos, qvel, ctrl, qacc_warmstart=None, literal=False) -> int:
        qpos, qvel, ctrl = _d(qpos, self.nq), _d(qvel, self.nv), _d(ctrl, self.nu)
        warm = _d(np.zeros(self.nv) if qacc_warmstart is None else qacc_warmstart, self.nv)
        return int(lib().orc_count_flops_step(C.byref(self.h), _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), int(literal)))


def quad_cost(xs, us, Q, Qf, R, xg):
    """StaticGoalQuadraticCost.cost (cost.py:62-85) in float64 numpy; xs (...,N+1,nx), us (...,N,nu)."""
    xs, us = np.asarray(xs, dtype=np.float64), np.asarray(us, dtype=np.float64)
    e = xs[..., :-1, :] - xg
    ef = xs[..., -1, :] - xg
    run = np.einsum("...ti,ij,...tj->...", e, Q, e)
    fin = np.einsum("...i,ij,...j->...", ef, Qf, ef)
    ctl = np.einsum("...ti,ij,...tj->...", us, R, us)
    return 0.5 * (run + fin + ctl)
