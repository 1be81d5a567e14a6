"""Drop-in for the slice of `mujoco.mjx` that ambersim's hot path uses.

    mjx.device_put(mj_model) -> Model     (reference call sites: io_utils.py:225, rl/base.py:52)
    mjx.make_data(model)     -> Data      (io_utils.py:226, shooting.py:34)
    mjx.forward(model, data) -> Data      (shooting.py:36, rl/base.py:85)
    mjx.step(model, data)    -> Data      (shooting.py:41, rl/base.py:93)

`Model` keeps the `mjx.Model` field names (numpy on the host, a device-resident blob behind the
C ABI); `Data` keeps the `mjx.Data` field names for the state that persists across steps, as torch
CUDA tensors (JAX is not present in this image; torch is the array/stream plumbing). All physics
runs in the CUDA engine (libabr.so); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum
from typing import Optional

import numpy as np
import torch

from ambersim_b200 import _abi, _lib
from ambersim_b200.utils.mjcf import MjModel, Option


class DisableBit(enum.IntFlag):
    """mujoco.mjx._src.types.DisableBit (used by tests/trajopt/test_predictive_sampler.py:7,29)."""

    CONSTRAINT = 1
    EQUALITY = 2
    FRICTIONLOSS = 4
    LIMIT = 8
    CONTACT = 16
    PASSIVE = 32
    GRAVITY = 64
    CLAMPCTRL = 128
    WARMSTART = 256
    FILTERPARENT = 512
    ACTUATION = 1024
    REFSAFE = 2048
    SENSOR = 4096
    EULERDAMP = 16384


class IntegratorType(enum.IntEnum):
    EULER = 0
    RK4 = 1


class SolverType(enum.IntEnum):
    CG = 1
    NEWTON = 2


_MODEL_FIELDS = [
    "nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "neq", "npair", "nvert", "nface", "nfacevert", "nedge",
    "body_parentid", "body_rootid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_pos",
    "body_quat", "body_ipos", "body_iquat", "body_mass", "body_subtreemass", "body_inertia", "body_invweight0",
    "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited", "jnt_solref", "jnt_solimp", "jnt_pos",
    "jnt_axis", "jnt_stiffness", "jnt_range", "jnt_margin",
    "dof_bodyid", "dof_jntid", "dof_parentid", "dof_armature", "dof_damping", "dof_invweight0",
    "geom_type", "geom_bodyid", "geom_size", "geom_pos", "geom_quat", "geom_vertadr", "geom_vertnum", "vert",
    "geom_faceadr", "geom_facenum", "face_vertadr", "face_vertnum", "face_vert", "face_normal", "geom_edgeadr", "geom_edgenum", "edge_vert",
    "pair_geom1", "pair_geom2", "pair_kind", "pair_condim", "pair_friction", "pair_solref", "pair_solimp",
    "pair_includemargin",
    "eq_type", "eq_obj1id", "eq_obj2id", "eq_active", "eq_solref", "eq_solimp", "eq_data",
    "actuator_trnid", "actuator_gaintype", "actuator_biastype", "actuator_ctrllimited", "actuator_forcelimited",
    "actuator_ctrlrange", "actuator_forcerange", "actuator_gainprm", "actuator_biasprm", "actuator_gear",
    "qpos0", "qpos_spring",
]


class _Handle:
    """Owns one AbrModel* (device-resident blob) and frees it with the Python object."""

    def __init__(self, model: "Model", device: int):
        L = _lib.lib()
        host, keep = _abi.pack_model(model, model.opt)
        ptr = C.c_void_p()
        _lib.check(L.abr_model_create(C.byref(host), device, C.byref(ptr)))
        self.ptr = ptr
        self.device = device
        info = [C.c_int() for _ in range(5)]
        _lib.check(L.abr_model_info(ptr, *[C.byref(i) for i in info]))
        self.ncon, self.ne, self.nl, self.nefc, self.depth = (i.value for i in info)
        del keep

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().abr_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Model:
    """Mirror of `mjx.Model`: same field names, numpy host arrays + a device blob created lazily."""

    def __init__(self, mj_model: MjModel, opt: Optional[Option] = None):
        for f in _MODEL_FIELDS:
            setattr(self, f, getattr(mj_model, f))
        self.opt = mj_model.opt if opt is None else opt
        self.stat = mj_model.stat
        self.names = mj_model.names
        self.keyframes = mj_model.keyframes
        self.n_unsupported_pairs = mj_model.n_unsupported_pairs
        self.unsupported_reason = mj_model.unsupported_reason
        self._mj = mj_model
        self._handles = {}
        self._lanes = 0

    @property
    def nx(self) -> int:
        return self.nq + self.nv

    def replace(self, **kw) -> "Model":
        """`model.replace(opt=model.opt.replace(...))` (tests/trajopt/test_predictive_sampler.py:22-31)."""
        m = Model(self._mj, opt=kw.pop("opt", self.opt))
        m._lanes = self._lanes
        for f in _MODEL_FIELDS:  # start from THIS model's fields, so chained replaces keep earlier overrides
            setattr(m, f, getattr(self, f))
        m.stat = self.stat
        for k, v in kw.items():
            if k not in _MODEL_FIELDS:
                raise AttributeError(f"Model has no field {k!r}")
            setattr(m, k, v)
        return m

    def set_lanes(self, lanes: int) -> "Model":
        """Pin the kernel family: 0 = auto (limb kernels when the model is eligible, else the generic kernels with
        an automatic group size), 1 = limb kernels (error if ineligible), 4/8/16/32 = generic kernels with that
        many lanes per world. Tuning knob, not part of mjx."""
        self._lanes = lanes
        for h in self._handles.values():
            _lib.check(_lib.lib().abr_model_set_lanes(h.ptr, lanes))
        return self

    def describe(self, device: Optional[int] = None) -> str:
        """Which kernel family / variant serves this model with its current options (engine extension)."""
        import ctypes

        buf = ctypes.create_string_buffer(256)
        _lib.check(_lib.lib().abr_model_describe(self.handle(device).ptr, buf, 256))
        return buf.value.decode()

    def handle(self, device: Optional[int] = None) -> _Handle:
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        h = self._handles.get(device)
        if h is None:
            contact_on = not (self.opt.disableflags & (DisableBit.CONTACT | DisableBit.CONSTRAINT))
            if self.n_unsupported_pairs and contact_on:
                raise NotImplementedError(
                    f"{self.n_unsupported_pairs} colliding geom pair(s) are outside the engine's collision "
                    f"functions ({self.unsupported_reason}); disable contacts or change the geoms")
            h = _Handle(self, device)
            if self._lanes:
                _lib.check(_lib.lib().abr_model_set_lanes(h.ptr, self._lanes))
            self._handles[device] = h
        return h


@dataclasses.dataclass
class Data:
    """Mirror of the `mjx.Data` fields that persist across steps, torch tensors.

    Leading batch dimensions are allowed on every field (the reference vmaps over them).
    """

    qpos: torch.Tensor
    qvel: torch.Tensor
    ctrl: torch.Tensor
    qacc: torch.Tensor
    qacc_warmstart: torch.Tensor
    time: torch.Tensor
    # derived fields (mjx.Data names and layouts), filled on request by forward(..., fields=...) / step(..., fields=...):
    # what an env's compute_obs / compute_reward reads (rl/base.py:98-125). None until requested.
    xpos: Optional[torch.Tensor] = None
    xquat: Optional[torch.Tensor] = None
    xipos: Optional[torch.Tensor] = None
    xanchor: Optional[torch.Tensor] = None
    xaxis: Optional[torch.Tensor] = None
    cinert: Optional[torch.Tensor] = None
    cdof: Optional[torch.Tensor] = None
    cvel: Optional[torch.Tensor] = None
    cdof_dot: Optional[torch.Tensor] = None
    qfrc_smooth: Optional[torch.Tensor] = None
    qacc_smooth: Optional[torch.Tensor] = None
    qfrc_constraint: Optional[torch.Tensor] = None
    efc_force: Optional[torch.Tensor] = None
    efc_D: Optional[torch.Tensor] = None
    efc_aref: Optional[torch.Tensor] = None
    contact_dist: Optional[torch.Tensor] = None
    contact_pos: Optional[torch.Tensor] = None
    contact_frame: Optional[torch.Tensor] = None

    def replace(self, **kw) -> "Data":
        return dataclasses.replace(self, **kw)


def _register_dataclass_pytree(cls) -> None:
    """Makes a dataclass a pytree node for `torch.utils._pytree` (tree_map / tree_flatten), the role flax.struct plays for
    mjx.Data and brax's State in the reference (rl/base.py:14-32): wrappers written as `tree_map(where_done, first, state)` work."""
    from torch.utils import _pytree

    names = [f.name for f in dataclasses.fields(cls)]

    def flatten(x):  # fields that are None (derived fields not requested) are structure, not leaves, as in JAX
        present = tuple(n for n in names if getattr(x, n) is not None)
        return [getattr(x, n) for n in present], present

    def unflatten(values, present):
        return cls(**dict(zip(present, values)))

    try:
        _pytree.register_pytree_node(cls, flatten, unflatten, serialized_type_name=f"{cls.__module__}.{cls.__name__}")
    except (ValueError, TypeError):  # registered already (module reloaded) or an older torch without the keyword
        try:
            _pytree.register_pytree_node(cls, flatten, unflatten)
        except ValueError:
            pass


_register_dataclass_pytree(Data)

DERIVED_FIELDS = ("xpos", "xquat", "xipos", "xanchor", "xaxis", "cinert", "cdof", "cvel", "cdof_dot", "qfrc_smooth", "qacc_smooth",
                  "qfrc_constraint", "efc_force", "efc_D", "efc_aref", "contact_dist", "contact_pos", "contact_frame")


def _field_shapes(m: "Model", h) -> dict:
    nb, nv, nj, ne, nc = m.nbody, m.nv, m.njnt, h.nefc, h.ncon
    return dict(xpos=(nb, 3), xquat=(nb, 4), xipos=(nb, 3), xanchor=(nj, 3), xaxis=(nj, 3), cinert=(nb, 10), cdof=(nv, 6), cvel=(nb, 6),
                cdof_dot=(nv, 6), qfrc_smooth=(nv,), qacc_smooth=(nv,), qfrc_constraint=(nv,), efc_force=(ne,), efc_D=(ne,), efc_aref=(ne,),
                contact_dist=(nc,), contact_pos=(nc, 3), contact_frame=(nc, 3, 3))


def _alloc_fields(m: "Model", h, fields, E: int, dev):
    """(AbrDataFields struct, {name: tensor [E, ...]}) for the requested derived fields."""
    S = _abi.structs()["AbrDataFields"]()
    shapes = _field_shapes(m, h)
    out = {}
    for name in fields:
        if name not in shapes:
            raise AttributeError(f"Data has no derived field {name!r} (available: {', '.join(DERIVED_FIELDS)})")
        t = torch.empty((E,) + shapes[name], dtype=torch.float32, device=dev)
        out[name] = t
        setattr(S, name, C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float)))
    return S, out


def limb_plan(mj_model: MjModel, opt: Optional[Option] = None) -> dict:
    """The limb-path decomposition the engine derives for a model (host-only, no device needed):
    eligibility, lanes per world, the serving kernel's (NL, NC), the sharing pattern and, per lane, the
    body at each chain position with its owner / sharing-level bits. Debugging aid, not part of mjx."""
    import ctypes as C

    host, _keep = _abi.pack_model(mj_model, opt)
    info = (C.c_int * 8)()
    body = (C.c_int * (8 * 17))()
    own, lvl = (C.c_int * 8)(), (C.c_int * 8)()
    _lib.check(_lib.lib().abr_limb_plan_host(C.byref(host), info, body, len(body), own, lvl))
    out = dict(eligible=bool(info[0]), lanes=1 << info[1], NL=info[2], NC=info[3], pattern=info[4], nefc=info[5], ncon=info[6],
               lanes_used=info[7], paths=[], own=[], level=[])
    if out["eligible"]:
        n = info[2] + 1
        for g in range(out["lanes"]):
            out["paths"].append([body[g * n + p] for p in range(n)])
            out["own"].append([(own[g] >> p) & 1 for p in range(n)])
            out["level"].append([(lvl[g] >> (2 * p)) & 3 for p in range(n)])
    return out


def device_put(mj_model: MjModel) -> Model:
    """mjx.device_put(mj_model): flatten the MjModel into the structure-of-arrays the engine uploads."""
    if isinstance(mj_model, Model):
        return mj_model
    from ambersim_b200.utils import mjmodel

    if mjmodel.looks_like_mjmodel(mj_model):  # a real mujoco.MjModel (io_utils.py:225, rl/base.py:52): flatten it first
        mj_model = mjmodel.from_mjmodel(mj_model)
    return Model(mj_model)


def _dev(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("ambersim_b200 needs a CUDA device: the engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def make_data(m: Model, device=None) -> Data:
    """mjx.make_data(m): zeros, qpos = qpos0."""
    dev = _dev(device)
    f = dict(dtype=torch.float32, device=dev)
    return Data(
        qpos=torch.tensor(np.asarray(m.qpos0), **f), qvel=torch.zeros(m.nv, **f), ctrl=torch.zeros(m.nu, **f),
        qacc=torch.zeros(m.nv, **f), qacc_warmstart=torch.zeros(m.nv, **f), time=torch.zeros((), **f),
    )


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev: torch.device):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _flat(t: torch.Tensor, n: int, dev: torch.device, batch: tuple) -> torch.Tensor:
    t = torch.as_tensor(t, dtype=torch.float32, device=dev)
    rows = int(np.prod(batch)) if batch else 1  # explicit: reshape(-1, 0) is ambiguous for a model without actuators (nu = 0)
    return t.expand(*batch, n).reshape(rows, n).contiguous().clone()


def _batch_shape(m: Model, d: Data) -> tuple:
    b1 = tuple(d.qpos.shape[:-1])
    b2 = tuple(torch.as_tensor(d.ctrl).shape[:-1])
    return b1 if len(b1) >= len(b2) else b2


def forward(m: Model, d: Data, fields=()) -> Data:
    """mjx.forward(m, d): fills qacc and qacc_warmstart; qpos gets its quaternions normalised.
    fields: names of derived mjx.Data fields (DERIVED_FIELDS) to fill as well, batched, in the same launch."""
    dev = d.qpos.device
    h = m.handle(dev.index or 0)
    batch = _batch_shape(m, d)
    qpos, qvel = _flat(d.qpos, m.nq, dev, batch), _flat(d.qvel, m.nv, dev, batch)
    ctrl, warm = _flat(d.ctrl, m.nu, dev, batch), _flat(d.qacc_warmstart, m.nv, dev, batch)
    qacc = torch.empty_like(qvel)
    E = qpos.shape[0]
    extra = {}
    if fields:
        S, extra = _alloc_fields(m, h, fields, E, dev)
        _lib.check(_lib.lib().abr_forward_fields_dev(h.ptr, _ptr(qpos), _ptr(qvel), _ptr(ctrl), _ptr(warm), _ptr(qacc), E, C.byref(S), _stream(dev)))
    else:
        _lib.check(_lib.lib().abr_forward_dev(h.ptr, _ptr(qpos), _ptr(qvel), _ptr(ctrl), _ptr(warm), _ptr(qacc), E, _stream(dev)))
    rs = lambda t, n: t.reshape(*batch, n)
    return d.replace(qpos=rs(qpos, m.nq), qvel=rs(qvel, m.nv), ctrl=rs(ctrl, m.nu), qacc=rs(qacc, m.nv),
                     qacc_warmstart=rs(warm, m.nv), **{k: v.reshape(*batch, *v.shape[1:]) for k, v in extra.items()})


def step(m: Model, d: Data, nsubsteps: int = 1, fields=()) -> Data:
    """mjx.step(m, d) (x nsubsteps with ctrl held: MjxEnv.pipeline_step, rl/base.py:88-96).
    fields: derived mjx.Data fields to fill as well: those of the last forward pass, i.e. of the state before the final
    integration, exactly what mjx.step leaves in mjx.Data."""
    dev = d.qpos.device
    h = m.handle(dev.index or 0)
    batch = _batch_shape(m, d)
    qpos, qvel = _flat(d.qpos, m.nq, dev, batch), _flat(d.qvel, m.nv, dev, batch)
    ctrl, warm = _flat(d.ctrl, m.nu, dev, batch), _flat(d.qacc_warmstart, m.nv, dev, batch)
    time = torch.as_tensor(d.time, dtype=torch.float32, device=dev).broadcast_to(batch).reshape(-1).contiguous().clone()
    E = qpos.shape[0]
    extra = {}
    if fields:
        S, extra = _alloc_fields(m, h, fields, E, dev)
        _lib.check(_lib.lib().abr_env_step_fields_dev(h.ptr, _ptr(qpos), _ptr(qvel), _ptr(warm), _ptr(time), _ptr(ctrl), E, int(nsubsteps),
                                                      C.byref(S), _stream(dev)))
    else:
        _lib.check(_lib.lib().abr_env_step_dev(h.ptr, _ptr(qpos), _ptr(qvel), _ptr(warm), _ptr(time), _ptr(ctrl), E,
                                               int(nsubsteps), None, None, None, None, _stream(dev)))
    rs = lambda t, n: t.reshape(*batch, n)
    return d.replace(qpos=rs(qpos, m.nq), qvel=rs(qvel, m.nv), ctrl=rs(ctrl, m.nu), qacc_warmstart=rs(warm, m.nv),
                     time=time.reshape(batch), **{k: v.reshape(*batch, *v.shape[1:]) for k, v in extra.items()})


def set_randomization(m: Model, dr: Optional[torch.Tensor]) -> None:
    """Per-env domain randomisation for the env entry points (forward / step with a leading batch of E envs):
    dr (E, 2) = {contact friction scale, actuator strength scale} or (E, 4) = those + {joint damping scale, joint armature scale}
    on the device, or None to switch it off. Like `model.replace(dof_damping=..., dof_armature=...)` in MJX the scales leave the
    compiled constants (invweight0, meaninertia) of the base model untouched. The tensor is referenced, not copied: it is kept
    alive on the model until replaced."""
    if dr is None:
        for h in m._handles.values():
            _lib.check(_lib.lib().abr_env_set_randomization(h.ptr, None, 0))
        m._dr = None
        return
    dr = dr.to(torch.float32).contiguous()
    assert dr.dim() == 2 and dr.shape[1] in (2, 4) and dr.is_cuda
    h = m.handle(dr.device.index or 0)
    _lib.check(_lib.lib().abr_env_set_randomization_ex(h.ptr, _ptr(dr), dr.shape[0], dr.shape[1]))
    m._dr = dr


def debug_forward(m: Model, qpos, qvel, ctrl=None, qacc_warmstart=None, names=()):
    """Stage dump for parity tests: one world through mjx.forward, named intermediates as numpy."""
    h = m.handle()
    L = _lib.lib()
    fp = C.POINTER(C.c_float)
    f32 = lambda a, n: np.ascontiguousarray(np.zeros(n) if a is None else a, dtype=np.float32)
    q, v = f32(qpos, m.nq), f32(qvel, m.nv)
    c, w = f32(ctrl, m.nu), f32(qacc_warmstart, m.nv)
    out = {}
    cap = max(64, h.nefc * m.nv + 16, m.nv * m.nv, 16 * m.nbody)
    for name in names:
        buf = np.zeros(cap, dtype=np.float32)
        n = C.c_int()
        _lib.check(L.abr_debug_forward_host(h.ptr, q.ctypes.data_as(fp), v.ctypes.data_as(fp), c.ctypes.data_as(fp),
                                            w.ctypes.data_as(fp), name.encode(), buf.ctypes.data_as(fp), cap, C.byref(n)))
        out[name] = buf[: n.value].copy()
    return out
