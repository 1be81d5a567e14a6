"""JAX binding of the engine: XLA FFI custom calls over the C ABI, so the reference's JAX callers stay unchanged.

The reference drives the hot path from JAX (`jit(vmap(shoot))`, `VanillaPredictiveSampler.optimize` under `jit` / `vmap`,
`MjxEnv.pipeline_step`; tests/trajopt/test_predictive_sampler.py:56-57,78, ambersim/rl/base.py:88-96). With JAX installed,
`build_shim()` compiles csrc/xla_ffi.cc against `jax.ffi.include_dir()` into libabr_xla.so, `register()` registers its three
handlers for the CUDA platform, and the functions below are drop-ins that trace, jit and vmap:

    shoot(m, x0, us) -> xs                                    ambersim/trajopt/shooting.py:22-48
    predictive_sample(m, cost, x0, us_guess, key, S, stdev)   shooting.py:119-157
    pipeline_step(m, qpos, qvel, qacc_warmstart, time, ctrl)  rl/base.py:88-96

vmap maps onto ONE launch: the engine is natively batched, so the handlers flatten leading batch dimensions
(`vmap_method="broadcast_all"`). JAX is NOT installable in this image (no wheel in /opt/wheelhouse): importing this module
without it raises ImportError from `_jax()`, nothing here is on the torch / ctypes path, and tests/test_jax_binding.py skips.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from ambersim_b200 import _lib

_DIR = Path(__file__).resolve().parent
_SHIM = _DIR / "libabr_xla.so"
_registered = False
TARGETS = {"abr_rollout": "abr_xla_rollout", "abr_predictive_sample": "abr_xla_predictive_sample", "abr_env_step": "abr_xla_env_step"}


def _jax():
    import jax  # noqa: F401  (ImportError here = JAX not installed: use the torch / ctypes binding instead)

    return jax


def include_dir():
    """`jax.ffi.include_dir()` (older releases: jax.extend.ffi), or None without JAX."""
    try:
        jax = _jax()
    except ImportError:
        return None
    ffi = getattr(jax, "ffi", None)
    if ffi is None:
        from jax.extend import ffi  # type: ignore
    return ffi.include_dir()


def build_shim(force: bool = False) -> Path:
    """Compile csrc/xla_ffi.cc into libabr_xla.so (needs the XLA FFI headers that ship with jaxlib)."""
    inc = include_dir()
    if inc is None:
        raise ImportError("JAX is not installed: the XLA FFI shim cannot be built (the ctypes / torch binding needs no JAX)")
    src = _DIR / "csrc" / "xla_ffi.cc"
    if force or not _SHIM.exists() or _SHIM.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", f"-I{inc}", f"-I{_DIR.parent / 'include'}",
                        "-o", str(_SHIM), str(src), f"-L{_DIR}", "-labr", f"-Xlinker=-rpath,{_DIR}"], check=True)
    return _SHIM


def register() -> None:
    """Register the three handlers with XLA (idempotent)."""
    global _registered
    if _registered:
        return
    jax = _jax()
    _lib.lib()  # libabr.so first: the shim links against it
    shim = C.CDLL(str(build_shim()))
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(shim, symbol)), platform="CUDA")
    _registered = True


def _handle_attr(h) -> np.int64:
    return np.int64(h.ptr.value if hasattr(h.ptr, "value") else int(h.ptr))


def shoot(m, x0, us, cost=None):
    """`shoot` (shooting.py:22-48) as a JAX function: x0 [..., nx], us [..., N, nu] -> xs [..., N+1, nx]
    (with `cost`: also the fused quadratic cost [...])."""
    jax = _jax()
    import jax.numpy as jnp

    register()
    x0, us = jnp.asarray(x0, jnp.float32), jnp.asarray(us, jnp.float32)
    batch, N = us.shape[:-2], us.shape[-2]
    out = (jax.ShapeDtypeStruct(batch + (N + 1, m.nx), jnp.float32), jax.ShapeDtypeStruct(batch, jnp.float32))
    dev = jax.devices("gpu")[0].id
    xs, costs = jax.ffi.ffi_call("abr_rollout", out, vmap_method="broadcast_all")(
        x0, us, model=_handle_attr(m.handle(dev)), cost=np.int64(0) if cost is None else _handle_attr(cost.device_cost(dev)))
    return xs if cost is None else (xs, costs)


def predictive_sample(m, cost, x0, us_guess, key, nsamples: int, stdev: float, sample_offset: int = 0, nsamples_total: int = 0):
    """`VanillaPredictiveSampler.optimize` (shooting.py:119-157): returns (xs_star, us_star, best_idx, best_cost).
    `key`: a Python int, or concrete jax key data (the noise is the engine's counter-based generator keyed by it)."""
    jax = _jax()
    import jax.numpy as jnp

    from ambersim_b200.trajopt.shooting import _seed_of

    register()
    x0, ug = jnp.asarray(x0, jnp.float32), jnp.asarray(us_guess, jnp.float32)
    batch, N, nu = ug.shape[:-2], ug.shape[-2], ug.shape[-1]
    out = (jax.ShapeDtypeStruct(batch + (N + 1, m.nx), jnp.float32), jax.ShapeDtypeStruct(batch + (N, nu), jnp.float32),
           jax.ShapeDtypeStruct(batch, jnp.int32), jax.ShapeDtypeStruct(batch, jnp.float32))
    dev = jax.devices("gpu")[0].id
    seed = _seed_of(np.asarray(jax.random.key_data(key)) if hasattr(key, "dtype") and jax.dtypes.issubdtype(key.dtype, jax.dtypes.prng_key) else key)
    return jax.ffi.ffi_call("abr_predictive_sample", out, vmap_method="broadcast_all")(
        x0, ug, model=_handle_attr(m.handle(dev)), cost=_handle_attr(cost.device_cost(dev)), seed=np.int64(seed & 0x7FFFFFFFFFFFFFFF),
        nsamples=np.int32(nsamples), stdev=np.float32(stdev), sample_offset=np.int32(sample_offset), nsamples_total=np.int32(nsamples_total))


def pipeline_step(m, qpos, qvel, qacc_warmstart, time, ctrl, nsubsteps: int = 1):
    """`MjxEnv.pipeline_step` (rl/base.py:88-96) on E envs: returns the stepped (qpos, qvel, qacc_warmstart, time).
    The state buffers are donated to the outputs (input_output_aliases), so the step runs in place."""
    jax = _jax()
    import jax.numpy as jnp

    register()
    args = [jnp.asarray(a, jnp.float32) for a in (qpos, qvel, qacc_warmstart, time, ctrl)]
    out = tuple(jax.ShapeDtypeStruct(a.shape, jnp.float32) for a in args[:4])
    dev = jax.devices("gpu")[0].id
    return jax.ffi.ffi_call("abr_env_step", out, vmap_method="broadcast_all", input_output_aliases={0: 0, 1: 1, 2: 2, 3: 3})(
        *args, model=_handle_attr(m.handle(dev)), nsubsteps=np.int32(nsubsteps))
