"""The XLA FFI shim (csrc/xla_ffi.cc + ambersim_b200/jax_binding.py). JAX is not installable in this image, so the
functional tests skip here; what always runs: the shim's source names exactly the handlers the binding registers, and
each handler forwards to an entry point include/abr.h declares."""
import re
from pathlib import Path

import pytest

from ambersim_b200 import _abi, jax_binding

SRC = (Path(jax_binding.__file__).parent / "csrc" / "xla_ffi.cc").read_text()


def test_shim_source_defines_the_registered_handlers():
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", SRC))
    assert defined == set(jax_binding.TARGETS.values())
    called = set(re.findall(r"\b(abr_\w+_dev)\(", SRC))
    assert called == {"abr_rollout_dev", "abr_predictive_sample_dev", "abr_env_step_dev"}
    assert called <= set(_abi.declared_functions())


def test_without_jax_the_binding_fails_loudly():
    if jax_binding.include_dir() is not None:
        pytest.skip("JAX is installed")
    with pytest.raises(ImportError):
        jax_binding.build_shim()


@pytest.mark.gpu
def test_jax_shoot_matches_torch_binding(load_model):
    jax = pytest.importorskip("jax")
    import numpy as np
    import torch

    from ambersim_b200 import mjx
    from ambersim_b200.trajopt.shooting import shoot

    mj = load_model("barkour")
    m = mjx.device_put(mj)
    rng = np.random.default_rng(0)
    x0 = np.concatenate([mj.key_qpos("home"), np.zeros(mj.nv)]).astype(np.float32)
    us = (mj.key_ctrl("home") + 0.1 * rng.standard_normal((5, 16, mj.nu))).astype(np.float32)
    ref = shoot(m, torch.tensor(x0, device="cuda").expand(5, -1).contiguous(), torch.tensor(us, device="cuda")).cpu().numpy()
    got = np.asarray(jax.jit(lambda a, b: jax_binding.shoot(m, a, b))(np.tile(x0, (5, 1)), us))
    assert np.array_equal(ref, got)
    vm = np.asarray(jax.vmap(lambda a, b: jax_binding.shoot(m, a, b))(np.tile(x0, (5, 1)), us))
    assert np.array_equal(ref, vm)
