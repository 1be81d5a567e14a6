"""The C-ABI library loads, exports every symbol include/abr.h declares, agrees with the ctypes
mirror on struct layout, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C

import pytest
import torch

from ambersim_b200 import _abi, _lib, mjx
from ambersim_b200.utils.io_utils import load_mj_model_from_file


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _abi.declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/abr.h but not exported by libabr.so"
    assert L.abr_version() == 100


def test_struct_layout_matches_header():
    L = _lib.lib()
    S = _abi.structs()
    assert L.abr_sizeof_model_host() == C.sizeof(S["AbrModelHost"])
    assert L.abr_sizeof_opt() == C.sizeof(S["AbrOpt"])
    m = load_mj_model_from_file("models/barkour_standin/barkour_vb_standin.xml")
    h, keep = _abi.pack_model(m)
    assert (h.nq, h.nv, h.nu, h.nbody, h.npair) == (19, 18, 12, 14, 4)
    assert abs(h.opt.timestep - 0.004) < 1e-9 and h.opt.disableflags == 16384
    assert h.body_parentid[1] == 0 and abs(h.qpos0[2] - 0.3573) < 1e-6


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    m = mjx.device_put(load_mj_model_from_file("models/pendulum/scene.xml"))
    with pytest.raises(_lib.AbrError) as e:
        m.handle(0)
    assert e.value.code == _lib.ABR_ENODEVICE
    with pytest.raises(RuntimeError):
        mjx.make_data(m)


def test_product_does_not_import_the_oracle():
    import pathlib
    import re

    root = pathlib.Path(_lib.__file__).resolve().parent
    for f in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cuh")) + list(root.rglob("*.h")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
        assert "abr_oracle" not in text, f


def test_argument_validation_needs_no_device():
    """Bad arguments are refused with ABR_EINVAL before any CUDA call (error codes + abr_last_error, never an exception or a crash)."""
    L = _lib.lib()
    assert L.abr_rollout_dev(None, None, 0, None, 0, 4, 4, None, None, None, None) == _lib.ABR_EINVAL
    assert b"abr_rollout_dev" in L.abr_last_error()
    assert L.abr_rollout_dev(None, None, 0, None, 0, -1, 4, None, None, None, None) == _lib.ABR_EINVAL
    assert L.abr_rollout_host(None, None, 0, None, 0, 4, 4, None, None, None) == _lib.ABR_EINVAL
    assert L.abr_forward_dev(None, None, None, None, None, None, 3, None) == _lib.ABR_EINVAL
    assert L.abr_env_step_dev(None, None, None, None, None, None, 3, 1, None, None, None, None, None) == _lib.ABR_EINVAL
    assert L.abr_env_task_step_dev(None, None, None, None, None, None, 3, 1, None, None, None, None, 0.0, 10, None, None, None, None, None,
                                   None) == _lib.ABR_EINVAL
    assert L.abr_env_task_step_dev(None, None, None, None, None, None, 3, 0, None, None, None, None, 0.0, 10, None, None, None, None, None,
                                   None) == _lib.ABR_EINVAL  # nsubsteps < 1
    assert L.abr_env_set_randomization(None, None, 0) == _lib.ABR_EINVAL
    assert L.abr_model_reserve(None, 4, 4, 1, 8) == _lib.ABR_EINVAL
    assert L.abr_model_describe(None, C.create_string_buffer(8), 8) == _lib.ABR_EINVAL
    assert L.abr_predictive_sample_dev(None, None, None, None, None, 0, 1, 8, 4, 0.1, 0, 8, None, None, None, None, None, None) == _lib.ABR_EINVAL
    assert L.abr_mpc_dev(None, None, None, None, 0, 8, 4, 0.1, 2, None, None, None, None, None) == _lib.ABR_EINVAL
    handle = C.create_string_buffer(_lib.xchg_handle_bytes())
    x = C.c_void_p()
    for nranks, rank, cap in ((0, 0, 16), (9, 0, 16), (2, 2, 16), (2, -1, 16), (2, 0, 0)):
        assert L.abr_xchg_create(0, nranks, rank, cap, C.byref(x), handle) == _lib.ABR_EINVAL
    assert L.abr_xchg_create(0, 2, 0, 16, None, handle) == _lib.ABR_EINVAL
    assert L.abr_xchg_connect(None, handle) == _lib.ABR_EINVAL
    assert L.abr_xchg_merge_best_dev(None, None, None, None, None, 1, 4, 2, None, None, None, None, None) == _lib.ABR_EINVAL
    assert L.abr_xchg_destroy(None) == _lib.ABR_OK and L.abr_model_destroy(None) == _lib.ABR_OK and L.abr_cost_destroy(None) == _lib.ABR_OK


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_exchange_refuses_without_a_device():
    L = _lib.lib()
    handle = C.create_string_buffer(_lib.xchg_handle_bytes())
    x = C.c_void_p()
    assert L.abr_xchg_create(0, 2, 0, 16, C.byref(x), handle) == _lib.ABR_ENODEVICE


def _build_c_client(tmp_path):
    import subprocess

    from ambersim_b200 import _abi

    root = _abi.REPO_ROOT
    exe = tmp_path / "abr_c_client"
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", f"-I{root / 'include'}", "-o", str(exe), str(root / "tests/c_client/abr_c_client.c"),
                    f"-L{root / 'ambersim_b200'}", "-labr", "-lm", f"-Wl,-rpath,{root / 'ambersim_b200'}"], check=True, capture_output=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)


def test_plain_c_client_builds_and_fails_loudly_without_a_device(tmp_path):
    """include/abr.h is plain C99 and libabr.so links into a C program with nothing else (no Python, torch or C++ runtime on the client
    side). Without a CUDA device the program gets ABR_ENODEVICE from abr_model_create, never a fallback."""
    import torch

    r = _build_c_client(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "ABR_ENODEVICE" in r.stdout


@pytest.mark.gpu
def test_plain_c_client_rolls_a_pendulum(tmp_path):
    """The same C program on a GPU: a hand-filled AbrModelHost of a damped pendulum, 8 worlds x 200 steps through abr_rollout_host;
    finite, never above the release angle, half a period where the physical pendulum's is, identical worlds bit-identical."""
    r = _build_c_client(tmp_path)
    assert r.returncode == 0 and "0 failed checks" in r.stdout, r.stdout + r.stderr
