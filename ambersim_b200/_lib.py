"""Loader for the engine's C-ABI shared library (ambersim_b200/libabr.so, built from csrc/).

The product path has no CPU fallback: if the library is missing this raises, and every compute
entry point returns ABR_ENODEVICE when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

from ambersim_b200 import _abi

_PKG = Path(__file__).resolve().parent
import os

# ABR_LIB: an alternative build of the same library (A/B experiments with compile-time knobs); the default is the in-tree one
LIB_PATH = Path(os.environ["ABR_LIB"]).resolve() if os.environ.get("ABR_LIB") else _PKG / "libabr.so"
_lib = None

ABR_OK, ABR_EINVAL, ABR_EUNSUPPORTED, ABR_ECUDA, ABR_ENODEVICE, ABR_ECAPACITY = 0, -1, -2, -3, -4, -5


class AbrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"abr error {code}: {msg}")
        self.code = code


def build(jobs: int = 8, force: bool = False) -> Path:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    args = ["make", "-C", str(_PKG / "csrc"), f"-j{jobs}"]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, capture_output=True, text=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The engine has no CPU fallback."
        )
    L = C.CDLL(str(LIB_PATH))
    S = _abi.structs()
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    L.abr_last_error.restype = C.c_char_p
    L.abr_sizeof_model_host.restype = C.c_size_t
    L.abr_sizeof_opt.restype = C.c_size_t
    L.abr_model_create.argtypes = [C.POINTER(S["AbrModelHost"]), C.c_int, C.POINTER(vp)]
    L.abr_model_destroy.argtypes = [vp]
    L.abr_model_set_opt.argtypes = [vp, C.POINTER(S["AbrOpt"])]
    L.abr_model_get_opt.argtypes = [vp, C.POINTER(S["AbrOpt"])]
    L.abr_model_info.argtypes = [vp, ip, ip, ip, ip, ip]
    L.abr_limb_plan_host.argtypes = [C.POINTER(S["AbrModelHost"]), ip, ip, C.c_int, ip, ip]
    L.abr_model_set_lanes.argtypes = [vp, C.c_int]
    L.abr_model_describe.argtypes = [vp, C.c_char_p, C.c_int]
    L.abr_model_reserve.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.abr_workspace_bytes.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    L.abr_model_set_workspace.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int]
    L.abr_cost_create.argtypes = [C.POINTER(S["AbrQuadCostHost"]), C.c_int, C.POINTER(vp)]
    L.abr_cost_destroy.argtypes = [vp]
    L.abr_rollout_dev.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.abr_rollout_host.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    ps = [vp, vp, vp, vp, vp, C.c_ulonglong, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.abr_predictive_sample_dev.argtypes = ps + [vp]
    L.abr_predictive_sample_host.argtypes = ps
    L.abr_mpc_dev.argtypes = [vp, vp, vp, vp, C.c_ulonglong, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp, vp, vp, vp]
    L.abr_forward_dev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.abr_env_step_dev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.abr_env_set_randomization.argtypes = [vp, vp, C.c_int]
    L.abr_env_set_randomization_ex.argtypes = [vp, vp, C.c_int, C.c_int]
    L.abr_forward_fields_dev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.POINTER(S["AbrDataFields"]), vp]
    L.abr_env_step_fields_dev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(S["AbrDataFields"]), vp]
    L.abr_env_task_step_dev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_float, C.c_int, vp, vp, vp, vp, vp, vp]
    L.abr_xchg_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.abr_xchg_connect.argtypes = [vp, C.c_char_p]
    L.abr_xchg_merge_best_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.abr_xchg_timed_out.argtypes = [vp, ip]
    L.abr_xchg_destroy.argtypes = [vp]
    L.abr_debug_forward_host.argtypes = [vp, fp, fp, fp, fp, C.c_char_p, fp, C.c_int, ip]
    L.abr_ffma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    if L.abr_sizeof_model_host() != C.sizeof(S["AbrModelHost"]) or L.abr_sizeof_opt() != C.sizeof(S["AbrOpt"]):
        raise ImportError("include/abr.h and libabr.so disagree on struct layout: rebuild the library")
    _lib = L
    return L


def xchg_handle_bytes() -> int:
    """ABR_XCHG_HANDLE_BYTES of include/abr.h (the header is the single source of the ABI)."""
    import re

    return int(re.search(r"#define ABR_XCHG_HANDLE_BYTES (\d+)", _abi.HEADER.read_text()).group(1))


def check(rc: int) -> None:
    if rc != ABR_OK:
        msg = lib().abr_last_error().decode()
        if rc == ABR_EUNSUPPORTED:
            raise NotImplementedError(msg)  # mirrors MJX device_put (io_utils.py:228-241)
        raise AbrError(rc, msg)
