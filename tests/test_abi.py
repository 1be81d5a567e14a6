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
