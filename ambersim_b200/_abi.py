"""ctypes mirror of include/abr.h.

The struct layouts are parsed from the header itself (between the ABR_STRUCT_BEGIN/END markers),
so the Python side cannot drift from the C ABI. No torch types cross this boundary: only ints,
floats and raw pointers.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

REPO_ROOT = Path(__file__).resolve().parents[1]
HEADER = REPO_ROOT / "include" / "abr.h"

_struct_cache: Dict[str, type] = {}


def _parse_structs() -> Dict[str, type]:
    if _struct_cache:
        return _struct_cache
    text = HEADER.read_text()
    for name, body in re.findall(r"/\* ABR_STRUCT_BEGIN (\w+) \*/(.*?)/\* ABR_STRUCT_END \*/", text, flags=re.S):
        inner = body[body.index("{") + 1: body.rindex("}")]
        inner = re.sub(r"/\*.*?\*/", "", inner, flags=re.S)
        fields: List[Tuple[str, object]] = []
        for decl in inner.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.fullmatch(r"(const )?(int|float|AbrOpt)(\*?) (\w+)(\[(\d+)\])?", decl)
            if not m:
                raise RuntimeError(f"cannot parse field {decl!r} of {name} in {HEADER}")
            _, base, ptr, fname, _, arr = m.groups()
            ctype = {"int": C.c_int, "float": C.c_float}.get(base) or _struct_cache[base]
            if ptr:
                ctype = C.POINTER(ctype)
            if arr:
                ctype = ctype * int(arr)
            fields.append((fname, ctype))
        _struct_cache[name] = type(name, (C.Structure,), {"_fields_": fields})
    return _struct_cache


def structs() -> Dict[str, type]:
    return _parse_structs()


def declared_functions() -> List[str]:
    """Names of every function the header declares (used by the CPU symbol-export test)."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(abr_\w+)\s*\(", text)))


def make_opt(opt, meaninertia: float):
    s = structs()["AbrOpt"]()
    s.timestep = float(opt.timestep)
    s.impratio = float(opt.impratio)
    s.tolerance = float(opt.tolerance)
    s.ls_tolerance = float(opt.ls_tolerance)
    for i in range(3):
        s.gravity[i] = float(opt.gravity[i])
    s.meaninertia = float(meaninertia)
    s.integrator = int(opt.integrator)
    s.cone = int(opt.cone)
    s.jacobian = int(opt.jacobian)
    s.solver = int(opt.solver)
    s.iterations = int(opt.iterations)
    s.ls_iterations = int(opt.ls_iterations)
    s.disableflags = int(opt.disableflags)
    return s


def pack_model(m, opt=None):
    """Flatten an `MjModel`-like object into an AbrModelHost. Returns (struct, keepalive)."""
    S = structs()["AbrModelHost"]
    h = S()
    keep = []
    opt = m.opt if opt is None else opt
    for fname, ftype in S._fields_:
        if fname == "opt":
            h.opt = make_opt(opt, m.stat.meaninertia)
        elif fname.startswith("reserved"):
            setattr(h, fname, 0)
        elif ftype is C.c_int:
            setattr(h, fname, int(getattr(m, fname)))
        else:
            is_int = ftype._type_ is C.c_int
            arr = np.ascontiguousarray(np.asarray(getattr(m, fname)), dtype=np.int32 if is_int else np.float32).ravel()
            if arr.size == 0:
                arr = np.zeros(1, dtype=arr.dtype)  # never hand C a NULL for an empty table
            keep.append(arr)
            setattr(h, fname, arr.ctypes.data_as(ftype))
    return h, keep


def pack_cost(Q, Qf, R, xg):
    S = structs()["AbrQuadCostHost"]
    c = S()
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float32)) for a in (Q, Qf, R, xg)]
    c.nx = arrs[0].shape[0]
    c.nu = arrs[2].shape[0]
    if arrs[0].shape != (c.nx, c.nx) or arrs[1].shape != (c.nx, c.nx) or arrs[2].shape != (c.nu, c.nu) or arrs[3].shape != (c.nx,):
        raise ValueError("StaticGoalQuadraticCost: Q,Qf must be (nx,nx), R (nu,nu), xg (nx,)")
    fp = C.POINTER(C.c_float)
    c.Q, c.Qf, c.R, c.xg = (a.ctypes.data_as(fp) for a in arrs)
    return c, arrs
