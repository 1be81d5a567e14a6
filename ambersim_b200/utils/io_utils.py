"""Model loader boundary (reference ambersim/utils/io_utils.py:139-248).

    load_mj_model_from_file(filepath, force_float, solver, iterations, ls_iterations) -> MjModel
    mj_to_mjx_model_and_data(mj_model) -> (mjx.Model, mjx.Data)
    load_mjx_model_and_data_from_file(filepath, force_float) -> (mjx.Model, mjx.Data)

`mujoco` is not installable here, so MJCF files are compiled by ambersim_b200.utils.mjcf and URDF files by the same compiler behind
ambersim_b200.utils.urdf (MuJoCo's URDF import conventions + the reference's actuators-from-transmissions and equalities-from-mimics,
io_utils.py:18-121); a real `mujoco.MjModel` handed in by a caller who has the package is flattened by ambersim_b200.utils.mjmodel.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple, Union

from ambersim_b200 import mjx
from ambersim_b200.utils import mjcf
from ambersim_b200.utils._internal_utils import _check_filepath

_SOLVERS = {"cg": 1, "newton": 2}


def load_mj_model_from_file(
    filepath: Union[str, Path],
    force_float: bool = False,
    solver: Optional[Union[str, int]] = None,
    iterations: Optional[int] = None,
    ls_iterations: Optional[int] = None,
) -> mjcf.MjModel:
    """Loads a model from an MJCF or URDF path (absolute, cwd-relative or package-root-relative).

    solver: 'newton' (default, as the reference does for mujoco >= 3.0.1) or 'cg'.
    iterations / ls_iterations: override the <option> values when given.
    """
    if solver is None:
        solver_id = 2
    elif isinstance(solver, str):
        if solver.lower() not in _SOLVERS:
            raise ValueError("Solver must be one of: ['cg', 'newton']!")
        solver_id = _SOLVERS[solver.lower()]
    else:
        solver_id = int(solver)
        assert solver_id in (1, 2)
    path = _check_filepath(filepath)
    ext = str(path).rsplit(".", 1)[-1]
    if ext not in ("urdf", "xml"):
        raise NotImplementedError
    m = mjcf.compile_urdf(path, force_float=force_float) if ext == "urdf" else mjcf.compile_mjcf(path, force_float=force_float)
    kw = dict(solver=solver_id)
    if iterations is not None:
        kw["iterations"] = iterations
    if ls_iterations is not None:
        kw["ls_iterations"] = ls_iterations
    m.opt = m.opt.replace(**kw)
    return m


def mj_to_mjx_model_and_data(mj_model, device=None) -> Tuple[mjx.Model, mjx.Data]:
    """Converts a host model to an (mjx.Model, mjx.Data) pair (reference io_utils.py:222-241)."""
    model = mjx.device_put(mj_model)  # this repo's compiled model, or a real mujoco.MjModel (flattened by utils/mjmodel.py)
    data = mjx.make_data(model, device=device)
    return model, data


def load_mjx_model_and_data_from_file(filepath: Union[str, Path], force_float: bool = False, device=None):
    """Convenience: file -> (mjx.Model, mjx.Data)."""
    return mj_to_mjx_model_and_data(load_mj_model_from_file(filepath, force_float=force_float), device=device)
