"""Name lists of a loaded model, in id order (the reference's ambersim/utils/introspection_utils.py:8-25, used by its loader tests,
tests/test_model_io.py:66-100): the host model keeps the names MuJoCo would return from mj_id2name."""
from typing import List


def _names(model, kind: str, n: int) -> List[str]:
    names = list(getattr(model, "names", {}).get(kind, []))
    if len(names) != n:
        raise ValueError(f"model carries {len(names)} {kind} names for {n} objects")
    return names


def get_actuator_names(model) -> List[str]:
    """All actuator names of a host (not device) model."""
    return _names(model, "actuator", model.nu)


def get_equality_names(model) -> List[str]:
    """All equality constraint names of a host model."""
    return _names(model, "equality", model.neq)


def get_geom_names(model) -> List[str]:
    """All geom names of a host model."""
    return _names(model, "geom", model.ngeom)


def get_joint_names(model) -> List[str]:
    """All joint names of a host model."""
    return _names(model, "joint", model.njnt)
