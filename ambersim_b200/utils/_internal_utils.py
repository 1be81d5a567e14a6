"""Path resolution rule of the reference loader (ambersim/utils/_internal_utils.py:7-19)."""
from pathlib import Path
from typing import Union

from ambersim_b200 import ROOT


def _check_filepath(filepath: Union[str, Path]) -> str:
    """Absolute path, else relative to the working directory, else relative to the package root."""
    p = Path(filepath)
    if p.is_absolute():
        if not p.exists():
            raise FileNotFoundError(f"{p} does not exist")
        return str(p)
    for cand in (Path.cwd() / p, Path(ROOT) / p):
        if cand.exists():
            return str(cand)
    raise FileNotFoundError(f"{filepath} not found (tried cwd and {ROOT})")
