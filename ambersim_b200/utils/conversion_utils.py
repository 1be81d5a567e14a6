"""URDF -> MJCF saving (the reference's ambersim/utils/conversion_utils.py:11-37; its convex-decomposition helpers wrap CoACD and are
offline asset preparation, outside this repo)."""
from __future__ import annotations

import xml.etree.ElementTree as ET
from pathlib import Path
from typing import Optional, Union

from ambersim_b200.utils._internal_utils import _check_filepath
from ambersim_b200.utils.urdf import urdf_to_mjcf


def save_model_xml(filepath: Union[str, Path], output_path: Optional[Union[str, Path]] = None) -> None:
    """Loads a URDF (or MJCF) model and saves it as an MJCF file, by default `<stem>.xml` in the working directory.

    Unlike the reference's `mj_saveLastXML` output the saved file already carries the motors made from `<transmission>`s and the joint
    equalities made from `<mimic>`s (the reference appends them in a second pass, io_utils.py:196-200): loading the saved file gives
    the model that loading the URDF gives. Mesh file names stay relative to the URDF's directory / `meshdir`."""
    path = Path(_check_filepath(filepath))
    out = Path(output_path) if output_path is not None else Path(path.name.split(".")[0] + ".xml")
    if path.suffix.lower() == ".urdf":
        root, base = urdf_to_mjcf(path)
        comp = root.find("compiler")
        meshdir = (base / comp.get("meshdir", "")).resolve()
        if root.find("asset") is not None and len(root.find("asset")):
            comp.set("meshdir", str(meshdir))  # the saved file may live elsewhere
    else:
        root = ET.parse(path).getroot()
    ET.indent(root, space="  ")
    out.write_text(ET.tostring(root, encoding="unicode") + "\n")
    print(f"XML file saved to {str(out)}!")
