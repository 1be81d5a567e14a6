"""Derive the URDF descriptions shipped in ambersim_b200/models/ (the reference loads its models from URDFs as well as from MJCF:
examples/load_from_file.py, tests/test_model_io.py:24-47).

Reads the reference's ambersim/models/pendulum/pendulum.urdf and ambersim/models/barrett_hand/bh280.urdf (MIT, Caltech-AMBER/ambersim)
and keeps what the dynamics need: links with their inertials, joints (origins, axes, limits, mimics), transmissions and the <mujoco>
compiler block; the Barrett hand's 98 mesh geoms are dropped like in the shipped bh280.xml (4 MB of OBJ files stay in the reference), the
pendulum keeps its collision capsule (visual geoms carry no dynamics and are dropped). Run once in the build container:  python tools/make_urdf_fixtures.py /root/reference
"""
import re
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "ambersim/models"
out = Path(__file__).resolve().parents[1] / "ambersim_b200/models"
for rel, drop_geoms in (("pendulum/pendulum.urdf", False), ("barrett_hand/bh280.urdf", True)):
    text = re.sub(r"<(/?)([A-Za-z_][\w.-]*):([A-Za-z_][\w.-]*)", r"<\1\2_\3", (ref / rel).read_text())
    root = ET.fromstring(text)
    for link in root.findall("link"):
        for el in list(link):
            if el.tag not in ("inertial", "collision") or (drop_geoms and el.tag != "inertial"):  # visuals carry no dynamics
                link.remove(el)
    comp = root.find("mujoco/compiler")
    if comp is not None and drop_geoms:
        comp.attrib.pop("meshdir", None)
    for tr in root.findall("transmission"):  # only the joint reference matters to the loader
        for el in list(tr):
            if el.tag != "joint":
                tr.remove(el)
        for j in tr.findall("joint"):
            for el in list(j):
                j.remove(el)
    ET.indent(root, space=" ")
    hdr = (f"<!-- {Path(rel).name}: links / inertials / joints / mimics / transmissions taken from the reference's ambersim/models/{rel}\n"
           f"     (MIT, Caltech-AMBER/ambersim) by tools/make_urdf_fixtures.py" + ("; mesh geometry dropped" if drop_geoms else "") + ". -->\n")
    (out / rel).write_text(hdr + ET.tostring(root, encoding="unicode") + "\n")
    print("wrote", out / rel)
