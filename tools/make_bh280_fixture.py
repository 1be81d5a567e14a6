"""Derive the mesh-free Barrett BH280 description shipped in ambersim_b200/models/barrett_hand/.

Reads the reference's ambersim/models/barrett_hand/bh280.xml (MIT, Caltech-AMBER/ambersim) and keeps
only what the contact-free dynamics need: bodies, inertials, joints, motors and joint equalities.
The 89 convex collision meshes (4 MB of OBJ) are dropped: mesh collision is outside the engine
 and the reference's own sampler test disables contacts
(tests/trajopt/test_predictive_sampler.py:29). Run once in the build container:
    python tools/make_bh280_fixture.py /root/reference
"""
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
src = ref / "ambersim/models/barrett_hand/bh280.xml"
root = ET.parse(src).getroot()
for tag in ("asset", "statistic", "visual"):
    for n in root.findall(tag):
        root.remove(n)
comp = root.find("compiler")
comp.attrib.pop("meshdir", None)
for parent in root.iter():
    for g in [c for c in parent if c.tag == "geom"]:
        parent.remove(g)
ET.indent(root, space="  ")
out = Path(__file__).resolve().parents[1] / "ambersim_b200/models/barrett_hand/bh280.xml"
hdr = ("<!-- Barrett BH280, mesh-free: bodies/inertials/joints/motors/equalities taken from the reference's\n"
       "     ambersim/models/barrett_hand/bh280.xml (MIT, Caltech-AMBER/ambersim) by tools/make_bh280_fixture.py;\n"
       "     collision meshes dropped (mesh collision is out of the engine's scope). -->\n")
out.write_text(hdr + ET.tostring(root, encoding="unicode") + "\n")
print("wrote", out)
