"""Derive tests/models/bh280_hulls.xml: the Barrett BH280 WITH the tractable part of its real collision geometry.

Reads the reference's ambersim/models/barrett_hand/bh280.xml and its OBJ collision meshes (MIT, Caltech-AMBER/ambersim) and keeps
bodies, inertials, joints, motors, joint equalities and those collision hulls whose convex hull has at most MAXV vertices (the finger
tips' and the inner finger links' small hulls), as inline `vertex=` meshes: the GPU box has no reference tree. With MuJoCo's default
contype / conaffinity these hulls give 19 hull - hull pairs (76 contact slots). The whole hand would give ~3 200 pairs of up to 1 006
vertices each, which is why the reference's own test switches contacts off (tests/trajopt/test_predictive_sampler.py:29).
Run once in the build container:  python tools/make_bh280_hulls_fixture.py /root/reference
"""
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ambersim_b200.utils import mjcf  # noqa: E402

MAXV = 48
ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
src = ref / "ambersim/models/barrett_hand/bh280.xml"
root = ET.parse(src).getroot()
meshdir = src.parent / root.find("compiler").get("meshdir", ".")
files = {m.get("name"): (m.get("file"), m.get("scale")) for m in root.find("asset").findall("mesh")}
hulls = {}
for name, (f, scale) in files.items():
    if "_col_" not in name:
        continue
    v = mjcf.read_obj_vertices(meshdir / f)
    if scale:
        v = v * np.array([float(x) for x in scale.split()])
    h = mjcf.convex_vertices(v)
    if len(h) <= MAXV:
        hulls[name] = h
for tag in ("asset", "statistic", "visual"):
    for n in root.findall(tag):
        root.remove(n)
root.find("compiler").attrib.pop("meshdir", None)
asset = ET.Element("asset")
for name, h in sorted(hulls.items()):
    ET.SubElement(asset, "mesh", name=name, vertex=" ".join(f"{x:.9g}" for x in h.ravel()))
root.insert(1, asset)
kept = 0
for parent in root.iter():
    for g in [c for c in parent if c.tag == "geom"]:
        if g.get("mesh") in hulls and g.get("contype") != "0":
            kept += 1
        else:
            parent.remove(g)
ET.indent(root, space="  ")
out = Path(__file__).resolve().parents[1] / "tests/models/bh280_hulls.xml"
hdr = ("<!-- Barrett BH280 with the small hulls of its real collision geometry (convex hulls of at most %d vertices, inline): bodies /\n"
       "     inertials / joints / motors / equalities and the hull vertices are taken from the reference's\n"
       "     ambersim/models/barrett_hand/bh280.xml and meshes/*_col_*.obj (MIT, Caltech-AMBER/ambersim) by tools/make_bh280_hulls_fixture.py. -->\n" % MAXV)
out.write_text(hdr + ET.tostring(root, encoding="unicode") + "\n")
print("wrote", out, "with", kept, "collision geoms of", len(hulls), "distinct hulls")
