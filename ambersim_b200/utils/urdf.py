"""URDF front end of the model loader (reference ambersim/utils/io_utils.py:18-121, 196-204).

The reference loads a URDF in three moves: MuJoCo's compiler imports it and saves an MJCF (`save_model_xml`), `_add_actuators`
appends one `<motor>` per `<transmission>` (ctrlrange = +- the joint's `limit effort`), `_add_mimics` appends one joint equality per
`<mimic>` (polycoef = offset, multiplier, 0, 0, 0); the result goes through `MjModel.from_xml_path`. `mujoco` is not installable
here, so `urdf_to_mjcf` restates the part of MuJoCo's URDF import the reference's models use and hands the MJCF tree to this repo's
MJCF compiler (`ambersim_b200.utils.mjcf`). What the import does (MuJoCo's `xml_urdf.cc` conventions):

* every `<link>` becomes a body named after it; the root link (never a `<child>`) hangs off the world body without a joint;
* a joint's `<origin xyz rpy>` (fixed-axis roll / pitch / yaw: R = Rz(y) Ry(p) Rx(r)) is the child body's frame in its parent; the joint
  sits at the child body's origin: `revolute` / `continuous` -> hinge, `prismatic` -> slide, `fixed` -> no joint (welded), `floating` ->
  free joint; `<axis xyz>` (default 1 0 0) in the child frame; `<limit lower upper>` is the range of revolute / prismatic joints,
  `<dynamics damping friction>` the damping / frictionloss;
* `<inertial>`: mass, centre `origin xyz`, and the tensor ixx .. izz given in the frame `origin rpy`, passed on as `fullinertia` in the
  body frame (the MJCF compiler diagonalises it, as MuJoCo does);
* `<collision>` geoms collide (default contype / conaffinity); `<visual>` geoms are kept only when `discardvisual="false"` and then never
  collide (contype = conaffinity = 0, group 1, density 0); sphere / box / cylinder / capsule / mesh geometry, mesh paths stripped to
  their file name (`strippath`, the URDF default) under `meshdir`;
* the `<mujoco><compiler .../></mujoco>` extension block carries `meshdir`, `discardvisual`, `fusestatic`, `strippath`. Bodies joined by
  fixed joints are NOT fused into their parents (MuJoCo's URDF default `fusestatic="true"` would): a welded body moves rigidly with its
  parent either way, so the dynamics are the same; only the body count differs. Both URDFs of the reference switch fusing off.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from pathlib import Path
from typing import Dict, Tuple

import numpy as np


def _vec(text, n, default):
    if text is None:
        return np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in text.split()], dtype=np.float64)
    if v.size != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return v


def _rpy_to_mat(rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _mat_to_quat(R) -> np.ndarray:
    from ambersim_b200.utils.mjcf import mat_to_quat

    return mat_to_quat(R)


def _fmt(v) -> str:
    return " ".join(f"{float(x):.12g}" for x in np.ravel(v))


def _bool(text, default):
    return default if text is None else text.strip().lower() == "true"


def _geom(elem: ET.Element, name, visual: bool, strippath: bool, meshes: Dict[str, Dict[str, str]]) -> Dict[str, str]:
    origin = elem.find("origin")
    pos = _vec(origin.get("xyz") if origin is not None else None, 3, [0, 0, 0])
    R = _rpy_to_mat(_vec(origin.get("rpy") if origin is not None else None, 3, [0, 0, 0]))
    g = elem.find("geometry")
    if g is None or len(g) == 0:
        raise ValueError("URDF visual / collision element without <geometry>")
    shape = g[0]
    a: Dict[str, str] = {"pos": _fmt(pos), "quat": _fmt(_mat_to_quat(R))}
    if name:
        a["name"] = name
    if shape.tag == "sphere":
        a.update(type="sphere", size=_fmt([float(shape.get("radius"))]))
    elif shape.tag == "box":
        a.update(type="box", size=_fmt(0.5 * _vec(shape.get("size"), 3, None)))
    elif shape.tag in ("cylinder", "capsule"):
        a.update(type=shape.tag, size=_fmt([float(shape.get("radius")), 0.5 * float(shape.get("length"))]))
    elif shape.tag == "mesh":
        fn = shape.get("filename")
        if fn is None:
            raise ValueError("URDF <mesh> without filename")
        fn = fn.split("://", 1)[-1]
        file = Path(fn).name if strippath else fn
        mname = Path(file).stem
        scale = shape.get("scale")
        key = mname if scale is None else f"{mname}_{len(meshes)}"
        if key not in meshes:
            meshes[key] = {"name": key, "file": file, **({"scale": scale} if scale else {})}
        a.update(type="mesh", mesh=key)
    else:
        raise NotImplementedError(f"URDF geometry <{shape.tag}>")
    if visual:
        a.update(contype="0", conaffinity="0", group="1", density="0")
    return a


def urdf_to_mjcf(path) -> Tuple[ET.Element, Path]:
    """Returns (the <mujoco> element tree equivalent to the URDF under the reference's loading pipeline, the directory its mesh
    paths are relative to)."""
    path = Path(path)
    # elements with an undeclared namespace prefix (e.g. <drake:declare_convex/>) are not well-formed XML; the reference reads such
    # files with lxml's recover=True (io_utils.py:28-31), here the prefix is folded into the tag name
    text = re.sub(r"<(/?)([A-Za-z_][\w.-]*):([A-Za-z_][\w.-]*)", r"<\1\2_\3", path.read_text())
    robot = ET.fromstring(text)
    if robot.tag != "robot":
        raise ValueError(f"{path} is not a URDF file")
    comp = robot.find("mujoco/compiler")
    cget = (lambda k, d=None: comp.get(k, d)) if comp is not None else (lambda k, d=None: d)
    discardvisual = _bool(cget("discardvisual"), True)
    strippath = _bool(cget("strippath"), True)
    meshdir = cget("meshdir", "")

    links = {ln.get("name"): ln for ln in robot.findall("link")}
    joints = robot.findall("joint")
    children: Dict[str, list] = {n: [] for n in links}
    is_child = set()
    for j in joints:
        p, c = j.find("parent").get("link"), j.find("child").get("link")
        if p not in links or c not in links:
            raise ValueError(f"joint {j.get('name')!r} refers to an unknown link")
        children[p].append(j)
        is_child.add(c)
    roots = [n for n in links if n not in is_child]
    if len(roots) != 1:
        raise ValueError(f"a URDF needs exactly one root link, found {roots}")

    mj = ET.Element("mujoco", {"model": robot.get("name", path.stem)})
    ET.SubElement(mj, "compiler", {"angle": "radian", "autolimits": "true", **({"meshdir": meshdir} if meshdir else {})})
    asset = ET.SubElement(mj, "asset")
    world = ET.SubElement(mj, "worldbody")
    meshes: Dict[str, Dict[str, str]] = {}

    def add_link(parent_elem: ET.Element, name: str, joint) -> None:
        attrs = {"name": name}
        if joint is not None:
            o = joint.find("origin")
            attrs["pos"] = _fmt(_vec(o.get("xyz") if o is not None else None, 3, [0, 0, 0]))
            attrs["quat"] = _fmt(_mat_to_quat(_rpy_to_mat(_vec(o.get("rpy") if o is not None else None, 3, [0, 0, 0]))))
        body = ET.SubElement(parent_elem, "body", attrs)
        ln = links[name]
        inode = ln.find("inertial")
        if inode is not None:
            o = inode.find("origin")
            R = _rpy_to_mat(_vec(o.get("rpy") if o is not None else None, 3, [0, 0, 0]))
            it = inode.find("inertia")
            I = np.array([[float(it.get("ixx", 0)), float(it.get("ixy", 0)), float(it.get("ixz", 0))],
                          [float(it.get("ixy", 0)), float(it.get("iyy", 0)), float(it.get("iyz", 0))],
                          [float(it.get("ixz", 0)), float(it.get("iyz", 0)), float(it.get("izz", 0))]])
            Ib = R @ I @ R.T  # the tensor in the body frame
            ET.SubElement(body, "inertial", {"pos": _fmt(_vec(o.get("xyz") if o is not None else None, 3, [0, 0, 0])),
                                             "mass": _fmt([float(inode.find("mass").get("value"))]),
                                             "fullinertia": _fmt([Ib[0, 0], Ib[1, 1], Ib[2, 2], Ib[0, 1], Ib[0, 2], Ib[1, 2]])})
        if joint is not None:
            jt = joint.get("type")
            if jt in ("revolute", "continuous", "prismatic"):
                ja = {"name": joint.get("name"), "type": "slide" if jt == "prismatic" else "hinge", "pos": "0 0 0"}
                ax = joint.find("axis")
                ja["axis"] = _fmt(_vec(ax.get("xyz") if ax is not None else None, 3, [1, 0, 0]))
                lim = joint.find("limit")
                if jt != "continuous" and lim is not None and lim.get("lower") is not None and lim.get("upper") is not None:
                    lo, hi = float(lim.get("lower")), float(lim.get("upper"))
                    if lo < hi:
                        ja["range"] = _fmt([lo, hi])
                dyn = joint.find("dynamics")
                if dyn is not None:
                    if dyn.get("damping") is not None:
                        ja["damping"] = dyn.get("damping")
                    if dyn.get("friction") is not None and float(dyn.get("friction")) != 0.0:
                        ja["frictionloss"] = dyn.get("friction")
                ET.SubElement(body, "joint", ja)
            elif jt == "floating":
                ET.SubElement(body, "freejoint", {"name": joint.get("name")})
            elif jt != "fixed":
                raise NotImplementedError(f"URDF joint type {jt!r}")
        for el in ln:  # geoms in document order, as MuJoCo's importer creates them
            if el.tag == "collision":
                ET.SubElement(body, "geom", _geom(el, el.get("name"), False, strippath, meshes))
            elif el.tag == "visual" and not discardvisual:
                ET.SubElement(body, "geom", _geom(el, el.get("name"), True, strippath, meshes))
        for j in children[name]:
            add_link(body, j.find("child").get("link"), j)

    add_link(world, roots[0], None)
    for m in meshes.values():
        ET.SubElement(asset, "mesh", m)

    # ambersim's _add_actuators (io_utils.py:18-69): one motor per <transmission>, limited to +- the joint's effort
    act = ET.SubElement(mj, "actuator")
    jmap = {j.get("name"): j for j in joints}
    for tr in robot.findall("transmission"):
        jn = tr.find("joint").get("name")
        lim = jmap[jn].find("limit") if jn in jmap else None
        effort = lim.get("effort") if lim is not None else None
        if effort is not None:
            ET.SubElement(act, "motor", {"name": jn + "_actuator", "ctrllimited": "true", "ctrlrange": f"-{effort} {effort}", "joint": jn})
        else:
            ET.SubElement(act, "motor", {"name": jn + "_actuator", "ctrllimited": "false", "joint": jn})
    # ambersim's _add_mimics (io_utils.py:72-121): one joint equality per <mimic>
    eq = ET.SubElement(mj, "equality")
    for j in joints:
        mim = j.find("mimic")
        if mim is not None:
            j1, j2 = j.get("name"), mim.get("joint")
            ET.SubElement(eq, "joint", {"name": f"{j1}_{j2}_equality", "joint1": j1, "joint2": j2,
                                        "polycoef": f"{mim.get('offset', '0')} {mim.get('multiplier', '1')} 0 0 0"})
    return mj, path.parent
