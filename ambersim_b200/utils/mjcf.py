"""Minimal MJCF compiler: XML -> flat `MjModel`-like description (numpy, float64/int32).

Stands in for `mujoco.MjModel.from_xml_path` (reference: ambersim/utils/io_utils.py:206), which
is a C library that is not installable in this image. It covers what the hot path needs:
kinematic trees of free/hinge/slide joints with explicit <inertial>, plane/sphere/capsule geoms,
motor/position/general joint actuators, joint equalities, <default> classes, <include>,
<keyframe>, <option>/<flag>, and the constants MuJoCo's compiler derives at qpos0
(`mj_setConst`: body_subtreemass, dof_invweight0, body_invweight0, stat.meaninertia).

One-off host work (numpy); none of it is on the rollout hot path.
"""
from __future__ import annotations

import copy
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

# enums (MuJoCo values)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
_GEOM_TYPES = {
    "plane": GEOM_PLANE, "hfield": GEOM_HFIELD, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH,
}
_JNT_TYPES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
(PAIR_PLANE_SPHERE, PAIR_PLANE_CAPSULE, PAIR_SPHERE_SPHERE, PAIR_SPHERE_CAPSULE, PAIR_CAPSULE_CAPSULE, PAIR_PLANE_CONVEX,
 PAIR_SPHERE_CONVEX, PAIR_CAPSULE_CONVEX, PAIR_CONVEX_CONVEX) = range(9)
_PAIR_KIND = {
    (GEOM_PLANE, GEOM_SPHERE): PAIR_PLANE_SPHERE,
    (GEOM_PLANE, GEOM_CAPSULE): PAIR_PLANE_CAPSULE,
    (GEOM_PLANE, GEOM_BOX): PAIR_PLANE_CONVEX,   # plane against a convex vertex set: up to 4 contacts (mjx collision_convex.plane_convex)
    (GEOM_PLANE, GEOM_MESH): PAIR_PLANE_CONVEX,
    (GEOM_SPHERE, GEOM_SPHERE): PAIR_SPHERE_SPHERE,
    (GEOM_SPHERE, GEOM_CAPSULE): PAIR_SPHERE_CAPSULE,
    (GEOM_CAPSULE, GEOM_CAPSULE): PAIR_CAPSULE_CAPSULE,
    # convex geoms against spheres, capsules and each other (mjx collision_convex.sphere_convex / capsule_convex / convex_convex)
    (GEOM_SPHERE, GEOM_BOX): PAIR_SPHERE_CONVEX,
    (GEOM_SPHERE, GEOM_MESH): PAIR_SPHERE_CONVEX,
    (GEOM_CAPSULE, GEOM_BOX): PAIR_CAPSULE_CONVEX,
    (GEOM_CAPSULE, GEOM_MESH): PAIR_CAPSULE_CONVEX,
    (GEOM_BOX, GEOM_BOX): PAIR_CONVEX_CONVEX,
    (GEOM_BOX, GEOM_MESH): PAIR_CONVEX_CONVEX,
    (GEOM_MESH, GEOM_MESH): PAIR_CONVEX_CONVEX,
}
PAIR_NCON = {PAIR_PLANE_SPHERE: 1, PAIR_PLANE_CAPSULE: 2, PAIR_SPHERE_SPHERE: 1, PAIR_SPHERE_CAPSULE: 1,
             PAIR_CAPSULE_CAPSULE: 1, PAIR_PLANE_CONVEX: 4, PAIR_SPHERE_CONVEX: 1, PAIR_CAPSULE_CONVEX: 2, PAIR_CONVEX_CONVEX: 4}
MAX_FACE_VERTS = 8   # polygon faces larger than this stay triangulated (the kernels clip into fixed-size local arrays)
MAX_CONVEX_VERTS = 64  # per geom, for the pairs that need the hull's faces and edges (sphere / capsule / convex - convex)
_DISABLE_BITS = {
    "constraint": 1, "equality": 2, "frictionloss": 4, "limit": 8, "contact": 16, "passive": 32,
    "gravity": 64, "clampctrl": 128, "warmstart": 256, "filterparent": 512, "actuation": 1024,
    "refsafe": 2048, "sensor": 4096, "eulerdamp": 16384,
}
_INTEGRATORS = {"euler": 0, "rk4": 1, "implicit": 2, "implicitfast": 3}
_SOLVERS = {"pgs": 0, "cg": 1, "newton": 2}
_CONES = {"pyramidal": 0, "elliptic": 1}
_JACOBIANS = {"dense": 0, "sparse": 1, "auto": 2}
MJ_MINVAL = 1e-15


# ---------------------------------------------------------------------------- quaternion helpers
def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def rotate(v, q):
    return quat_to_mat(q) @ v


def mat_to_quat(m):
    # robust conversion for proper rotations
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def _quat_z_to_vec(vec):
    """Quaternion rotating the z axis onto `vec` (MuJoCo's mjuu_z2quat)."""
    v = np.asarray(vec, dtype=np.float64)
    n = np.linalg.norm(v)
    if n < MJ_MINVAL:
        return np.array([1.0, 0, 0, 0])
    v = v / n
    z = np.array([0.0, 0, 1.0])
    axis = np.cross(z, v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1.0, 0, 0])
    axis = axis / s
    ang = np.arctan2(s, v[2])
    return np.concatenate([[np.cos(ang / 2)], axis * np.sin(ang / 2)])


# ---------------------------------------------------------------------------- model container
@dataclass
class Option:
    """Mirror of mujoco's `MjOption` fields used by the hot path (+ `stat.meaninertia`)."""

    timestep: float = 0.002
    impratio: float = 1.0
    tolerance: float = 1e-8
    ls_tolerance: float = 0.01
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    integrator: int = 0
    cone: int = 0
    jacobian: int = 2
    solver: int = 2
    iterations: int = 100
    ls_iterations: int = 50
    disableflags: int = 0

    def replace(self, **kw) -> "Option":
        o = copy.copy(self)
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(f"Option has no field {k!r}")
            setattr(o, k, int(v) if k in ("integrator", "cone", "jacobian", "solver", "iterations",
                                          "ls_iterations", "disableflags") else v)
        return o


@dataclass
class Statistic:
    meaninertia: float = 1.0


class MjModel:
    """Flat host-side model with MuJoCo field names (stand-in for `mujoco.MjModel`)."""

    def __init__(self) -> None:
        self.opt = Option()
        self.stat = Statistic()
        self.names: Dict[str, List[str]] = {}
        self.keyframes: Dict[str, Dict[str, np.ndarray]] = {}
        self.n_unsupported_pairs = 0
        self.unsupported_reason = ""

    # convenience
    @property
    def nx(self) -> int:
        return self.nq + self.nv

    def key_qpos(self, name: str) -> np.ndarray:
        return self.keyframes[name]["qpos"].copy()

    def key_ctrl(self, name: str) -> np.ndarray:
        return self.keyframes[name]["ctrl"].copy()


# ---------------------------------------------------------------------------- parsing helpers
def _floats(s: Optional[str], n: Optional[int] = None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(t) for t in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        if default is not None and v.size < n:  # MuJoCo pads partial vectors with defaults
            full = np.array(default, dtype=np.float64)
            full[: v.size] = v
            return full
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def _bool(s: Optional[str], default=None):
    if s is None:
        return default
    if s in ("true", "1"):
        return True
    if s in ("false", "0"):
        return False
    if s == "auto":
        return None
    raise ValueError(f"bad boolean {s!r}")


class _Defaults:
    """<default> class tree: class name -> {element tag -> attribute dict}, with inheritance."""

    def __init__(self) -> None:
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}
        self.parent: Dict[str, Optional[str]] = {"main": None}

    def load(self, node: ET.Element, parent: Optional[str] = None) -> None:
        name = node.get("class", "main" if parent is None else None)
        if name is None:
            raise ValueError("nested <default> needs a class name")
        if name not in self.classes:
            self.classes[name] = {}
            self.parent[name] = parent
        if name != "main" or parent is not None:
            self.parent.setdefault(name, parent)
        for child in node:
            if child.tag == "default":
                self.load(child, name)
            else:
                self.classes[name].setdefault(child.tag, {}).update(child.attrib)

    def resolve(self, tag: str, cls: Optional[str]) -> Dict[str, str]:
        chain = []
        c = cls if cls in self.classes else "main"
        while c is not None:
            chain.append(c)
            c = self.parent.get(c)
        out: Dict[str, str] = {}
        for c in reversed(chain):
            out.update(self.classes[c].get(tag, {}))
        return out


def convex_vertices(points: np.ndarray) -> np.ndarray:
    """Vertices of the convex hull of a point set, in their original order (MuJoCo / MJX collide a mesh geom as the convex
    hull of its vertices). Degenerate sets (fewer than 4 points, coplanar) are returned unchanged."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if len(pts) < 4:
        return pts
    try:
        from scipy.spatial import ConvexHull

        keep = np.sort(np.unique(ConvexHull(pts).vertices))
        return pts[keep]
    except Exception:
        return pts


def convex_topology(verts: np.ndarray):
    """Faces and edges of the convex hull of `verts` (all of them hull vertices): what sphere - convex, capsule - convex and
    convex - convex collision need beyond the vertex set (mjx keeps the same per-mesh tables: polygon faces, face normals, unique
    edges). Returns (faces, normals, edges): `faces` a list of vertex-index lists, counter-clockwise seen from outside, coplanar
    hull triangles merged into one polygon (up to MAX_FACE_VERTS vertices; larger polygons stay triangles), `normals` the outward
    unit normals, `edges` the unique polygon edges as (i, j) pairs with i < j. Degenerate sets (flat, fewer than 4 points) have no
    faces."""
    pts = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    if len(pts) < 4:
        return [], np.zeros((0, 3)), np.zeros((0, 2), dtype=np.int32)
    try:
        from scipy.spatial import ConvexHull

        hull = ConvexHull(pts)
    except Exception:
        return [], np.zeros((0, 3)), np.zeros((0, 2), dtype=np.int32)
    scale = max(1e-12, float(np.abs(pts).max()))
    # coplanar hull triangles -> one group per supporting plane. Candidates are bucketed on the rounded plane equation and then
    # compared exactly against the bucket's groups, so the merge stays a tolerance test but costs O(triangles)
    groups, buckets = [], {}  # group: [normal, offset, set of vertex ids, list of triangles]
    for tri, eq in zip(hull.simplices, hull.equations):
        n, d = eq[:3], eq[3]
        key0 = np.round(np.concatenate([n, [d / scale]]) * 1e6).astype(np.int64)
        hit = None
        for dk in ((0, 0, 0, 0),) + tuple(tuple(int(i == j) * sg for j in range(4)) for i in range(4) for sg in (-1, 1)):
            for gi in buckets.get(tuple(key0 + np.array(dk)), ()):
                grp = groups[gi]
                if np.dot(grp[0], n) > 1.0 - 1e-9 and abs(grp[1] - d) < 1e-9 * scale + 1e-12:
                    hit = grp
                    break
            if hit is not None:
                break
        if hit is not None:
            hit[2].update(int(i) for i in tri)
            hit[3].append(tri)
        else:
            groups.append([n.copy(), float(d), set(int(i) for i in tri), [tri]])
            buckets.setdefault(tuple(key0), []).append(len(groups) - 1)
    faces, normals = [], []

    def ordered(ids, n):
        ids = sorted(ids)
        c = pts[ids].mean(axis=0)
        u = pts[ids[0]] - c
        u = u - n * np.dot(u, n)
        u /= max(np.linalg.norm(u), 1e-300)
        w = np.cross(n, u)
        ang = [np.arctan2(np.dot(pts[i] - c, w), np.dot(pts[i] - c, u)) for i in ids]
        return [ids[k] for k in np.argsort(ang, kind="stable")]

    for n, d, ids, tris in groups:
        n = n / np.linalg.norm(n)
        if len(ids) <= MAX_FACE_VERTS:
            faces.append(ordered(ids, n))
            normals.append(n)
        else:
            for tri in tris:
                faces.append(ordered(set(int(i) for i in tri), n))
                normals.append(n)
    # canonical order (independent of qhull's facet order): by the direction of the normal, then by the lowest vertex id
    order = sorted(range(len(faces)), key=lambda f: (tuple(np.round(-normals[f], 9)), min(faces[f])))
    faces = [faces[f] for f in order]
    normals = np.array([normals[f] for f in order]).reshape(-1, 3)
    faces = [f[f.index(min(f)):] + f[:f.index(min(f))] for f in faces]  # start every polygon at its lowest vertex id
    edges = sorted({(min(a, b), max(a, b)) for f in faces for a, b in zip(f, f[1:] + f[:1])})
    if any(len(f) > MAX_FACE_VERTS for f in faces):
        raise AssertionError("polygon larger than MAX_FACE_VERTS")
    return faces, normals, np.array(edges, dtype=np.int32).reshape(-1, 2)


def attach_convex_topology(m) -> None:
    """Face / edge tables of every convex (box / mesh) geom of `m`, in the flat layout the model blob carries:
    geom_faceadr / geom_facenum -> face_vertadr / face_vertnum -> face_vert (vertex ids LOCAL to the geom's vertex set),
    face_normal (geom frame); geom_edgeadr / geom_edgenum -> edge_vert."""
    fadr, fnum, fva, fvn, fv, fn, eadr, enum_, ev = [], [], [], [], [], [], [], [], []
    for g in range(m.ngeom):
        a, n = int(m.geom_vertadr[g]), int(m.geom_vertnum[g])
        faces, normals, edges = convex_topology(m.vert[a:a + n]) if n else ([], np.zeros((0, 3)), np.zeros((0, 2), dtype=np.int32))
        fadr.append(len(fva)); fnum.append(len(faces))
        for f, nn in zip(faces, normals):
            fva.append(len(fv)); fvn.append(len(f)); fv.extend(f); fn.append(nn)
        eadr.append(len(ev)); enum_.append(len(edges)); ev.extend(edges.tolist())
    i32 = lambda x: np.array(x, dtype=np.int32)
    m.geom_faceadr, m.geom_facenum, m.face_vertadr, m.face_vertnum, m.face_vert = i32(fadr), i32(fnum), i32(fva), i32(fvn), i32(fv)
    m.face_normal = np.array(fn, dtype=np.float64).reshape(-1, 3)
    m.geom_edgeadr, m.geom_edgenum, m.edge_vert = i32(eadr), i32(enum_), i32(ev).reshape(-1, 2)
    m.nface, m.nfacevert, m.nedge = len(fva), len(fv), len(ev)


def box_vertices(size) -> np.ndarray:
    """The 8 corners of a box geom (half extents `size`), in the order mjx builds them: x slowest, z fastest."""
    s = np.asarray(size, dtype=np.float64)[:3]
    return np.array([[sx * s[0], sy * s[1], sz * s[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])


def read_obj_vertices(path) -> np.ndarray:
    """`v x y z` lines of a Wavefront OBJ file."""
    out = []
    for line in Path(path).read_text().splitlines():
        t = line.split()
        if len(t) >= 4 and t[0] == "v":
            out.append([float(t[1]), float(t[2]), float(t[3])])
    if not out:
        raise ValueError(f"{path}: no vertices")
    return np.array(out)


def read_stl_vertices(path) -> np.ndarray:
    """Vertices of an STL file, binary or ASCII (every triangle corner; the hull reduction follows)."""
    raw = Path(path).read_bytes()
    if len(raw) >= 84:
        ntri = int(np.frombuffer(raw[80:84], dtype="<u4")[0])
        if len(raw) == 84 + 50 * ntri and ntri > 0:  # binary: 80-byte header, count, 50 bytes per triangle (normal, 3 corners, attribute)
            rec = np.frombuffer(raw[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
            return np.unique(rec["v"].reshape(-1, 3).astype(np.float64), axis=0)
    out = []
    for line in raw.decode("ascii", errors="ignore").splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == "vertex":
            out.append([float(t[1]), float(t[2]), float(t[3])])
    if not out:
        raise ValueError(f"{path}: not an STL file")
    return np.unique(np.array(out), axis=0)


def _expand_includes(root: ET.Element, base: Path) -> None:
    for parent in list(root.iter()):
        for i, child in enumerate(list(parent)):
            if child.tag == "include":
                inc_path = base / child.get("file")
                inc_root = ET.parse(inc_path).getroot()
                _expand_includes(inc_root, inc_path.parent)
                idx = list(parent).index(child)
                parent.remove(child)
                for k, sub in enumerate(list(inc_root)):
                    parent.insert(idx + k, sub)


# ---------------------------------------------------------------------------- the compiler
def enumerate_pairs(m: "MjModel", G, excludes, explicit=()) -> None:
    """Static enumeration of colliding geom pairs + parameter mixing, shared by the MJCF compiler and the
    mujoco.MjModel adapter (utils/mjmodel.py). G: one dict per geom (type, body, contype, conaffinity, condim, priority,
    friction[3], solmix, solref[2], solimp[5], margin, gap, name); excludes: set of (body, body) pairs; explicit:
    pre-mixed <contact><pair> entries (dicts with g1, g2, condim, friction[5], solref, solimp, includemargin).

    Restates what MJX does at trace time: contype/conaffinity filter, same
    weld-body and parent-child filters, type-pair dispatch, friction=max, solref/solimp mixed
    by solmix, includemargin = max(margin) - max(gap), condim = max; priority overrides.
    """
    pairs = []
    unsupported = 0
    reason = ""
    filterparent = not (m.opt.disableflags & _DISABLE_BITS["filterparent"])
    for i in range(m.ngeom):
        for k in range(i + 1, m.ngeom):
            a, b = G[i], G[k]
            if not ((a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"])):
                continue
            b1, b2 = a["body"], b["body"]
            w1, w2 = m.body_weldid[b1], m.body_weldid[b2]
            if w1 == w2:
                continue
            if filterparent and w1 != 0 and w2 != 0:
                wp1 = m.body_weldid[m.body_parentid[w1]]
                wp2 = m.body_weldid[m.body_parentid[w2]]
                if wp1 == w2 or wp2 == w1:
                    continue
            if (min(b1, b2), max(b1, b2)) in excludes:
                continue
            g1, g2 = (i, k) if a["type"] <= b["type"] else (k, i)
            key = (G[g1]["type"], G[g2]["type"])
            if key not in _PAIR_KIND:
                unsupported += 1
                reason = f"geom pair types {key} ({G[g1]['name']}, {G[g2]['name']})"
                continue
            A, B = G[g1], G[g2]
            if A["priority"] != B["priority"]:
                hi = A if A["priority"] > B["priority"] else B
                friction, solref, solimp, condim = hi["friction"], hi["solref"], hi["solimp"], hi["condim"]
            else:
                s1, s2 = A["solmix"], B["solmix"]
                if s1 >= MJ_MINVAL and s2 >= MJ_MINVAL:
                    mix = s1 / (s1 + s2)
                elif s1 < MJ_MINVAL and s2 < MJ_MINVAL:
                    mix = 0.5
                elif s1 < MJ_MINVAL:
                    mix = 0.0
                else:
                    mix = 1.0
                friction = np.maximum(A["friction"], B["friction"])
                if A["solref"][0] > 0 and B["solref"][0] > 0:
                    solref = mix * A["solref"] + (1 - mix) * B["solref"]
                else:
                    solref = np.minimum(A["solref"], B["solref"])
                solimp = mix * A["solimp"] + (1 - mix) * B["solimp"]
                condim = max(A["condim"], B["condim"])
            if condim not in (1, 3):
                unsupported += 1
                reason = f"condim {condim}"
                continue
            margin = max(A["margin"], B["margin"])
            gap = max(A["gap"], B["gap"])
            pairs.append(dict(g1=g1, g2=g2, kind=_PAIR_KIND[key], condim=condim,
                              friction=np.array([friction[0], friction[0], friction[1], friction[2], friction[2]]),
                              solref=solref, solimp=solimp, includemargin=margin - gap))
    for e in explicit:
        g1, g2 = (e["g1"], e["g2"]) if G[e["g1"]]["type"] <= G[e["g2"]]["type"] else (e["g2"], e["g1"])
        key = (G[g1]["type"], G[g2]["type"])
        if key not in _PAIR_KIND or e["condim"] not in (1, 3):
            unsupported += 1
            reason = f"explicit pair of geom types {key}, condim {e['condim']}"
            continue
        pairs.append(dict(g1=g1, g2=g2, kind=_PAIR_KIND[key], condim=int(e["condim"]), friction=np.asarray(e["friction"], dtype=np.float64),
                          solref=np.asarray(e["solref"], dtype=np.float64), solimp=np.asarray(e["solimp"], dtype=np.float64),
                          includemargin=float(e["includemargin"])))
    # sphere / capsule / convex - convex pairs work on the hull's faces and edges: the geom needs a proper (3-D) hull of bounded size
    kept = []
    for pr in pairs:
        if pr["kind"] in (PAIR_SPHERE_CONVEX, PAIR_CAPSULE_CONVEX, PAIR_CONVEX_CONVEX):
            bad = [g for g in ((pr["g2"],) if pr["kind"] != PAIR_CONVEX_CONVEX else (pr["g1"], pr["g2"]))
                   if m.geom_facenum[g] == 0 or m.geom_vertnum[g] > MAX_CONVEX_VERTS]
            if bad:
                unsupported += 1
                reason = (f"convex geom {G[bad[0]]['name']} has {int(m.geom_vertnum[bad[0]])} hull vertices and {int(m.geom_facenum[bad[0]])} faces "
                          f"(needs a 3-D hull of at most {MAX_CONVEX_VERTS} vertices)")
                continue
        kept.append(pr)
    pairs = kept
    npair = len(pairs)
    m.npair = npair
    m.pair_geom1 = np.array([p["g1"] for p in pairs], dtype=np.int32)
    m.pair_geom2 = np.array([p["g2"] for p in pairs], dtype=np.int32)
    m.pair_kind = np.array([p["kind"] for p in pairs], dtype=np.int32)
    m.pair_condim = np.array([p["condim"] for p in pairs], dtype=np.int32)
    m.pair_friction = np.array([p["friction"] for p in pairs]).reshape(npair, 5)
    m.pair_solref = np.array([p["solref"] for p in pairs]).reshape(npair, 2)
    m.pair_solimp = np.array([p["solimp"] for p in pairs]).reshape(npair, 5)
    m.pair_includemargin = np.array([p["includemargin"] for p in pairs], dtype=np.float64)
    m.n_unsupported_pairs = unsupported
    m.unsupported_reason = reason
    if m.opt.cone != 0:
        m.n_unsupported_pairs += 1
        m.unsupported_reason = "elliptic friction cone"



class _Compiler:
    def __init__(self, root: ET.Element, force_float: bool = False) -> None:
        self.root = root
        self.force_float = force_float
        comp = root.find("compiler")
        self.degrees = True
        self.autolimits = True
        self.eulerseq = "xyz"
        if comp is not None:
            self.degrees = comp.get("angle", "degree") == "degree"
            self.autolimits = _bool(comp.get("autolimits"), True)
            self.eulerseq = comp.get("eulerseq", "xyz")
        self.defaults = _Defaults()
        for d in root.findall("default"):
            self.defaults.load(d)

        self.bodies: List[dict] = []
        self.joints: List[dict] = []
        self.geoms: List[dict] = []
        # <asset><mesh>: vertices inline (`vertex="x y z ..."`) or from an OBJ file under <compiler meshdir>, optional `scale`
        self.meshes: Dict[str, np.ndarray] = {}
        self.base_dir = Path(".")
        self.meshdir = comp.get("meshdir", "") if comp is not None else ""

    # -- orientation of a frame-carrying element (quat | euler | axisangle | zaxis | xyaxes)
    def _orientation(self, attrs: Dict[str, str]) -> np.ndarray:
        if "quat" in attrs:
            q = _floats(attrs["quat"], 4)
            return q / np.linalg.norm(q)
        if "euler" in attrs:
            e = _floats(attrs["euler"], 3)
            if self.degrees:
                e = np.deg2rad(e)
            q = np.array([1.0, 0, 0, 0])
            for ax, ang in zip(self.eulerseq, e):
                axis = np.zeros(3)
                axis["xyz".index(ax.lower())] = 1.0
                r = np.concatenate([[np.cos(ang / 2)], axis * np.sin(ang / 2)])
                q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
            return q
        if "axisangle" in attrs:
            a = _floats(attrs["axisangle"], 4)
            ang = np.deg2rad(a[3]) if self.degrees else a[3]
            ax = a[:3] / np.linalg.norm(a[:3])
            return np.concatenate([[np.cos(ang / 2)], ax * np.sin(ang / 2)])
        if "zaxis" in attrs:
            return _quat_z_to_vec(_floats(attrs["zaxis"], 3))
        if "xyaxes" in attrs:
            a = _floats(attrs["xyaxes"], 6)
            x = a[:3] / np.linalg.norm(a[:3])
            y = a[3:] - x * np.dot(x, a[3:])
            y = y / np.linalg.norm(y)
            return mat_to_quat(np.stack([x, y, np.cross(x, y)], axis=1))
        return np.array([1.0, 0, 0, 0])

    def _walk(self, node: ET.Element, parent_id: int, childclass: Optional[str]) -> None:
        for bnode in node.findall("body"):
            cc = bnode.get("childclass", childclass)
            bid = len(self.bodies)
            body = dict(
                name=bnode.get("name", f"body{bid}"), parent=parent_id,
                pos=_floats(bnode.get("pos"), 3, [0, 0, 0]), quat=self._orientation(bnode.attrib),
                ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3),
                has_inertial=False, joints=[],
            )
            inode = bnode.find("inertial")
            if inode is not None:
                body["has_inertial"] = True
                body["ipos"] = _floats(inode.get("pos"), 3, [0, 0, 0])
                body["mass"] = float(inode.get("mass"))
                if inode.get("fullinertia") is not None:
                    f = _floats(inode.get("fullinertia"), 6)
                    full = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                    w, v = np.linalg.eigh(full)
                    order = np.argsort(-w)
                    w, v = w[order], v[:, order]
                    if np.linalg.det(v) < 0:
                        v[:, 2] = -v[:, 2]
                    body["inertia"] = w
                    body["iquat"] = mat_to_quat(v)
                else:
                    body["inertia"] = _floats(inode.get("diaginertia"), 3)
                    body["iquat"] = self._orientation(inode.attrib)
            self.bodies.append(body)

            jnodes = [c for c in bnode if c.tag in ("joint", "freejoint")]
            if self.force_float and parent_id == 0 and not any(b["parent"] == 0 for b in self.bodies[1:-1]):
                # reference: _modify_robot_float_base adds a <freejoint> to the first body
                # (ambersim/utils/io_utils.py:120-136)
                if not any(j.tag == "freejoint" or j.get("type") == "free" for j in jnodes):
                    jnodes = [ET.Element("freejoint", {"name": "freejoint"})] + jnodes
            for jn in jnodes:
                if jn.tag == "freejoint":
                    attrs = {"type": "free"}
                    attrs.update(jn.attrib)
                else:
                    attrs = self.defaults.resolve("joint", jn.get("class", cc))
                    attrs.update(jn.attrib)
                jtype = _JNT_TYPES[attrs.get("type", "hinge")]
                rng = _floats(attrs.get("range"), 2, [0, 0])
                if self.degrees and jtype == JNT_HINGE:
                    rng = np.deg2rad(rng)
                limited = _bool(attrs.get("limited"))
                if limited is None:
                    limited = self.autolimits and ("range" in attrs)
                axis = _floats(attrs.get("axis"), 3, [0, 0, 1])
                if jtype in (JNT_HINGE, JNT_SLIDE):
                    axis = axis / np.linalg.norm(axis)
                ref = float(attrs.get("ref", 0.0))
                springref = float(attrs.get("springref", 0.0))
                if self.degrees and jtype == JNT_HINGE:
                    ref, springref = np.deg2rad(ref), np.deg2rad(springref)
                if float(attrs.get("frictionloss", 0.0)) != 0.0:
                    raise NotImplementedError("joint frictionloss is not supported (as in MJX 3.0.x)")
                j = dict(
                    name=attrs.get("name", f"joint{len(self.joints)}"), type=jtype, body=bid,
                    pos=_floats(attrs.get("pos"), 3, [0, 0, 0]), axis=axis, range=rng, limited=bool(limited),
                    stiffness=float(attrs.get("stiffness", 0.0)), damping=float(attrs.get("damping", 0.0)),
                    armature=float(attrs.get("armature", 0.0)), margin=float(attrs.get("margin", 0.0)),
                    ref=ref, springref=springref,
                    solref=_floats(attrs.get("solreflimit"), 2, [0.02, 1.0]),
                    solimp=_floats(attrs.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                )
                if jtype == JNT_FREE:
                    j["pos"] = np.zeros(3)
                    j["axis"] = np.array([0.0, 0, 1.0])
                    j["limited"] = False
                body["joints"].append(len(self.joints))
                self.joints.append(j)

            self._geoms(bnode, bid, cc)
            if not body["has_inertial"] and body["joints"]:
                raise NotImplementedError(
                    f"body {body['name']!r} moves but has no <inertial>; inertia-from-geom is not supported")
            self._walk(bnode, bid, cc)

    def _load_meshes(self) -> None:
        for anode in self.root.findall("asset"):
            for mn in anode.findall("mesh"):
                attrs = self.defaults.resolve("mesh", mn.get("class"))
                attrs.update(mn.attrib)
                if "vertex" in attrs:
                    v = _floats(attrs["vertex"]).reshape(-1, 3)
                elif "file" in attrs:
                    f = self.base_dir / self.meshdir / attrs["file"]
                    if f.suffix.lower() == ".obj":
                        v = read_obj_vertices(f)
                    elif f.suffix.lower() == ".stl":
                        v = read_stl_vertices(f)
                    else:
                        raise NotImplementedError(f"mesh file {f.name}: Wavefront OBJ and STL are read")
                else:
                    raise ValueError("<mesh> needs `vertex` or `file`")
                scale = _floats(attrs.get("scale"), 3, [1, 1, 1])
                name = attrs.get("name") or Path(attrs.get("file", f"mesh{len(self.meshes)}")).stem
                self.meshes[name] = convex_vertices(v * scale)

    def _geoms(self, node: ET.Element, bid: int, childclass: Optional[str]) -> None:
        for gn in node.findall("geom"):
            attrs = self.defaults.resolve("geom", gn.get("class", childclass))
            attrs.update(gn.attrib)
            gtype = _GEOM_TYPES[attrs.get("type", "mesh" if "mesh" in attrs else "sphere")]
            size = np.zeros(3)
            s = _floats(attrs.get("size"))
            if s is not None:
                size[: s.size] = s
            pos = _floats(attrs.get("pos"), 3, [0, 0, 0])
            quat = self._orientation(attrs)
            if "fromto" in attrs:
                ft = _floats(attrs["fromto"], 6)
                a, b = ft[:3], ft[3:]
                pos = 0.5 * (a + b)
                quat = _quat_z_to_vec(b - a)
                size[1] = 0.5 * np.linalg.norm(b - a)
            verts = np.zeros((0, 3))
            if gtype == GEOM_BOX:
                verts = box_vertices(size)
            elif gtype == GEOM_MESH:
                if attrs.get("mesh") not in self.meshes:
                    raise ValueError(f"geom {attrs.get('name')!r}: unknown mesh {attrs.get('mesh')!r}")
                verts = self.meshes[attrs["mesh"]]  # kept in the file's coordinates = the geom frame (MuJoCo re-centres meshes on their
                #                                     inertial frame and compensates in geom_pos / geom_quat: same world-frame vertices)
            self.geoms.append(dict(
                name=attrs.get("name", f"geom{len(self.geoms)}"), type=gtype, body=bid, size=size, pos=pos, verts=verts,
                quat=quat, contype=int(attrs.get("contype", 1)), conaffinity=int(attrs.get("conaffinity", 1)),
                condim=int(attrs.get("condim", 3)), priority=int(attrs.get("priority", 0)),
                friction=_floats(attrs.get("friction"), 3, [1.0, 0.005, 0.0001]),
                margin=float(attrs.get("margin", 0.0)), gap=float(attrs.get("gap", 0.0)),
                solmix=float(attrs.get("solmix", 1.0)),
                solref=_floats(attrs.get("solref"), 2, [0.02, 1.0]),
                solimp=_floats(attrs.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
            ))

    # ------------------------------------------------------------------ main entry
    def compile(self) -> MjModel:
        root = self.root
        m = MjModel()

        # ---- <option>
        opt = Option()
        for onode in root.findall("option"):
            if onode.get("timestep"):
                opt.timestep = float(onode.get("timestep"))
            if onode.get("gravity"):
                opt.gravity = _floats(onode.get("gravity"), 3)
            for k in ("impratio", "tolerance", "ls_tolerance"):
                if onode.get(k):
                    setattr(opt, k, float(onode.get(k)))
            for k in ("iterations", "ls_iterations"):
                if onode.get(k):
                    setattr(opt, k, int(onode.get(k)))
            if onode.get("integrator"):
                opt.integrator = _INTEGRATORS[onode.get("integrator").lower()]
            if onode.get("solver"):
                opt.solver = _SOLVERS[onode.get("solver").lower()]
            if onode.get("cone"):
                opt.cone = _CONES[onode.get("cone").lower()]
            if onode.get("jacobian"):
                opt.jacobian = _JACOBIANS[onode.get("jacobian").lower()]
            for fnode in onode.findall("flag"):
                for k, v in fnode.attrib.items():
                    if k in _DISABLE_BITS:
                        if v == "disable":
                            opt.disableflags |= _DISABLE_BITS[k]
                        else:
                            opt.disableflags &= ~_DISABLE_BITS[k]
        m.opt = opt

        # ---- bodies / joints / geoms
        self.bodies.append(dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                                ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]), mass=0.0,
                                inertia=np.zeros(3), has_inertial=False, joints=[]))
        self._load_meshes()
        for wb in root.findall("worldbody"):
            self._geoms(wb, 0, None)
            self._walk(wb, 0, None)

        nbody, njnt, ngeom = len(self.bodies), len(self.joints), len(self.geoms)
        m.nbody, m.njnt, m.ngeom = nbody, njnt, ngeom
        # MuJoCo stores joints grouped by body in body order; our DFS already emits them that way
        # only if bodies are numbered depth-first, which they are.
        m.body_parentid = np.array([b["parent"] for b in self.bodies], dtype=np.int32)
        m.body_pos = np.array([b["pos"] for b in self.bodies])
        m.body_quat = np.array([b["quat"] for b in self.bodies])
        m.body_ipos = np.array([b["ipos"] for b in self.bodies])
        m.body_iquat = np.array([b["iquat"] for b in self.bodies])
        m.body_mass = np.array([b["mass"] for b in self.bodies], dtype=np.float64)
        m.body_inertia = np.array([b["inertia"] for b in self.bodies])
        rootid = np.zeros(nbody, dtype=np.int32)
        weldid = np.zeros(nbody, dtype=np.int32)
        for i in range(1, nbody):
            p = m.body_parentid[i]
            rootid[i] = i if p == 0 else rootid[p]
            weldid[i] = i if self.bodies[i]["joints"] else weldid[p]
        m.body_rootid, m.body_weldid = rootid, weldid

        jnt_qposadr, jnt_dofadr = [], []
        nq = nv = 0
        body_jntnum = np.zeros(nbody, dtype=np.int32)
        body_jntadr = np.full(nbody, -1, dtype=np.int32)
        body_dofnum = np.zeros(nbody, dtype=np.int32)
        body_dofadr = np.full(nbody, -1, dtype=np.int32)
        dof_bodyid, dof_jntid, dof_parentid = [], [], []
        body_lastdof = np.full(nbody, -1, dtype=np.int32)  # last dof of nearest moving ancestor-or-self
        for bid, b in enumerate(self.bodies):
            last = body_lastdof[b["parent"]] if bid > 0 else -1
            if b["joints"]:
                body_jntadr[bid] = b["joints"][0]
                body_jntnum[bid] = len(b["joints"])
                body_dofadr[bid] = nv
            for jid in b["joints"]:
                j = self.joints[jid]
                if j["type"] == JNT_BALL:
                    raise NotImplementedError("ball joints are not supported by the engine")
                if j["type"] == JNT_FREE and (bid == 0 or b["parent"] != 0 or len(b["joints"]) != 1):
                    raise ValueError("free joint must be the only joint of a top-level body")
                jnt_qposadr.append(nq)
                jnt_dofadr.append(nv)
                w = 6 if j["type"] == JNT_FREE else 1
                for _ in range(w):
                    dof_bodyid.append(bid)
                    dof_jntid.append(jid)
                    dof_parentid.append(last)
                    last = nv
                    nv += 1
                nq += 7 if j["type"] == JNT_FREE else 1
            body_dofnum[bid] = nv - body_dofadr[bid] if b["joints"] else 0
            body_lastdof[bid] = last
        m.nq, m.nv = nq, nv
        m.body_jntnum, m.body_jntadr, m.body_dofnum, m.body_dofadr = body_jntnum, body_jntadr, body_dofnum, body_dofadr
        m.body_lastdof = body_lastdof
        m.jnt_type = np.array([j["type"] for j in self.joints], dtype=np.int32)
        m.jnt_qposadr = np.array(jnt_qposadr, dtype=np.int32)
        m.jnt_dofadr = np.array(jnt_dofadr, dtype=np.int32)
        m.jnt_bodyid = np.array([j["body"] for j in self.joints], dtype=np.int32)
        m.jnt_limited = np.array([j["limited"] for j in self.joints], dtype=np.int32)
        m.jnt_solref = np.array([j["solref"] for j in self.joints]).reshape(njnt, 2)
        m.jnt_solimp = np.array([j["solimp"] for j in self.joints]).reshape(njnt, 5)
        m.jnt_pos = np.array([j["pos"] for j in self.joints]).reshape(njnt, 3)
        m.jnt_axis = np.array([j["axis"] for j in self.joints]).reshape(njnt, 3)
        m.jnt_stiffness = np.array([j["stiffness"] for j in self.joints], dtype=np.float64)
        m.jnt_range = np.array([j["range"] for j in self.joints]).reshape(njnt, 2)
        m.jnt_margin = np.array([j["margin"] for j in self.joints], dtype=np.float64)
        m.dof_bodyid = np.array(dof_bodyid, dtype=np.int32)
        m.dof_jntid = np.array(dof_jntid, dtype=np.int32)
        m.dof_parentid = np.array(dof_parentid, dtype=np.int32)
        m.dof_armature = np.array([self.joints[j]["armature"] for j in dof_jntid], dtype=np.float64)
        m.dof_damping = np.array([self.joints[j]["damping"] for j in dof_jntid], dtype=np.float64)
        for jid, j in enumerate(self.joints):
            if j["type"] == JNT_FREE and j["stiffness"] != 0.0:
                raise NotImplementedError("free-joint stiffness is not supported")

        qpos0 = np.zeros(nq)
        qpos_spring = np.zeros(nq)
        for jid, j in enumerate(self.joints):
            a = m.jnt_qposadr[jid]
            if j["type"] == JNT_FREE:
                b = self.bodies[j["body"]]
                qpos0[a:a + 3] = b["pos"]
                qpos0[a + 3:a + 7] = b["quat"]
                qpos_spring[a:a + 7] = qpos0[a:a + 7]
            else:
                qpos0[a] = j["ref"]
                qpos_spring[a] = j["springref"]
        m.qpos0, m.qpos_spring = qpos0, qpos_spring

        m.geom_type = np.array([g["type"] for g in self.geoms], dtype=np.int32)
        m.geom_bodyid = np.array([g["body"] for g in self.geoms], dtype=np.int32)
        m.geom_size = np.array([g["size"] for g in self.geoms]).reshape(ngeom, 3)
        m.geom_pos = np.array([g["pos"] for g in self.geoms]).reshape(ngeom, 3)
        m.geom_quat = np.array([g["quat"] for g in self.geoms]).reshape(ngeom, 4)
        m.geom_contype = np.array([g["contype"] for g in self.geoms], dtype=np.int32)
        m.geom_conaffinity = np.array([g["conaffinity"] for g in self.geoms], dtype=np.int32)
        m.geom_condim = np.array([g["condim"] for g in self.geoms], dtype=np.int32)
        m.geom_friction = np.array([g["friction"] for g in self.geoms]).reshape(ngeom, 3)
        m.geom_priority = np.array([g["priority"] for g in self.geoms], dtype=np.int32)
        m.geom_solmix = np.array([g["solmix"] for g in self.geoms], dtype=np.float64)
        m.geom_solref = np.array([g["solref"] for g in self.geoms]).reshape(ngeom, 2)
        m.geom_solimp = np.array([g["solimp"] for g in self.geoms]).reshape(ngeom, 5)
        m.geom_margin = np.array([g["margin"] for g in self.geoms], dtype=np.float64)
        m.geom_gap = np.array([g["gap"] for g in self.geoms], dtype=np.float64)
        # convex vertex sets of box / mesh geoms, in the geom frame (the engine's model blob carries them)
        adr, num, pool = [], [], []
        for g in self.geoms:
            adr.append(sum(num))
            num.append(len(g["verts"]))
            pool.extend(g["verts"].tolist())
        m.geom_vertadr = np.array(adr, dtype=np.int32)
        m.geom_vertnum = np.array(num, dtype=np.int32)
        m.nvert = int(sum(num))
        m.vert = np.array(pool, dtype=np.float64).reshape(m.nvert, 3)
        attach_convex_topology(m)

        m.names = dict(
            body=[b["name"] for b in self.bodies], joint=[j["name"] for j in self.joints],
            geom=[g["name"] for g in self.geoms],
        )

        self._actuators(m)
        self._equalities(m)
        self._pairs(m)
        self._set_const(m)
        self._keyframes(m)
        return m

    def _joint_id(self, m: MjModel, name: str) -> int:
        try:
            return m.names["joint"].index(name)
        except ValueError:
            raise ValueError(f"unknown joint {name!r}") from None

    def _actuators(self, m: MjModel) -> None:
        acts = []
        for anode in self.root.findall("actuator"):
            for a in anode:
                attrs = self.defaults.resolve(a.tag, a.get("class"))
                attrs.update(a.attrib)
                if "joint" not in attrs:
                    raise NotImplementedError("only joint transmissions are supported")
                jid = self._joint_id(m, attrs["joint"])
                if m.jnt_type[jid] not in (JNT_HINGE, JNT_SLIDE):
                    raise NotImplementedError("actuators on free joints are not supported")
                gainprm = np.zeros(3)
                biasprm = np.zeros(3)
                gaintype, biastype = 0, 0
                if a.tag == "motor":
                    gainprm[0] = 1.0
                elif a.tag == "position":
                    kp = float(attrs.get("kp", 1.0))
                    kv = float(attrs.get("kv", 0.0))
                    gainprm[0] = kp
                    biasprm[:] = [0.0, -kp, -kv]
                    biastype = 1
                elif a.tag == "velocity":
                    kv = float(attrs.get("kv", 1.0))
                    gainprm[0] = kv
                    biasprm[:] = [0.0, 0.0, -kv]
                    biastype = 1
                elif a.tag == "general":
                    gp = _floats(attrs.get("gainprm"))
                    bp = _floats(attrs.get("biasprm"))
                    gainprm[0] = 1.0
                    if gp is not None:
                        gainprm[:] = 0
                        gainprm[: min(3, gp.size)] = gp[:3]
                    if bp is not None:
                        biasprm[: min(3, bp.size)] = bp[:3]
                    gaintype = {"fixed": 0, "affine": 1}[attrs.get("gaintype", "fixed")]
                    biastype = {"none": 0, "affine": 1}[attrs.get("biastype", "none")]
                    if attrs.get("dyntype", "none") != "none":
                        raise NotImplementedError("stateful actuators (dyntype) are not supported")
                else:
                    raise NotImplementedError(f"actuator <{a.tag}> is not supported")
                ctrlrange = _floats(attrs.get("ctrlrange"), 2, [0, 0])
                forcerange = _floats(attrs.get("forcerange"), 2, [0, 0])
                cl = _bool(attrs.get("ctrllimited"))
                fl = _bool(attrs.get("forcelimited"))
                if cl is None:
                    cl = self.autolimits and "ctrlrange" in attrs
                if fl is None:
                    fl = self.autolimits and "forcerange" in attrs
                gear = _floats(attrs.get("gear"))
                acts.append(dict(name=attrs.get("name", f"actuator{len(acts)}"), jid=jid, gaintype=gaintype,
                                 biastype=biastype, gainprm=gainprm, biasprm=biasprm, ctrlrange=ctrlrange,
                                 forcerange=forcerange, ctrllimited=bool(cl), forcelimited=bool(fl),
                                 gear=1.0 if gear is None else float(gear[0])))
        nu = len(acts)
        m.nu, m.na = nu, 0
        m.actuator_trnid = np.array([a["jid"] for a in acts], dtype=np.int32)
        m.actuator_gaintype = np.array([a["gaintype"] for a in acts], dtype=np.int32)
        m.actuator_biastype = np.array([a["biastype"] for a in acts], dtype=np.int32)
        m.actuator_ctrllimited = np.array([a["ctrllimited"] for a in acts], dtype=np.int32)
        m.actuator_forcelimited = np.array([a["forcelimited"] for a in acts], dtype=np.int32)
        m.actuator_ctrlrange = np.array([a["ctrlrange"] for a in acts]).reshape(nu, 2)
        m.actuator_forcerange = np.array([a["forcerange"] for a in acts]).reshape(nu, 2)
        m.actuator_gainprm = np.array([a["gainprm"] for a in acts]).reshape(nu, 3)
        m.actuator_biasprm = np.array([a["biasprm"] for a in acts]).reshape(nu, 3)
        m.actuator_gear = np.array([a["gear"] for a in acts], dtype=np.float64)
        m.names["actuator"] = [a["name"] for a in acts]

    def _equalities(self, m: MjModel) -> None:
        eqs = []
        for enode in self.root.findall("equality"):
            for e in enode:
                attrs = self.defaults.resolve(e.tag, e.get("class"))
                attrs.update(e.attrib)
                if e.tag != "joint":
                    raise NotImplementedError(f"equality <{e.tag}> is not supported (joint only)")
                j1 = self._joint_id(m, attrs["joint1"])
                j2 = self._joint_id(m, attrs["joint2"]) if "joint2" in attrs else -1
                data = np.zeros(11)
                data[:5] = _floats(attrs.get("polycoef"), 5, [0, 1, 0, 0, 0])
                eqs.append(dict(name=attrs.get("name", f"eq{len(eqs)}"), type=2, obj1=j1, obj2=j2,
                                active=_bool(attrs.get("active"), True), data=data,
                                solref=_floats(attrs.get("solref"), 2, [0.02, 1.0]),
                                solimp=_floats(attrs.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0])))
        neq = len(eqs)
        m.neq = neq
        m.eq_type = np.array([e["type"] for e in eqs], dtype=np.int32)
        m.eq_obj1id = np.array([e["obj1"] for e in eqs], dtype=np.int32)
        m.eq_obj2id = np.array([e["obj2"] for e in eqs], dtype=np.int32)
        m.eq_active = np.array([e["active"] for e in eqs], dtype=np.int32)
        m.eq_solref = np.array([e["solref"] for e in eqs]).reshape(neq, 2)
        m.eq_solimp = np.array([e["solimp"] for e in eqs]).reshape(neq, 5)
        m.eq_data = np.array([e["data"] for e in eqs]).reshape(neq, 11)
        m.names["equality"] = [e["name"] for e in eqs]

    def _pairs(self, m: MjModel) -> None:
        excludes = set()
        for cnode in self.root.findall("contact"):
            for ex in cnode.findall("exclude"):
                b1 = m.names["body"].index(ex.get("body1"))
                b2 = m.names["body"].index(ex.get("body2"))
                excludes.add((min(b1, b2), max(b1, b2)))
        m.exclude_signature = np.array(sorted((b1 << 16) + b2 for b1, b2 in excludes), dtype=np.int32)  # mjModel.exclude_signature
        enumerate_pairs(m, self.geoms, excludes)

    # ------------------------------------------------------------------ mj_setConst at qpos0
    def _set_const(self, m: MjModel) -> None:
        nb, nv = m.nbody, m.nv
        # kinematics at qpos0 (every hinge/slide sits at its reference => local transforms only)
        xpos = np.zeros((nb, 3))
        xquat = np.zeros((nb, 4))
        xquat[0] = [1, 0, 0, 0]
        for b in range(1, nb):
            p = m.body_parentid[b]
            xpos[b] = xpos[p] + rotate(m.body_pos[b], xquat[p])
            xquat[b] = quat_mul(xquat[p], m.body_quat[b])
        xipos = np.array([xpos[b] + rotate(m.body_ipos[b], xquat[b]) for b in range(nb)])
        ximat = np.array([quat_to_mat(quat_mul(xquat[b], m.body_iquat[b])) for b in range(nb)])

        subtreemass = m.body_mass.copy()
        for b in range(nb - 1, 0, -1):
            subtreemass[m.body_parentid[b]] += subtreemass[b]
        m.body_subtreemass = subtreemass

        # world-frame spatial Jacobian columns of each dof: (angular axis, point on axis)
        dof_axis = np.zeros((nv, 3))
        dof_point = np.zeros((nv, 3))
        dof_rot = np.zeros(nv, dtype=bool)
        for jid in range(m.njnt):
            b = m.jnt_bodyid[jid]
            d = m.jnt_dofadr[jid]
            R = quat_to_mat(xquat[b])
            if m.jnt_type[jid] == JNT_FREE:
                for k in range(3):
                    dof_axis[d + k] = np.eye(3)[k]
                    dof_axis[d + 3 + k] = R[:, k]
                    dof_point[d + 3 + k] = xpos[b]
                    dof_rot[d + 3 + k] = True
            else:
                dof_axis[d] = R @ m.jnt_axis[jid]
                dof_point[d] = xpos[b] + R @ m.jnt_pos[jid]
                dof_rot[d] = m.jnt_type[jid] == JNT_HINGE

        def body_jac(b: int, point: np.ndarray) -> np.ndarray:
            """6 x nv Jacobian (translational rows first, like mj_jacBodyCom) of `point` on body b."""
            J = np.zeros((6, nv))
            d = m.body_lastdof[b]
            while d >= 0:
                if dof_rot[d]:
                    J[:3, d] = np.cross(dof_axis[d], point - dof_point[d])
                    J[3:, d] = dof_axis[d]
                else:
                    J[:3, d] = dof_axis[d]
                d = m.dof_parentid[d]
            return J

        M = np.zeros((nv, nv))
        for b in range(1, nb):
            if m.body_mass[b] == 0 and not np.any(m.body_inertia[b]):
                continue
            J = body_jac(b, xipos[b])
            Iw = ximat[b] @ np.diag(m.body_inertia[b]) @ ximat[b].T
            M += m.body_mass[b] * J[:3].T @ J[:3] + J[3:].T @ Iw @ J[3:]
        M += np.diag(m.dof_armature)
        m.qM0 = M
        if nv:
            Minv = np.linalg.inv(M)
            inv0 = np.diag(Minv).copy()
            for jid in range(m.njnt):
                if m.jnt_type[jid] == JNT_FREE:
                    d = m.jnt_dofadr[jid]
                    inv0[d:d + 3] = inv0[d:d + 3].mean()
                    inv0[d + 3:d + 6] = inv0[d + 3:d + 6].mean()
            m.dof_invweight0 = inv0
            m.stat.meaninertia = max(MJ_MINVAL, float(np.mean(np.diag(M))))
        else:
            Minv = np.zeros((0, 0))
            m.dof_invweight0 = np.zeros(0)
            m.stat.meaninertia = 1.0
        biw = np.zeros((nb, 2))
        for b in range(1, nb):
            if m.body_weldid[b] == 0:
                continue
            J = body_jac(b, xipos[b])
            A = J @ Minv @ J.T
            biw[b, 0] = max(MJ_MINVAL, np.trace(A[:3, :3]) / 3)
            biw[b, 1] = max(MJ_MINVAL, np.trace(A[3:, 3:]) / 3)
        m.body_invweight0 = biw

    def _keyframes(self, m: MjModel) -> None:
        for knode in self.root.findall("keyframe"):
            for k in knode.findall("key"):
                name = k.get("name", f"key{len(m.keyframes)}")
                m.keyframes[name] = dict(
                    qpos=_floats(k.get("qpos"), m.nq, m.qpos0.tolist()) if k.get("qpos") else m.qpos0.copy(),
                    qvel=_floats(k.get("qvel"), m.nv) if k.get("qvel") else np.zeros(m.nv),
                    ctrl=_floats(k.get("ctrl"), m.nu) if k.get("ctrl") else np.zeros(m.nu),
                )


def compile_urdf(path, force_float: bool = False) -> MjModel:
    """Compile a URDF file the way the reference loads one (MuJoCo's URDF import + ambersim's actuators-from-transmissions and
    equalities-from-mimics, io_utils.py:18-121, 196-204): ambersim_b200.utils.urdf builds the equivalent MJCF tree."""
    from ambersim_b200.utils.urdf import urdf_to_mjcf

    root, base = urdf_to_mjcf(path)
    comp = _Compiler(root, force_float=force_float)
    comp.base_dir = base
    m = comp.compile()
    m.source = str(path)
    return m


def compile_mjcf(path, force_float: bool = False) -> MjModel:
    """Compile an MJCF file into a flat `MjModel` (stand-in for MjModel.from_xml_path)."""
    path = Path(path)
    root = ET.parse(path).getroot()
    if root.tag != "mujoco":
        raise ValueError(f"{path} is not an MJCF file")
    _expand_includes(root, path.parent)
    comp = _Compiler(root, force_float=force_float)
    comp.base_dir = path.parent
    m = comp.compile()
    m.source = str(path)
    return m
