"""MJCF mini-compiler: MJCF XML -> flat model constants (numpy, float64).

The reference builds its model with the MuJoCo C compiler (``mujoco.MjSpec`` /
``dm_control.mjcf`` -> ``mjModel``; /root/reference/envs/fruitfly.py:381-413,
/root/reference/envs/rodent.py:51-84).  MuJoCo is not available in this image, so
this module restates the subset of MuJoCo's model compiler that the rodent and
fruit-fly assets exercise:

  * nested ``<default class>`` with ``childclass`` inheritance,
  * body tree in depth-first order, ``<freejoint>`` / hinge joints,
  * geoms sphere/capsule/ellipsoid/box/cylinder/plane/mesh with
    ``size/fromto/pos/quat/euler/zaxis``, ``density|mass`` -> body inertial frame,
  * ``<contact><exclude>``, contype/conaffinity/parent filters -> static geom pairs
    with mixed contact parameters,
  * ``<tendon><fixed>``, ``<general>/<motor>`` actuators,
  * dm_control ``rescale.rescale_subtree`` (rodent.py:60-64) and MjSpec free-joint
    deletion (fruitfly.py:381-387),
  * ``mj_setConst`` quantities: ``dof_invweight0``, ``body_invweight0``,
    ``stat.meaninertia``.

Field names follow ``mjModel``.  Nothing here runs in the timed path; the compiled
model is packed once into a device blob by ``model.py``.
"""
from __future__ import annotations

import copy
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

mjMINVAL = 1e-15

# geom types (mjtGeom)
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
_GEOM_TYPES = {
    "plane": GEOM_PLANE, "hfield": GEOM_HFIELD, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH,
}
# joint types (mjtJoint)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
# actuator enums
DYN_NONE, DYN_INTEGRATOR, DYN_FILTER = 0, 1, 2
GAIN_FIXED, GAIN_AFFINE = 0, 1
BIAS_NONE, BIAS_AFFINE = 0, 1
TRN_JOINT, TRN_TENDON = 0, 1
CONE_PYRAMIDAL, CONE_ELLIPTIC = 0, 1


# ----------------------------------------------------------------------------------------------
# small math helpers (float64)
# ----------------------------------------------------------------------------------------------
def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.replace(",", " ").split()], dtype=np.float64)
    if n is not None and len(v) < n and default is not None:
        out = np.array(default, dtype=np.float64)
        out[: len(v)] = v
        return out
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(m):
    # robust conversion (Shepperd)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
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


def axisangle_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n < mjMINVAL:
        return np.array([1.0, 0, 0, 0])
    axis = axis / n
    s = np.sin(angle * 0.5)
    return np.array([np.cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s])


def z_to_quat(vec):
    """Minimal rotation taking +z to `vec` (mjuu_z2quat)."""
    vec = np.asarray(vec, dtype=np.float64)
    n = np.linalg.norm(vec)
    if n < mjMINVAL:
        return np.array([1.0, 0, 0, 0])
    vec = vec / n
    z = np.array([0.0, 0, 1.0])
    axis = np.cross(z, vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        if vec[2] > 0:
            return np.array([1.0, 0, 0, 0])
        return np.array([0.0, 1.0, 0, 0])
    ang = np.arctan2(s, vec[2])
    return axisangle_quat(axis / s, ang)


def rotate(v, q):
    return quat_to_mat(q) @ v


# ----------------------------------------------------------------------------------------------
# model container
# ----------------------------------------------------------------------------------------------
@dataclass
class Model:
    """Flat model constants; names follow mjModel.  All float arrays are float64 here."""
    # sizes
    nq: int = 0
    nv: int = 0
    nu: int = 0
    na: int = 0
    nbody: int = 0
    njnt: int = 0
    ngeom: int = 0
    ntendon: int = 0
    nM: int = 0
    # options
    timestep: float = 0.002
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    density: float = 0.0
    viscosity: float = 0.0
    cone: int = CONE_PYRAMIDAL
    impratio: float = 1.0
    tolerance: float = 1e-8
    ls_tolerance: float = 0.01
    iterations: int = 100
    ls_iterations: int = 50
    meaninertia: float = 1.0
    # everything else lives in this dict of numpy arrays / lists
    a: Dict[str, np.ndarray] = field(default_factory=dict)
    names: Dict[str, List[str]] = field(default_factory=dict)

    def __getattr__(self, k):
        a = self.__dict__.get("a")
        if a is not None and k in a:
            return a[k]
        raise AttributeError(k)

    def name2id(self, kind: str, name: str) -> int:
        """mj_name2id: -1 when absent (fruitfly.py:419-428 relies on that)."""
        try:
            return self.names[kind].index(name)
        except ValueError:
            return -1


# ----------------------------------------------------------------------------------------------
# defaults handling
# ----------------------------------------------------------------------------------------------
_ACT_TAGS = ("general", "motor", "position", "velocity", "adhesion", "intvelocity", "damper", "cylinder", "muscle")


class _Defaults:
    def __init__(self, root: ET.Element):
        self.cls: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}
        for d in root.findall("default"):
            self._parse(d, "main", top=True)

    def _parse(self, elem, parent, top=False):
        name = elem.get("class", "main") if not top else elem.get("class", "main")
        if name not in self.cls:
            self.cls[name] = copy.deepcopy(self.cls[parent])
        cur = self.cls[name]
        for c in elem:
            if c.tag == "default":
                continue
            tag = "general" if c.tag in _ACT_TAGS else c.tag
            cur.setdefault(tag, {})
            cur[tag].update(c.attrib)
        for c in elem:
            if c.tag == "default":
                self._parse(c, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cname = elem.get("class", childclass or "main")
        dtag = "general" if tag in _ACT_TAGS else tag
        out = dict(self.cls.get(cname, self.cls["main"]).get(dtag, {}))
        out.update(elem.attrib)
        return out


def _orientation(attrs, eulerseq="xyz", degree=False):
    if "quat" in attrs:
        q = _vec(attrs["quat"])
        return q / np.linalg.norm(q)
    if "euler" in attrs:
        e = _vec(attrs["euler"])
        if degree:
            e = np.deg2rad(e)
        q = np.array([1.0, 0, 0, 0])
        for i, ch in enumerate(eulerseq):
            ax = {"x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1]}[ch.lower()]
            t = axisangle_quat(ax, e[i])
            q = quat_mul(q, t) if ch.islower() else quat_mul(t, q)
        return q / np.linalg.norm(q)
    if "axisangle" in attrs:
        aa = _vec(attrs["axisangle"])
        ang = np.deg2rad(aa[3]) if degree else aa[3]
        return axisangle_quat(aa[:3], ang)
    if "xyaxes" in attrs:
        xy = _vec(attrs["xyaxes"])
        x = xy[:3] / np.linalg.norm(xy[:3])
        y = xy[3:] - x * np.dot(x, xy[3:])
        y = y / np.linalg.norm(y)
        z = np.cross(x, y)
        return mat_to_quat(np.stack([x, y, z], axis=1))
    if "zaxis" in attrs:
        return z_to_quat(_vec(attrs["zaxis"]))
    return np.array([1.0, 0, 0, 0])


# ----------------------------------------------------------------------------------------------
# geom mass properties
# ----------------------------------------------------------------------------------------------
def _geom_volume_inertia(gtype, size):
    """Returns (volume, unit-mass diagonal inertia in the geom frame)."""
    if gtype == GEOM_SPHERE:
        r = size[0]
        return 4.0 / 3.0 * np.pi * r ** 3, np.full(3, 0.4 * r * r)
    if gtype == GEOM_CAPSULE:
        r, hh = size[0], size[1]
        h = 2 * hh
        vol = np.pi * (r * r * h + 4.0 / 3.0 * r ** 3)
        sm = 4 * r / (4 * r + 3 * h) if (4 * r + 3 * h) > 0 else 1.0  # sphere mass fraction
        cm = 1.0 - sm
        ix = cm * (3 * r * r + h * h) / 12.0 + 2 * sm * r * r / 5.0 + sm * h * (3 * r + 2 * h) / 8.0
        iz = cm * r * r / 2.0 + 2 * sm * r * r / 5.0
        return vol, np.array([ix, ix, iz])
    if gtype == GEOM_ELLIPSOID:
        a, b, c = size[:3]
        return 4.0 / 3.0 * np.pi * a * b * c, np.array([b * b + c * c, a * a + c * c, a * a + b * b]) / 5.0
    if gtype == GEOM_CYLINDER:
        r, hh = size[0], size[1]
        h = 2 * hh
        ix = (3 * r * r + h * h) / 12.0
        return np.pi * r * r * h, np.array([ix, ix, r * r / 2.0])
    if gtype == GEOM_BOX:
        a, b, c = size[:3]
        return 8 * a * b * c, np.array([b * b + c * c, a * a + c * c, a * a + b * b]) / 3.0
    return 0.0, np.zeros(3)


def _load_obj(path):
    verts, faces = [], []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append([float(p[1]), float(p[2]), float(p[3])])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int64)


def _mesh_mass_props(verts, faces):
    """Signed-tetrahedron volume, centre of mass and unit-density inertia about the com."""
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    center = verts.mean(axis=0)
    a, b, c = a - center, b - center, c - center
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    vol = det.sum() / 6.0
    if abs(vol) < mjMINVAL:
        return 0.0, center, np.zeros((3, 3))
    if vol < 0:  # inward-facing winding
        det, vol = -det, -vol
    com = (det[:, None] * (a + b + c) / 4.0).sum(axis=0) / (6.0 * vol)
    # second moments: integral of x_i x_j over a tetra with apex at origin = det/120 * (sum ...)
    P = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            s = (
                2 * a[:, i] * a[:, j] + 2 * b[:, i] * b[:, j] + 2 * c[:, i] * c[:, j]
                + a[:, i] * b[:, j] + a[:, j] * b[:, i] + a[:, i] * c[:, j] + a[:, j] * c[:, i]
                + b[:, i] * c[:, j] + b[:, j] * c[:, i]
            )
            P[i, j] = (det * s).sum() / 120.0
    P -= vol * np.outer(com, com)
    inertia = np.trace(P) * np.eye(3) - P
    return vol, com + center, inertia


# ----------------------------------------------------------------------------------------------
# the compiler
# ----------------------------------------------------------------------------------------------
class _Body:
    def __init__(self):
        self.name = ""
        self.parent = -1
        self.pos = np.zeros(3)
        self.quat = np.array([1.0, 0, 0, 0])
        self.joints = []  # list of attr dicts (resolved)
        self.geoms = []
        self.inertial = None
        self.children = []


def _rescale_subtree(elem: ET.Element, pos_factor: float, size_factor: float):
    """dm_control.locomotion.walkers.rescale.rescale_subtree (rodent.py:60-64): only attributes
    *explicitly present* on elements are touched (class defaults are not)."""
    for child in list(elem):
        if child.get("fromto") is not None:
            ft = _vec(child.get("fromto"))
            new_pos = pos_factor * 0.5 * (ft[3:] + ft[:3])
            new_size = size_factor * 0.5 * (ft[3:] - ft[:3])
            child.set("fromto", " ".join(repr(float(x)) for x in np.concatenate([new_pos - new_size, new_pos + new_size])))
        if child.get("pos") is not None:
            child.set("pos", " ".join(repr(float(x)) for x in _vec(child.get("pos")) * pos_factor))
        if child.get("size") is not None:
            child.set("size", " ".join(repr(float(x)) for x in _vec(child.get("size")) * size_factor))
        if child.tag in ("body", "worldbody"):
            _rescale_subtree(child, pos_factor, size_factor)


def _expand_replicate(elem: ET.Element) -> List[Tuple[str, int]]:
    """``<replicate count sep offset euler>``: the children are instantiated ``count`` times with ``sep + k`` appended to every
    name; copy k sits in frame ``T_k = T_{k-1} o (offset, euler)`` (``T_0`` = identity), which is FOLDED into the pose of each
    direct child (MuJoCo frames are not bodies).  ``euler`` is in the compiler's angle unit like every other angle: with
    ``<compiler angle="radian"/>`` the ``euler="0 0 180"`` of assets/rodent_pair.xml:163 is 180 rad (= 233.2 deg), not 180 deg.
    Returns [(sep, count)] of the replicates met."""
    found = []
    comp = {}
    for c in elem.findall("compiler"):
        comp.update(c.attrib)
    degree = comp.get("angle", "degree") == "degree"
    eulerseq = comp.get("eulerseq", "xyz")
    changed = True
    while changed:
        changed = False
        for parent in elem.iter():
            for idx, child in enumerate(list(parent)):
                if child.tag != "replicate":
                    continue
                count = int(child.get("count", "1"))
                sep = child.get("sep", "")
                offset = _vec(child.get("offset")) if child.get("offset") else np.zeros(3)
                dq = _orientation(child.attrib, eulerseq, degree)
                parent.remove(child)
                found.append((sep, count))
                ins = idx
                fpos, fquat = np.zeros(3), np.array([1.0, 0.0, 0.0, 0.0])
                for k in range(count):
                    if k > 0:
                        fpos = fpos + rotate(offset, fquat)
                        fquat = quat_mul(fquat, dq)
                    for sub in child:
                        c = copy.deepcopy(sub)
                        for e in c.iter():
                            if e.get("name") is not None:
                                e.set("name", f"{e.get('name')}{sep}{k}")
                        if c.tag in ("body", "geom", "site", "camera", "light"):
                            if c.get("fromto") is not None:
                                raise NotImplementedError("fromto on a direct child of <replicate>")
                            pos = fpos + rotate(_vec(c.get("pos"), default=[0, 0, 0]), fquat)
                            quat = quat_mul(fquat, _orientation(c.attrib, eulerseq, degree))
                            for k_ in ("euler", "axisangle", "xyaxes", "zaxis"):
                                if k_ in c.attrib:
                                    del c.attrib[k_]
                            c.set("pos", " ".join(repr(float(x)) for x in pos))
                            c.set("quat", " ".join(repr(float(x)) for x in quat))
                        parent.insert(ins, c)
                        ins += 1
                changed = True
                break
            if changed:
                break
    return found


def compile_mjcf(
    path: str,
    scale_factor: Optional[float] = None,
    delete_free_joint_of: Optional[str] = None,
    overrides: Optional[dict] = None,
    missing_mesh: str = "error",
    replicate_actuators: bool = False,
    extra_pairs: Optional[List[Tuple[str, str]]] = None,
) -> Model:
    """Compile an MJCF file.

    scale_factor: apply dm_control ``rescale_subtree(root, s, s)`` first (rodent.py:60-64).
    delete_free_joint_of: body whose first joint is deleted when it is a free joint named 'free'
        (fruitfly.py:381-387, the tethered fly).
    overrides: option overrides applied after compile (solver iterations etc., fruitfly.py:71-78).
    missing_mesh: 'error' | 'skip' -- what to do with mesh files absent from the checkout
        (six fly meshes).  'skip' treats the geom as massless unless it has explicit
        mass, in which case a sphere-equivalent inertia is used (declared deviation).
    replicate_actuators: assets/rodent_pair.xml replicates the animal (``sep="-"``) but its ``<actuator>`` block still
        names the un-suffixed joints, so MuJoCo itself rejects the file.  With this flag an actuator whose joint
        only exists with the replicate suffixes is instantiated once per copy (``name-k`` on ``joint-k``), copy 0 first.
    extra_pairs: explicit contact pairs by geom name, as ``<contact><pair geom1 geom2/>`` would add them (they bypass the
        contype / conaffinity filter; unspecified parameters follow the geoms with the rules of dynamic pairs).
    """
    tree = ET.parse(path)
    root = tree.getroot()
    basedir = os.path.dirname(os.path.abspath(path))
    replicates = _expand_replicate(root)
    if scale_factor is not None:
        _rescale_subtree(root, scale_factor, scale_factor)

    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    degree = comp.get("angle", "degree") == "degree"
    eulerseq = comp.get("eulerseq", "xyz")
    autolimits = comp.get("autolimits", "true") == "true"
    meshdir = comp.get("meshdir", "")
    defaults = _Defaults(root)

    m = Model()
    opt = {}
    for o in root.findall("option"):
        opt.update(o.attrib)
    m.timestep = float(opt.get("timestep", 0.002))
    m.gravity = _vec(opt.get("gravity"), default=[0, 0, -9.81])
    m.density = float(opt.get("density", 0.0))
    m.viscosity = float(opt.get("viscosity", 0.0))
    m.cone = CONE_ELLIPTIC if opt.get("cone", "pyramidal") == "elliptic" else CONE_PYRAMIDAL
    m.impratio = float(opt.get("impratio", 1.0))
    m.tolerance = float(opt.get("tolerance", 1e-8))
    m.ls_tolerance = float(opt.get("ls_tolerance", 0.01))
    m.iterations = int(opt.get("iterations", 100))
    m.ls_iterations = int(opt.get("ls_iterations", 50))

    # ---- meshes
    meshes = {}
    asset = root.find("asset")
    if asset is not None:
        for me in asset.findall("mesh"):
            at = defaults.resolve("mesh", me, None)
            name = at.get("name") or os.path.splitext(os.path.basename(at["file"]))[0]
            meshes[name] = at

    mesh_cache = {}

    def mesh_props(name):
        if name in mesh_cache:
            return mesh_cache[name]
        at = meshes[name]
        fpath = os.path.join(basedir, meshdir, at["file"])
        if not os.path.exists(fpath):
            if missing_mesh == "error":
                raise FileNotFoundError(fpath)
            mesh_cache[name] = None
            return None
        v, f = _load_obj(fpath)
        scale = _vec(at.get("scale"), default=[1, 1, 1])
        v = v * scale
        if np.prod(scale) < 0:
            f = f[:, ::-1]
        mesh_cache[name] = _mesh_mass_props(v, f)
        return mesh_cache[name]

    # ---- walk the body tree depth-first
    bodies: List[_Body] = []
    world = _Body()
    world.name = "world"
    bodies.append(world)

    def parse_geom(at):
        g = {}
        g["name"] = at.get("name", "")
        gtype = _GEOM_TYPES[at.get("type", "sphere")]
        if "mesh" in at and "type" not in at:
            gtype = GEOM_MESH
        g["type"] = gtype
        size = np.zeros(3)
        if "size" in at:
            s = _vec(at["size"])
            size[: min(3, len(s))] = s[:3]
        pos = _vec(at.get("pos"), default=[0, 0, 0])
        quat = _orientation(at, eulerseq, degree)
        if "fromto" in at:
            ft = _vec(at["fromto"])
            pos = 0.5 * (ft[:3] + ft[3:])
            vec = ft[:3] - ft[3:]
            size[1] = 0.5 * np.linalg.norm(vec)
            if gtype in (GEOM_BOX, GEOM_ELLIPSOID):
                size[2] = size[1]
                size[1] = size[0]
            quat = z_to_quat(vec)
        g["size"], g["pos"], g["quat"] = size, pos, quat
        g["contype"] = int(at.get("contype", 1))
        g["conaffinity"] = int(at.get("conaffinity", 1))
        g["condim"] = int(at.get("condim", 3))
        g["priority"] = int(at.get("priority", 0))
        fr = _vec(at.get("friction"), 3, [1, 0.005, 0.0001])
        g["friction"] = fr
        g["solref"] = _vec(at.get("solref"), 2, [0.02, 1])
        g["solimp"] = _vec(at.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2])
        g["solmix"] = float(at.get("solmix", 1.0))
        g["margin"] = float(at.get("margin", 0.0))
        g["gap"] = float(at.get("gap", 0.0))
        g["group"] = int(at.get("group", 0))
        # mass properties (full inertia tensor about geom com, in geom frame)
        mass = None
        if "mass" in at:
            mass = float(at["mass"])
        density = float(at.get("density", 1000.0))
        com_local = np.zeros(3)
        if gtype == GEOM_MESH:
            mp = mesh_props(at["mesh"])
            if mp is None or mp[0] <= 0:
                if mass is not None and mass > 0:
                    # declared deviation: unknown shape -> sphere of the same mass at geom origin
                    r = (3 * mass / (4 * np.pi * density)) ** (1 / 3) if density > 0 else 0.0
                    inertia = np.eye(3) * 0.4 * mass * r * r
                else:
                    mass, inertia = 0.0, np.zeros((3, 3))
            else:
                vol, com_local, iner = mp
                if mass is None:
                    mass = density * vol
                inertia = iner * (mass / vol)
        else:
            vol, unit = _geom_volume_inertia(gtype, size)
            if mass is None:
                mass = density * vol
            inertia = np.diag(unit * mass)
        g["mass"], g["inertia"], g["com_local"] = mass, inertia, com_local
        return g

    def walk(elem, parent_id, childclass):
        for c in elem:
            if c.tag == "body":
                b = _Body()
                b.name = c.get("name", f"body{len(bodies)}")
                b.parent = parent_id
                b.pos = _vec(c.get("pos"), default=[0, 0, 0])
                b.quat = _orientation(c.attrib, eulerseq, degree)
                cc = c.get("childclass", childclass)
                bid = len(bodies)
                bodies.append(b)
                for e in c:
                    if e.tag == "joint":
                        b.joints.append(defaults.resolve("joint", e, cc))
                    elif e.tag == "freejoint":
                        at = dict(e.attrib)
                        at["type"] = "free"
                        b.joints.append(at)
                    elif e.tag == "geom":
                        b.geoms.append(parse_geom(defaults.resolve("geom", e, cc)))
                    elif e.tag == "inertial":
                        b.inertial = dict(e.attrib)
                walk(c, bid, cc)
            elif c.tag == "geom" and parent_id == 0 and elem.tag == "worldbody":
                bodies[0].geoms.append(parse_geom(defaults.resolve("geom", c, childclass)))

    wb = root.find("worldbody")
    walk(wb, 0, wb.get("childclass"))

    if delete_free_joint_of is not None:
        for b in bodies:
            if b.name == delete_free_joint_of and b.joints:
                j0 = b.joints[0]
                if j0.get("type", "hinge") == "free" and j0.get("name") == "free":
                    b.joints.pop(0)

    nbody = len(bodies)
    a = m.a
    m.nbody = nbody
    a["body_parentid"] = np.array([max(b.parent, 0) for b in bodies], dtype=np.int32)
    a["body_pos"] = np.stack([b.pos for b in bodies])
    a["body_quat"] = np.stack([b.quat for b in bodies])
    m.names["body"] = [b.name for b in bodies]

    # ---- joints / dofs
    jnt_type, jnt_qposadr, jnt_dofadr, jnt_bodyid, jnt_pos, jnt_axis = [], [], [], [], [], []
    jnt_range, jnt_limited, jnt_stiffness, jnt_solref, jnt_solimp, jnt_margin = [], [], [], [], [], []
    jnt_names = []
    qpos0, qpos_spring = [], []
    dof_bodyid, dof_jntid, dof_parentid, dof_armature, dof_damping, dof_frictionloss = [], [], [], [], [], []
    body_jntadr = np.full(nbody, -1, dtype=np.int32)
    body_jntnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = np.full(nbody, -1, dtype=np.int32)
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    body_lastdof = np.full(nbody, -1, dtype=np.int32)  # last dof on the chain root->body (incl. own)
    nq = nv = 0
    for bid, b in enumerate(bodies):
        par_last = body_lastdof[b.parent] if bid > 0 else -1
        last = par_last
        if b.joints:
            body_jntadr[bid] = len(jnt_type)
            body_dofadr[bid] = nv
        for j in b.joints:
            jt = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[j.get("type", "hinge")]
            if jt in (JNT_BALL, JNT_SLIDE):
                raise NotImplementedError("ball/slide joints are not used by the rodent/fly assets")
            jid = len(jnt_type)
            jnt_type.append(jt)
            jnt_names.append(j.get("name", f"joint{jid}"))
            jnt_qposadr.append(nq)
            jnt_dofadr.append(nv)
            jnt_bodyid.append(bid)
            jnt_pos.append(_vec(j.get("pos"), default=[0, 0, 0]))
            ax = _vec(j.get("axis"), default=[0, 0, 1])
            n = np.linalg.norm(ax)
            jnt_axis.append(ax / n if n > 0 else ax)
            rng = _vec(j.get("range"), default=[0, 0])
            if degree and jt == JNT_HINGE:
                rng = np.deg2rad(rng)
            jnt_range.append(rng)
            lim = j.get("limited", "auto")
            if lim == "auto":
                limited = autolimits and ("range" in j)
            else:
                limited = lim == "true"
            if jt == JNT_FREE:
                limited = False
            jnt_limited.append(limited)
            stiff = float(j.get("stiffness", 0.0))
            damp = float(j.get("damping", 0.0))
            if "springdamper" in j:
                sd = _vec(j["springdamper"])
                if sd[0] > 0 and sd[1] > 0:
                    raise NotImplementedError("springdamper needs mass; unused by the selected assets")
            jnt_stiffness.append(stiff if jt != JNT_FREE else 0.0)
            jnt_solref.append(_vec(j.get("solreflimit"), 2, [0.02, 1]))
            jnt_solimp.append(_vec(j.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
            jnt_margin.append(float(j.get("margin", 0.0)))
            arm = float(j.get("armature", 0.0))
            fl = float(j.get("frictionloss", 0.0))
            if jt == JNT_FREE:
                q0 = np.concatenate([b.pos, b.quat])
                qpos0.extend(q0)
                qpos_spring.extend(q0)
                for k in range(6):
                    dof_bodyid.append(bid)
                    dof_jntid.append(jid)
                    dof_parentid.append(last)
                    last = nv + k
                    dof_armature.append(arm)
                    dof_damping.append(0.0 if "damping" not in j else damp)
                    dof_frictionloss.append(fl)
                nq += 7
                nv += 6
            else:
                ref = float(j.get("ref", 0.0))
                sref = float(j.get("springref", 0.0))
                if degree:
                    ref, sref = np.deg2rad(ref), np.deg2rad(sref)
                qpos0.append(ref)
                qpos_spring.append(sref)
                dof_bodyid.append(bid)
                dof_jntid.append(jid)
                dof_parentid.append(last)
                last = nv
                dof_armature.append(arm)
                dof_damping.append(damp)
                dof_frictionloss.append(fl)
                nq += 1
                nv += 1
        body_jntnum[bid] = len(b.joints)
        body_dofnum[bid] = nv - body_dofadr[bid] if b.joints else 0
        body_lastdof[bid] = last
    m.nq, m.nv, m.njnt = nq, nv, len(jnt_type)
    a["jnt_type"] = np.array(jnt_type, dtype=np.int32)
    a["jnt_qposadr"] = np.array(jnt_qposadr, dtype=np.int32)
    a["jnt_dofadr"] = np.array(jnt_dofadr, dtype=np.int32)
    a["jnt_bodyid"] = np.array(jnt_bodyid, dtype=np.int32)
    a["jnt_pos"] = np.array(jnt_pos).reshape(-1, 3)
    a["jnt_axis"] = np.array(jnt_axis).reshape(-1, 3)
    a["jnt_range"] = np.array(jnt_range).reshape(-1, 2)
    a["jnt_limited"] = np.array(jnt_limited, dtype=np.int32)
    a["jnt_stiffness"] = np.array(jnt_stiffness)
    a["jnt_solref"] = np.array(jnt_solref).reshape(-1, 2)
    a["jnt_solimp"] = np.array(jnt_solimp).reshape(-1, 5)
    a["jnt_margin"] = np.array(jnt_margin)
    a["qpos0"] = np.array(qpos0)
    a["qpos_spring"] = np.array(qpos_spring)
    a["dof_bodyid"] = np.array(dof_bodyid, dtype=np.int32)
    a["dof_jntid"] = np.array(dof_jntid, dtype=np.int32)
    a["dof_parentid"] = np.array(dof_parentid, dtype=np.int32)
    a["dof_armature"] = np.array(dof_armature)
    a["dof_damping"] = np.array(dof_damping)
    a["dof_frictionloss"] = np.array(dof_frictionloss)
    a["body_jntadr"], a["body_jntnum"] = body_jntadr, body_jntnum
    a["body_dofadr"], a["body_dofnum"] = body_dofadr, body_dofnum
    a["body_lastdof"] = body_lastdof
    m.names["joint"] = jnt_names
    if np.any(a["dof_frictionloss"] > 0):
        raise NotImplementedError("frictionloss is unused by the selected assets")

    # body_rootid / weldid / subtree size (DFS order => subtree is a contiguous id range)
    rootid = np.zeros(nbody, dtype=np.int32)
    weldid = np.zeros(nbody, dtype=np.int32)
    for bid in range(1, nbody):
        p = a["body_parentid"][bid]
        rootid[bid] = bid if p == 0 else rootid[p]
        weldid[bid] = bid if body_jntnum[bid] > 0 else weldid[p]
    a["body_rootid"], a["body_weldid"] = rootid, weldid
    sub = np.ones(nbody, dtype=np.int32)
    for bid in range(nbody - 1, 0, -1):
        sub[a["body_parentid"][bid]] += sub[bid]
    a["body_subtreenum"] = sub
    depth = np.zeros(nbody, dtype=np.int32)
    for bid in range(1, nbody):
        depth[bid] = depth[a["body_parentid"][bid]] + 1
    a["body_depth"] = depth

    # sparse M layout (mj: row i = [M(i,i), M(i,parent(i)), ...])
    dof_Madr = np.zeros(nv, dtype=np.int32)
    dof_depth = np.zeros(nv, dtype=np.int32)
    nM = 0
    for i in range(nv):
        dof_Madr[i] = nM
        j, cnt = i, 0
        while j >= 0:
            cnt += 1
            j = a["dof_parentid"][j]
        dof_depth[i] = cnt - 1
        nM += cnt
    m.nM = nM
    a["dof_Madr"], a["dof_depth"] = dof_Madr, dof_depth
    dsub = np.ones(nv, dtype=np.int32)
    for i in range(nv - 1, -1, -1):
        p = a["dof_parentid"][i]
        if p >= 0:
            dsub[p] += dsub[i]
    a["dof_subtreenum"] = dsub

    # ---- body inertial frames
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    for bid, b in enumerate(bodies):
        if bid == 0:
            continue
        if b.inertial is not None:
            it = b.inertial
            body_mass[bid] = float(it["mass"])
            body_ipos[bid] = _vec(it.get("pos"), default=[0, 0, 0])
            q = _orientation(it, eulerseq, degree)
            if "fullinertia" in it:
                fi = _vec(it["fullinertia"])
                I = np.array([[fi[0], fi[3], fi[4]], [fi[3], fi[1], fi[5]], [fi[4], fi[5], fi[2]]])
                w, V = np.linalg.eigh(I)
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                body_inertia[bid] = w
                q = quat_mul(q, mat_to_quat(V))
            else:
                body_inertia[bid] = _vec(it["diaginertia"])
            body_iquat[bid] = q
            continue
        mass = 0.0
        com = np.zeros(3)
        parts = []
        for g in b.geoms:
            if g["mass"] <= 0:
                continue
            R = quat_to_mat(g["quat"])
            c = g["pos"] + R @ g["com_local"]
            parts.append((g["mass"], c, R @ g["inertia"] @ R.T))
            mass += g["mass"]
            com += g["mass"] * c
        if mass <= 0:
            continue
        com /= mass
        I = np.zeros((3, 3))
        for (gm, c, gi) in parts:
            d = c - com
            I += gi + gm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        w, V = np.linalg.eigh(I)
        # MuJoCo sorts principal moments in decreasing order; keep a right-handed frame
        order = np.argsort(-w)
        w, V = w[order], V[:, order]
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        body_mass[bid], body_ipos[bid], body_inertia[bid], body_iquat[bid] = mass, com, w, mat_to_quat(V)
    a["body_mass"], a["body_ipos"], a["body_iquat"], a["body_inertia"] = body_mass, body_ipos, body_iquat, body_inertia
    for bid in range(1, nbody):
        if body_jntnum[bid] > 0 and body_mass[a["body_weldid"] == bid].sum() <= 0:
            # a moving body needs mass somewhere in its welded group + descendants; checked loosely
            pass

    # ---- geoms (flat)
    glist = []
    for bid, b in enumerate(bodies):
        for g in b.geoms:
            g = dict(g)
            g["bodyid"] = bid
            glist.append(g)
    m.ngeom = len(glist)
    a["geom_type"] = np.array([g["type"] for g in glist], dtype=np.int32)
    a["geom_bodyid"] = np.array([g["bodyid"] for g in glist], dtype=np.int32)
    a["geom_pos"] = np.array([g["pos"] for g in glist]).reshape(-1, 3)
    a["geom_quat"] = np.array([g["quat"] for g in glist]).reshape(-1, 4)
    a["geom_size"] = np.array([g["size"] for g in glist]).reshape(-1, 3)
    a["geom_contype"] = np.array([g["contype"] for g in glist], dtype=np.int32)
    a["geom_conaffinity"] = np.array([g["conaffinity"] for g in glist], dtype=np.int32)
    a["geom_condim"] = np.array([g["condim"] for g in glist], dtype=np.int32)
    a["geom_priority"] = np.array([g["priority"] for g in glist], dtype=np.int32)
    a["geom_friction"] = np.array([g["friction"] for g in glist]).reshape(-1, 3)
    a["geom_solref"] = np.array([g["solref"] for g in glist]).reshape(-1, 2)
    a["geom_solimp"] = np.array([g["solimp"] for g in glist]).reshape(-1, 5)
    a["geom_solmix"] = np.array([g["solmix"] for g in glist])
    a["geom_margin"] = np.array([g["margin"] for g in glist])
    a["geom_gap"] = np.array([g["gap"] for g in glist])
    m.names["geom"] = [g["name"] for g in glist]

    # ---- contact pairs (static list; MJX collision_driver semantics)
    excl = set()
    con = root.find("contact")
    if con is not None:
        for e in con.findall("exclude"):
            b1, b2 = m.name2id("body", e.get("body1")), m.name2id("body", e.get("body2"))
            if b1 >= 0 and b2 >= 0:
                excl.add((min(b1, b2), max(b1, b2)))
        if con.findall("pair"):
            raise NotImplementedError("explicit <pair> is unused by the selected assets")
    pairs = []
    for g1 in range(m.ngeom):
        for g2 in range(g1 + 1, m.ngeom):
            b1, b2 = a["geom_bodyid"][g1], a["geom_bodyid"][g2]
            if not ((a["geom_contype"][g1] & a["geom_conaffinity"][g2]) or (a["geom_contype"][g2] & a["geom_conaffinity"][g1])):
                continue
            w1, w2 = a["body_weldid"][b1], a["body_weldid"][b2]
            if w1 == w2:
                continue
            if (min(b1, b2), max(b1, b2)) in excl:
                continue
            # parent-child filter (on welded groups), unless the parent is the static world group
            pw1 = a["body_weldid"][a["body_parentid"][w1]]
            pw2 = a["body_weldid"][a["body_parentid"][w2]]
            if (w1 != 0 and w2 != 0) and (pw1 == w2 or pw2 == w1):
                continue
            i, j = g1, g2
            if a["geom_type"][i] > a["geom_type"][j]:
                i, j = j, i
            pairs.append((i, j))
    for n1, n2 in (extra_pairs or []):
        i, j = m.name2id("geom", n1), m.name2id("geom", n2)
        if i < 0 or j < 0:
            raise ValueError(f"extra pair ({n1}, {n2}): geom not found")
        if a["geom_type"][i] > a["geom_type"][j]:
            i, j = j, i
        if (i, j) in pairs or (j, i) in pairs:
            raise ValueError(f"extra pair ({n1}, {n2}) is already a dynamic pair")
        pairs.append((i, j))
    P = len(pairs)
    pair_geom = np.array(pairs, dtype=np.int32).reshape(-1, 2)
    pair_condim = np.zeros(P, dtype=np.int32)
    pair_friction = np.zeros((P, 5))
    pair_solref = np.zeros((P, 2))
    pair_solimp = np.zeros((P, 5))
    pair_margin = np.zeros(P)
    pair_gap = np.zeros(P)
    pair_ncon = np.zeros(P, dtype=np.int32)
    for k, (i, j) in enumerate(pairs):
        p1, p2 = a["geom_priority"][i], a["geom_priority"][j]
        f1, f2 = a["geom_friction"][i], a["geom_friction"][j]
        if p1 == p2:
            fr = np.maximum(f1, f2)
            pair_condim[k] = max(a["geom_condim"][i], a["geom_condim"][j])
            s1, s2 = a["geom_solmix"][i], a["geom_solmix"][j]
            mix = s1 / (s1 + s2) if (s1 + s2) > mjMINVAL else 0.5
            r1, r2 = a["geom_solref"][i], a["geom_solref"][j]
            if r1[0] > 0 and r2[0] > 0:
                pair_solref[k] = mix * r1 + (1 - mix) * r2
            else:
                pair_solref[k] = np.minimum(r1, r2)
            pair_solimp[k] = mix * a["geom_solimp"][i] + (1 - mix) * a["geom_solimp"][j]
        else:
            hi = i if p1 > p2 else j
            fr = a["geom_friction"][hi]
            pair_condim[k] = a["geom_condim"][hi]
            pair_solref[k] = a["geom_solref"][hi]
            pair_solimp[k] = a["geom_solimp"][hi]
        pair_friction[k] = [fr[0], fr[0], fr[1], fr[2], fr[2]]
        pair_margin[k] = max(a["geom_margin"][i], a["geom_margin"][j])
        pair_gap[k] = max(a["geom_gap"][i], a["geom_gap"][j])
        t1, t2 = a["geom_type"][i], a["geom_type"][j]
        if (t1, t2) == (GEOM_PLANE, GEOM_CAPSULE):
            pair_ncon[k] = 2
        elif (t1, t2) in ((GEOM_PLANE, GEOM_ELLIPSOID), (GEOM_PLANE, GEOM_SPHERE), (GEOM_CAPSULE, GEOM_CAPSULE),
                          (GEOM_SPHERE, GEOM_SPHERE), (GEOM_SPHERE, GEOM_CAPSULE)):
            pair_ncon[k] = 1
        else:
            raise NotImplementedError(f"collision pair types {(t1, t2)} ({m.names['geom'][i]}, {m.names['geom'][j]})")
    a["pair_geom"], a["pair_condim"], a["pair_friction"] = pair_geom, pair_condim, pair_friction
    a["pair_solref"], a["pair_solimp"], a["pair_margin"], a["pair_gap"] = pair_solref, pair_solimp, pair_margin, pair_gap
    a["pair_ncon"] = pair_ncon

    # ---- tendons (fixed only)
    ten_adr, ten_num, wrap_jnt, wrap_coef, ten_names = [], [], [], [], []
    tnd = root.find("tendon")
    if tnd is not None:
        for t in tnd:
            if t.tag != "fixed":
                raise NotImplementedError("only <fixed> tendons are used by the selected assets")
            at = defaults.resolve("tendon", t, None)
            if at.get("limited", "false") == "true":
                raise NotImplementedError("tendon limits are unused by the selected assets")
            ten_names.append(at.get("name", ""))
            ten_adr.append(len(wrap_jnt))
            n = 0
            for jn in t.findall("joint"):
                jid = m.name2id("joint", jn.get("joint"))
                if jid < 0:
                    raise ValueError(f"tendon joint {jn.get('joint')} not found")
                wrap_jnt.append(jid)
                wrap_coef.append(float(jn.get("coef")))
                n += 1
            ten_num.append(n)
    m.ntendon = len(ten_adr)
    a["tendon_adr"] = np.array(ten_adr, dtype=np.int32)
    a["tendon_num"] = np.array(ten_num, dtype=np.int32)
    a["wrap_jntid"] = np.array(wrap_jnt, dtype=np.int32)
    a["wrap_coef"] = np.array(wrap_coef)
    m.names["tendon"] = ten_names

    # ---- actuators
    acts = []
    act_root = root.find("actuator")
    act_elems = []   # (element, name suffix)
    if act_root is not None:
        plain = list(act_root)
        if replicate_actuators and replicates:
            sep, count = replicates[0]
            for k in range(count):
                for e in plain:
                    jn = e.get("joint")
                    if jn is not None and m.name2id("joint", jn) < 0 and m.name2id("joint", f"{jn}{sep}{k}") >= 0:
                        act_elems.append((e, f"{sep}{k}"))
                    elif k == 0:
                        act_elems.append((e, ""))
        else:
            act_elems = [(e, "") for e in plain]
    for e, sfx in act_elems:
        at = defaults.resolve(e.tag, e, None)
        u = {"name": at.get("name", "") + sfx}
        if "joint" in at:
            jname = at["joint"] + sfx
            jid = m.name2id("joint", jname)
            if jid < 0:
                raise ValueError(f"actuator joint {jname} not found")
            u["trntype"], u["trnid"] = TRN_JOINT, jid
        elif "tendon" in at:
            tid = m.name2id("tendon", at["tendon"])
            if tid < 0:
                raise ValueError(f"actuator tendon {at['tendon']} not found")
            u["trntype"], u["trnid"] = TRN_TENDON, tid
        else:
            raise NotImplementedError("only joint/tendon transmissions are used")
        u["gear"] = _vec(at.get("gear"), 6, [1, 0, 0, 0, 0, 0])[0]
        gainprm = _vec(at.get("gainprm"), 3, [1, 0, 0])
        biasprm = _vec(at.get("biasprm"), 3, [0, 0, 0])
        dynprm = _vec(at.get("dynprm"), 3, [1, 0, 0])
        dyntype = {"none": DYN_NONE, "integrator": DYN_INTEGRATOR, "filter": DYN_FILTER}[at.get("dyntype", "none")]
        gaintype = {"fixed": GAIN_FIXED, "affine": GAIN_AFFINE}[at.get("gaintype", "fixed")]
        biastype = {"none": BIAS_NONE, "affine": BIAS_AFFINE}[at.get("biastype", "none")]
        if e.tag == "motor":
            dyntype, gaintype, biastype = DYN_NONE, GAIN_FIXED, BIAS_NONE
            gainprm = np.array([1.0, 0, 0])
            biasprm = np.zeros(3)
        elif e.tag != "general":
            raise NotImplementedError(f"actuator shortcut <{e.tag}> is unused by the selected assets")
        if dyntype == DYN_INTEGRATOR:
            raise NotImplementedError("integrator dyntype unused")
        cl = at.get("ctrllimited", "auto")
        ctrllimited = (autolimits and "ctrlrange" in at) if cl == "auto" else (cl == "true")
        fl = at.get("forcelimited", "auto")
        forcelimited = (autolimits and "forcerange" in at) if fl == "auto" else (fl == "true")
        u.update(dyntype=dyntype, gaintype=gaintype, biastype=biastype, gainprm=gainprm, biasprm=biasprm,
                 dynprm=dynprm, ctrllimited=ctrllimited, forcelimited=forcelimited,
                 ctrlrange=_vec(at.get("ctrlrange"), 2, [0, 0]), forcerange=_vec(at.get("forcerange"), 2, [0, 0]))
        acts.append(u)
    m.nu = len(acts)
    m.na = sum(1 for u in acts if u["dyntype"] != DYN_NONE)
    a["actuator_trntype"] = np.array([u["trntype"] for u in acts], dtype=np.int32)
    a["actuator_trnid"] = np.array([u["trnid"] for u in acts], dtype=np.int32)
    a["actuator_gear"] = np.array([u["gear"] for u in acts])
    a["actuator_dyntype"] = np.array([u["dyntype"] for u in acts], dtype=np.int32)
    a["actuator_gaintype"] = np.array([u["gaintype"] for u in acts], dtype=np.int32)
    a["actuator_biastype"] = np.array([u["biastype"] for u in acts], dtype=np.int32)
    a["actuator_gainprm"] = np.array([u["gainprm"] for u in acts]).reshape(-1, 3)
    a["actuator_biasprm"] = np.array([u["biasprm"] for u in acts]).reshape(-1, 3)
    a["actuator_dynprm"] = np.array([u["dynprm"] for u in acts]).reshape(-1, 3)
    a["actuator_ctrllimited"] = np.array([u["ctrllimited"] for u in acts], dtype=np.int32)
    a["actuator_forcelimited"] = np.array([u["forcelimited"] for u in acts], dtype=np.int32)
    a["actuator_ctrlrange"] = np.array([u["ctrlrange"] for u in acts]).reshape(-1, 2)
    a["actuator_forcerange"] = np.array([u["forcerange"] for u in acts]).reshape(-1, 2)
    # activation address: stateful actuators must come last in MuJoCo; here all or none are stateful
    actadr = np.full(m.nu, -1, dtype=np.int32)
    k = 0
    for i, u in enumerate(acts):
        if u["dyntype"] != DYN_NONE:
            actadr[i] = k
            k += 1
    a["actuator_actadr"] = actadr
    m.names["actuator"] = [u["name"] for u in acts]

    if overrides:
        for k, v in overrides.items():
            setattr(m, k, v)

    _set_const(m)
    return m


# ----------------------------------------------------------------------------------------------
# mj_setConst restatement (float64, single configuration qpos0)
# ----------------------------------------------------------------------------------------------
def kinematics_np(m: Model, qpos):
    """Forward kinematics at one configuration.  Returns dict of world-frame arrays."""
    a = m.a
    nb = m.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
    xanchor = np.zeros((m.njnt, 3))
    xaxis = np.zeros((m.njnt, 3))
    for b in range(1, nb):
        p = a["body_parentid"][b]
        pos = xpos[p] + rotate(a["body_pos"][b], xquat[p])
        quat = quat_mul(xquat[p], a["body_quat"][b])
        for k in range(a["body_jntnum"][b]):
            j = a["body_jntadr"][b] + k
            qa = a["jnt_qposadr"][j]
            if a["jnt_type"][j] == JNT_FREE:
                xanchor[j] = qpos[qa:qa + 3]
                xaxis[j] = [0, 0, 1]
                pos = np.array(qpos[qa:qa + 3], dtype=np.float64)
                quat = np.array(qpos[qa + 3:qa + 7], dtype=np.float64)
                quat = quat / np.linalg.norm(quat)
            else:
                anchor = rotate(a["jnt_pos"][j], quat) + pos
                axis = rotate(a["jnt_axis"][j], quat)
                xanchor[j], xaxis[j] = anchor, axis
                quat = quat_mul(quat, axisangle_quat(a["jnt_axis"][j], qpos[qa] - a["qpos0"][qa]))
                pos = anchor - rotate(a["jnt_pos"][j], quat)
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    xipos = np.array([xpos[b] + rotate(a["body_ipos"][b], xquat[b]) for b in range(nb)])
    ximat = np.array([quat_to_mat(quat_mul(xquat[b], a["body_iquat"][b])) for b in range(nb)])
    return dict(xpos=xpos, xquat=xquat, xanchor=xanchor, xaxis=xaxis, xipos=xipos, ximat=ximat)


def dense_mass_matrix_np(m: Model, qpos):
    """Dense joint-space inertia via body Jacobians (independent of the CRB code paths)."""
    a = m.a
    k = kinematics_np(m, qpos)
    nv = m.nv
    M = np.zeros((nv, nv))
    jacs = body_jacobians_np(m, k)
    for b in range(1, m.nbody):
        if a["body_mass"][b] <= 0:
            continue
        Jp, Jr = jacs[b]
        R = k["ximat"][b]
        I = R @ np.diag(a["body_inertia"][b]) @ R.T
        M += a["body_mass"][b] * Jp.T @ Jp + Jr.T @ I @ Jr
    M += np.diag(a["dof_armature"])
    return M, k, jacs


def body_jacobians_np(m: Model, k):
    """Per body: (Jp at xipos [3,nv], Jr [3,nv])."""
    a = m.a
    out = [None] * m.nbody
    for b in range(m.nbody):
        Jp = np.zeros((3, m.nv))
        Jr = np.zeros((3, m.nv))
        d = a["body_lastdof"][b]
        point = k["xipos"][b]
        while d >= 0:
            j = a["dof_jntid"][d]
            if a["jnt_type"][j] == JNT_FREE:
                kk = d - a["jnt_dofadr"][j]
                if kk < 3:
                    Jp[kk, d] = 1.0
                else:
                    R = quat_to_mat(k["xquat"][a["jnt_bodyid"][j]])
                    ax = R[:, kk - 3]
                    Jr[:, d] = ax
                    Jp[:, d] = np.cross(ax, point - k["xanchor"][j])
            else:
                ax = k["xaxis"][j]
                Jr[:, d] = ax
                Jp[:, d] = np.cross(ax, point - k["xanchor"][j])
            d = a["dof_parentid"][d]
        out[b] = (Jp, Jr)
    return out


def _set_const(m: Model):
    a = m.a
    nv = m.nv
    if nv == 0:
        a["dof_invweight0"] = np.zeros(0)
        a["body_invweight0"] = np.zeros((m.nbody, 2))
        m.meaninertia = 1.0
        return
    M, k, jacs = dense_mass_matrix_np(m, a["qpos0"])
    Minv = np.linalg.inv(M)
    m.meaninertia = float(np.mean(np.diag(M)))
    inv = np.diag(Minv).copy()
    for j in range(m.njnt):
        if a["jnt_type"][j] == JNT_FREE:
            d = a["jnt_dofadr"][j]
            inv[d:d + 3] = inv[d:d + 3].mean()
            inv[d + 3:d + 6] = inv[d + 3:d + 6].mean()
    a["dof_invweight0"] = inv
    biw = np.zeros((m.nbody, 2))
    for b in range(1, m.nbody):
        if a["body_weldid"][b] == 0:
            continue
        Jp, Jr = jacs[b]
        biw[b, 0] = np.trace(Jp @ Minv @ Jp.T) / 3.0
        biw[b, 1] = np.trace(Jr @ Minv @ Jr.T) / 3.0
    a["body_invweight0"] = biw
    a["dof_M0"] = np.diag(M).copy()
