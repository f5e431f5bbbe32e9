"""MJCF-subset model compiler (host side, NumPy).

The reference delegates model loading to ``mj.MjModel.from_xml_path`` / ``from_xml_string``
(reference ``mujoco_template/model.py:22-31``); MuJoCo is not available offline, so this
module restates the part of MuJoCo's XML compiler the four example models
(``examples/{pendulum,cartpole,drone,humanoid}``) and the reference test fixture
(``tests/test_mujoco_template.py:40-61``) exercise:

* ``<compiler angle autolimits eulerseq>``, ``<option>``, ``<include>``, the ``<default>``
  class tree with ``childclass``, bodies / joints / freejoints / geoms (``fromto``) / sites,
  inertia-from-geoms, motors (joint and site transmission) and position/velocity/general
  affine actuators, fixed tendons, ``<contact><exclude>``, keyframes;
* ``mj_setConst`` quantities: ``dof_invweight0``, ``body_invweight0``,
  ``tendon_invweight0``, ``stat.meaninertia``;
* the static part of collision filtering (``mj_collision`` body/geom filters) and
  ``mj_contactParam`` mixing, folded into a pre-mixed candidate pair list.

Anything outside the subset raises ``ConfigError`` (no CPU fallback, no silent ignore of
physics-relevant features).  The output is a dict of NumPy arrays (``compile_*``) that
``_layout.pack`` serialises for the C-ABI.
"""

from __future__ import annotations

import math
import os
import warnings
import xml.etree.ElementTree as ET

import numpy as np

from .exceptions import ConfigError

mjMINVAL = 1e-15

# enums (values match MuJoCo's mjtGeom / mjtJoint / mjtTrn / mjtIntegrator)
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
GEOM_TYPES = {
    "plane": GEOM_PLANE, "hfield": GEOM_HFIELD, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH,
}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
JNT_TYPES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
# sensor tag -> (type code shared with the kernels and the oracle, output width)
SENS_JOINTPOS, SENS_JOINTVEL, SENS_FRAMEPOS, SENS_FRAMEQUAT, SENS_GYRO, SENS_VELOCIMETER, SENS_ACCELEROMETER = range(7)
SENSOR_TYPES = {
    "jointpos": (SENS_JOINTPOS, 1), "jointvel": (SENS_JOINTVEL, 1), "framepos": (SENS_FRAMEPOS, 3),
    "framequat": (SENS_FRAMEQUAT, 4), "gyro": (SENS_GYRO, 3), "velocimeter": (SENS_VELOCIMETER, 3),
    "accelerometer": (SENS_ACCELEROMETER, 3),
}
TRN_JOINT, TRN_SITE = 0, 4
INT_EULER, INT_RK4 = 0, 1

# narrow-phase routines available in the CUDA library / oracle, keyed by (type1 <= type2)
SUPPORTED_PAIRS = {
    (GEOM_PLANE, GEOM_SPHERE), (GEOM_PLANE, GEOM_CAPSULE), (GEOM_PLANE, GEOM_BOX),
    (GEOM_PLANE, GEOM_ELLIPSOID), (GEOM_SPHERE, GEOM_SPHERE), (GEOM_SPHERE, GEOM_CAPSULE),
    (GEOM_CAPSULE, GEOM_CAPSULE),
}


# ----------------------------------------------------------------------------- math
def _floats(text: str | None, n: int | None = None, default=None) -> np.ndarray | None:
    if text is None:
        return None if default is None else np.array(default, dtype=float)
    vals = np.array([float(t) for t in text.replace(",", " ").split()], dtype=float)
    if n is not None and vals.size != n:
        raise ConfigError(f"expected {n} numbers, got {vals.size}: {text!r}")
    return vals


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


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w>=0 branch of the standard trace method)."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def axisangle_to_quat(axis, angle):
    axis = np.asarray(axis, dtype=float)
    n = np.linalg.norm(axis)
    if n < mjMINVAL:
        return np.array([1.0, 0, 0, 0])
    axis = axis / n
    return np.concatenate([[math.cos(angle / 2)], axis * math.sin(angle / 2)])


def z_to_quat(vec):
    """Quaternion rotating (0,0,1) onto ``vec`` (MuJoCo mjuu_z2quat)."""
    vec = np.asarray(vec, dtype=float)
    vec = vec / np.linalg.norm(vec)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(z, vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / s
    ang = math.atan2(s, vec[2])
    return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


# ----------------------------------------------------------------------------- defaults
_ACTUATOR_TAGS = ("general", "motor", "position", "velocity")
_DEFAULT_TAGS = ("geom", "joint", "site", "tendon", "mesh", "camera", "light", "material", "pair", "equality") + _ACTUATOR_TAGS


class _Defaults:
    def __init__(self) -> None:
        self.classes: dict[str, dict[str, dict[str, str]]] = {"main": {}}
        self.parent: dict[str, str | None] = {"main": None}

    def add(self, node: ET.Element, parent: str | None) -> None:
        """Register one <default> node; nested classes start as a copy of their parent."""
        name = node.get("class")
        if parent is None:
            name = name or "main"
            if name != "main":
                raise ConfigError("top-level <default> must be the 'main' class")
        else:
            if not name:
                raise ConfigError("nested <default> requires a class attribute")
            if name in self.classes:
                raise ConfigError(f"repeated default class name: {name}")
            self.classes[name] = {k: dict(v) for k, v in self.classes[parent].items()}
            self.parent[name] = parent
        cur = self.classes[name]
        for child in node:
            if child.tag == "default":
                continue
            tag = "actuator" if child.tag in _ACTUATOR_TAGS else child.tag
            cur.setdefault(tag, {}).update(child.attrib)
        for child in node:
            if child.tag == "default":
                self.add(child, name)

    def get(self, cls: str | None, tag: str) -> dict[str, str]:
        cls = cls or "main"
        if cls not in self.classes:
            raise ConfigError(f"unknown default class: {cls}")
        return self.classes[cls].get(tag, {})


# ----------------------------------------------------------------------------- compiler
class _Compiler:
    def __init__(self) -> None:
        self.angle_deg = True
        self.autolimits = True
        self.eulerseq = "xyz"
        self.defaults = _Defaults()
        self.opt = dict(
            timestep=0.002, gravity=np.array([0, 0, -9.81]), wind=np.zeros(3), density=0.0,
            viscosity=0.0, integrator=INT_EULER, iterations=100, ls_iterations=50,
            tolerance=1e-8, ls_tolerance=0.01, impratio=1.0, disableflags=0,
        )
        self.bodies: list[dict] = []
        self.joints: list[dict] = []
        self.geoms: list[dict] = []
        self.sites: list[dict] = []
        self.actuators: list[dict] = []
        self.tendons: list[dict] = []
        self.excludes: list[tuple[str, str]] = []
        self.keys: list[dict] = []
        self.sensors: list[dict] = []
        self.model_name = "model"

    # ---- xml loading with <include>
    def load(self, root: ET.Element, basedir: str) -> None:
        self._expand_includes(root, basedir)
        if root.tag != "mujoco":
            raise ConfigError(f"root element must be <mujoco>, got <{root.tag}>")
        self.model_name = root.get("model", "model")
        for node in root.findall("compiler"):
            self._compiler(node)
        for node in root.findall("option"):
            self._option(node)
        for node in root.findall("default"):
            self.defaults.add(node, None)
        world = dict(name="world", parent=-1, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                     childclass=None, joints=[], geoms=[], explicit_inertial=None)
        self.bodies.append(world)
        for node in root.findall("worldbody"):
            self._body_children(node, 0, None)
        for node in root.findall("actuator"):
            for a in node:
                self._actuator(a)
        for node in root.findall("tendon"):
            for t in node:
                self._tendon(t)
        for node in root.findall("contact"):
            for c in node:
                if c.tag == "exclude":
                    self.excludes.append((c.get("body1"), c.get("body2")))
                else:
                    raise ConfigError(f"<contact><{c.tag}> is not supported by the B200 path")
        for node in root.findall("keyframe"):
            for k in node.findall("key"):
                self.keys.append(dict(k.attrib))
        for node in root.findall("sensor"):
            for s in node:
                self.sensors.append(dict(s.attrib, tag=s.tag))
        for node in root.findall("equality"):
            if len(node):
                raise ConfigError("<equality> constraints are not supported by the B200 path")

    def _expand_includes(self, node: ET.Element, basedir: str) -> None:
        i = 0
        while i < len(node):
            child = node[i]
            if child.tag == "include":
                path = os.path.join(basedir, child.get("file"))
                try:
                    inc = ET.parse(path).getroot()
                except (OSError, ET.ParseError) as exc:
                    raise ConfigError(f"cannot include {path}: {exc}") from exc
                self._expand_includes(inc, os.path.dirname(path))
                node.remove(child)
                if inc.tag == "mujoco" and node.tag == "mujoco":
                    if inc.get("model") and not node.get("model"):
                        node.set("model", inc.get("model"))
                items = list(inc) if inc.tag in ("mujoco", "mujocoinclude") else [inc]
                for k, item in enumerate(items):
                    node.insert(i + k, item)
                i += len(items)
            else:
                self._expand_includes(child, basedir)
                i += 1

    def _compiler(self, node: ET.Element) -> None:
        if "angle" in node.attrib:
            self.angle_deg = node.get("angle") == "degree"
        if "autolimits" in node.attrib:
            self.autolimits = node.get("autolimits") == "true"
        if "eulerseq" in node.attrib:
            self.eulerseq = node.get("eulerseq")
        if node.get("coordinate", "local") != "local":
            raise ConfigError("compiler coordinate='global' is not supported")
        if node.get("inertiafromgeom", "auto") == "false":
            raise ConfigError("inertiafromgeom='false' is not supported")

    def _option(self, node: ET.Element) -> None:
        o = self.opt
        for key in ("timestep", "density", "viscosity", "tolerance", "ls_tolerance", "impratio"):
            if key in node.attrib:
                o[key] = float(node.get(key))
        for key in ("gravity", "wind"):
            if key in node.attrib:
                o[key] = _floats(node.get(key), 3)
        for key in ("iterations", "ls_iterations"):
            if key in node.attrib:
                o[key] = int(node.get(key))
        if "integrator" in node.attrib:
            name = node.get("integrator")
            if name == "Euler":
                o["integrator"] = INT_EULER
            elif name == "RK4":
                o["integrator"] = INT_RK4
            else:
                raise ConfigError(f"integrator {name!r} is not supported (Euler, RK4 only)")
        if node.get("cone", "pyramidal") != "pyramidal":
            raise ConfigError("only cone='pyramidal' is supported")
        if node.get("solver", "Newton") != "Newton":
            raise ConfigError("only solver='Newton' is supported")
        if int(node.get("noslip_iterations", "0")) != 0:
            raise ConfigError("noslip iterations are not supported")
        for flag in node.findall("flag"):
            for k, v in flag.attrib.items():
                default_on = k not in ("override", "energy", "fwdinv", "invdiscrete", "multiccd", "island")
                if (v == "enable") != default_on:
                    raise ConfigError(f"<flag {k}={v!r}> is not supported")

    # ---- orientation helpers
    def _angle(self, v):
        return np.deg2rad(v) if self.angle_deg else v

    def _orientation(self, attr: dict[str, str]) -> np.ndarray:
        if "quat" in attr:
            q = _floats(attr["quat"], 4)
            n = np.linalg.norm(q)
            if n < mjMINVAL:
                raise ConfigError("zero quaternion")
            return q / n
        if "axisangle" in attr:
            v = _floats(attr["axisangle"], 4)
            return axisangle_to_quat(v[:3], float(self._angle(v[3])))
        if "euler" in attr:
            e = self._angle(_floats(attr["euler"], 3))
            q = np.array([1.0, 0, 0, 0])
            for ch, ang in zip(self.eulerseq, e):
                ax = {"x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1]}[ch.lower()]
                r = axisangle_to_quat(ax, float(ang))
                q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
            return q
        if "xyaxes" in attr:
            v = _floats(attr["xyaxes"], 6)
            x = v[:3] / np.linalg.norm(v[:3])
            y = v[3:] - x * np.dot(x, v[3:])
            y = y / np.linalg.norm(y)
            z = np.cross(x, y)
            return mat_to_quat(np.stack([x, y, z], axis=1))
        if "zaxis" in attr:
            return z_to_quat(_floats(attr["zaxis"], 3))
        return np.array([1.0, 0, 0, 0])

    def _resolve(self, node: ET.Element, tag: str, childclass: str | None) -> dict[str, str]:
        cls = node.get("class") or childclass
        merged = dict(self.defaults.get(cls, tag))
        merged.update(node.attrib)
        return merged

    # ---- kinematic tree
    def _body_children(self, node: ET.Element, bid: int, childclass: str | None) -> None:
        for child in node:
            tag = child.tag
            if tag == "body":
                cc = child.get("childclass") or childclass
                body = dict(name=child.get("name"), parent=bid,
                            pos=_floats(child.get("pos"), 3, [0, 0, 0]),
                            quat=self._orientation(child.attrib), childclass=cc,
                            joints=[], geoms=[], explicit_inertial=None)
                if child.get("mocap", "false") == "true":
                    raise ConfigError("mocap bodies are not supported")
                self.bodies.append(body)
                self._body_children(child, len(self.bodies) - 1, cc)
            elif tag == "inertial":
                a = child.attrib
                if "fullinertia" in a:
                    raise ConfigError("<inertial fullinertia> is not supported")
                self.bodies[bid]["explicit_inertial"] = dict(
                    pos=_floats(a.get("pos"), 3, [0, 0, 0]), quat=self._orientation(a),
                    mass=float(a["mass"]), inertia=_floats(a.get("diaginertia"), 3))
            elif tag in ("joint", "freejoint"):
                if bid == 0:
                    raise ConfigError("joints cannot be attached to the world body")
                self._joint(child, bid, childclass)
            elif tag == "geom":
                self._geom(child, bid, childclass)
            elif tag == "site":
                a = self._resolve(child, "site", childclass)
                self.sites.append(dict(name=a.get("name"), body=bid,
                                       pos=_floats(a.get("pos"), 3, [0, 0, 0]),
                                       quat=self._orientation(a)))
            elif tag in ("camera", "light"):
                continue
            else:
                raise ConfigError(f"<{tag}> inside <body> is not supported")

    def _joint(self, node: ET.Element, bid: int, childclass: str | None) -> None:
        if node.tag == "freejoint":
            a = {"type": "free"}
            if node.get("name"):
                a["name"] = node.get("name")
        else:
            a = self._resolve(node, "joint", childclass)
        jtype = JNT_TYPES.get(a.get("type", "hinge"))
        if jtype is None:
            raise ConfigError(f"unknown joint type {a.get('type')!r}")
        if jtype == JNT_BALL:
            raise ConfigError("ball joints are not supported by the B200 path")
        axis = _floats(a.get("axis"), 3, [0, 0, 1])
        if jtype in (JNT_SLIDE, JNT_HINGE):
            n = np.linalg.norm(axis)
            if n < mjMINVAL:
                raise ConfigError("joint axis too small")
            axis = axis / n
        else:
            axis = np.array([0.0, 0, 1])
        rng = _floats(a.get("range"), 2)
        lim = a.get("limited", "auto")
        if lim == "auto":
            if rng is not None and not self.autolimits:
                raise ConfigError("joint has range but limited is unspecified and autolimits is off")
            limited = rng is not None and self.autolimits
        else:
            limited = lim == "true"
        if rng is None:
            rng = np.zeros(2)
        ref = float(a.get("ref", 0.0))
        springref = float(a.get("springref", 0.0))
        if jtype == JNT_HINGE:
            rng = self._angle(rng)
            ref = float(self._angle(ref))
            springref = float(self._angle(springref))
        if float(a.get("frictionloss", 0.0)) != 0.0:
            raise ConfigError("joint frictionloss is not supported")
        if jtype == JNT_FREE and limited:
            raise ConfigError("limited free joints are not supported")
        j = dict(name=a.get("name"), type=jtype, body=bid, pos=_floats(a.get("pos"), 3, [0, 0, 0]),
                 axis=axis, range=np.asarray(rng, float), limited=bool(limited), ref=ref,
                 springref=springref, stiffness=float(a.get("stiffness", 0.0)),
                 damping=float(a.get("damping", 0.0)), armature=float(a.get("armature", 0.0)),
                 margin=float(a.get("margin", 0.0)),
                 solref=_floats(a.get("solreflimit"), None, [0.02, 1.0]),
                 solimp=self._solimp(a.get("solimplimit")))
        if jtype == JNT_FREE:
            j["pos"] = np.zeros(3)
            if self.bodies[bid]["parent"] != 0 or self.bodies[bid]["joints"]:
                raise ConfigError("free joint must be the only joint of a top-level body")
        self.bodies[bid]["joints"].append(len(self.joints))
        self.joints.append(j)

    @staticmethod
    def _solimp(text: str | None) -> np.ndarray:
        base = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        if text is not None:
            v = _floats(text)
            base[: v.size] = v
        return base

    def _geom(self, node: ET.Element, bid: int, childclass: str | None) -> None:
        a = self._resolve(node, "geom", childclass)
        gtype = GEOM_TYPES.get(a.get("type", "sphere"))
        if gtype is None:
            raise ConfigError(f"unknown geom type {a.get('type')!r}")
        if gtype == GEOM_HFIELD:
            raise ConfigError("hfield geoms are not supported")
        size = np.zeros(3)
        sz = _floats(a.get("size"))
        if sz is not None:
            size[: sz.size] = sz
        pos = _floats(a.get("pos"), 3, [0, 0, 0])
        quat = self._orientation(a)
        if "fromto" in a:
            if gtype not in (GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_ELLIPSOID):
                raise ConfigError("fromto requires capsule, cylinder, box or ellipsoid")
            ft = _floats(a["fromto"], 6)
            vec = ft[:3] - ft[3:]  # MuJoCo: z axis points from 'to' towards 'from'
            length = np.linalg.norm(vec)
            if length < mjMINVAL:
                raise ConfigError("fromto points too close")
            half = length / 2
            if gtype in (GEOM_CAPSULE, GEOM_CYLINDER):
                size[1] = half
            else:
                size[2] = half
                size[1] = size[0]
            pos = 0.5 * (ft[:3] + ft[3:])
            quat = z_to_quat(vec)
        contype = int(a.get("contype", 1))
        conaffinity = int(a.get("conaffinity", 1))
        density = float(a.get("density", 1000.0))
        mass_attr = a.get("mass")
        vol, unit_inertia = _geom_volume_inertia(gtype, size)
        if gtype == GEOM_MESH:
            if mass_attr is None or float(mass_attr) != 0.0 or contype or conaffinity:
                raise ConfigError("mesh geoms are supported only as massless, non-colliding visuals")
            mass = 0.0
        elif mass_attr is not None:
            mass = float(mass_attr)
        else:
            mass = density * vol
        inertia = unit_inertia * mass
        friction = np.array([1.0, 0.005, 0.0001])
        fr = _floats(a.get("friction"))
        if fr is not None:
            friction[: fr.size] = fr
        if gtype == GEOM_SPHERE:
            rbound = size[0]
        elif gtype == GEOM_CAPSULE:
            rbound = size[0] + size[1]
        elif gtype == GEOM_CYLINDER:
            rbound = math.hypot(size[0], size[1])
        elif gtype == GEOM_ELLIPSOID:
            rbound = float(np.max(size))
        elif gtype == GEOM_BOX:
            rbound = float(np.linalg.norm(size))
        else:
            rbound = 0.0
        g = dict(name=a.get("name"), type=gtype, body=bid, size=size, pos=pos, quat=quat,
                 mass=mass, inertia=inertia, contype=contype, conaffinity=conaffinity,
                 condim=int(a.get("condim", 3)), friction=friction,
                 solref=_floats(a.get("solref"), None, [0.02, 1.0]), solimp=self._solimp(a.get("solimp")),
                 solmix=float(a.get("solmix", 1.0)), margin=float(a.get("margin", 0.0)),
                 gap=float(a.get("gap", 0.0)), priority=int(a.get("priority", 0)), rbound=rbound,
                 group=int(a.get("group", 0)))
        if g["condim"] not in (1, 3):
            raise ConfigError("only condim 1 and 3 are supported")
        self.bodies[bid]["geoms"].append(len(self.geoms))
        self.geoms.append(g)

    def _actuator(self, node: ET.Element) -> None:
        if node.tag not in _ACTUATOR_TAGS:
            raise ConfigError(f"actuator type <{node.tag}> is not supported")
        cls = node.get("class")
        a = dict(self.defaults.get(cls, "actuator"))
        a.pop("__shortcut__", None)
        a.update(node.attrib)
        gear = np.zeros(6)
        gear[0] = 1.0
        gv = _floats(a.get("gear"))
        if gv is not None:
            gear[:] = 0.0
            gear[: gv.size] = gv
        gain = 1.0
        bias = np.zeros(3)
        if node.tag == "motor":
            pass
        elif node.tag == "position":
            kp = float(a.get("kp", 1.0))
            kv = float(a.get("kv", 0.0))
            gain, bias = kp, np.array([0.0, -kp, -kv])
        elif node.tag == "velocity":
            kv = float(a.get("kv", 1.0))
            gain, bias = kv, np.array([0.0, 0.0, -kv])
        else:
            if a.get("dyntype", "none") != "none" or a.get("gaintype", "fixed") != "fixed":
                raise ConfigError("only dyntype=none, gaintype=fixed actuators are supported")
            gp = _floats(a.get("gainprm"))
            gain = float(gp[0]) if gp is not None else 1.0
            bt = a.get("biastype", "none")
            if bt == "affine":
                bp = _floats(a.get("biasprm"))
                if bp is not None:
                    bias[: min(3, bp.size)] = bp[:3]
            elif bt != "none":
                raise ConfigError("only biastype none/affine is supported")
        if a.get("dyntype", "none") != "none":
            raise ConfigError("actuator dynamics (activations) are not supported")
        ctrlrange = _floats(a.get("ctrlrange"), 2)
        forcerange = _floats(a.get("forcerange"), 2)

        def lim(flag: str, rng) -> bool:
            v = a.get(flag, "auto")
            if v == "auto":
                if rng is not None and not self.autolimits:
                    raise ConfigError(f"{flag} unspecified with a range and autolimits off")
                return rng is not None and self.autolimits
            return v == "true"

        act = dict(name=a.get("name"), gear=gear, gain=gain, bias=bias,
                   ctrllimited=lim("ctrllimited", ctrlrange), forcelimited=lim("forcelimited", forcerange),
                   ctrlrange=ctrlrange if ctrlrange is not None else np.zeros(2),
                   forcerange=forcerange if forcerange is not None else np.zeros(2),
                   group=int(a.get("group", 0)))
        if "joint" in a:
            act["trntype"], act["target"] = TRN_JOINT, a["joint"]
        elif "site" in a:
            if "refsite" in a:
                raise ConfigError("refsite transmissions are not supported")
            act["trntype"], act["target"] = TRN_SITE, a["site"]
        else:
            raise ConfigError("actuator needs a joint= or site= transmission")
        self.actuators.append(act)

    def _tendon(self, node: ET.Element) -> None:
        if node.tag != "fixed":
            raise ConfigError("only <fixed> tendons are supported")
        a = dict(self.defaults.get(node.get("class"), "tendon"))
        a.update(node.attrib)
        rng = _floats(a.get("range"), 2)
        lim = a.get("limited", "auto")
        limited = (rng is not None and self.autolimits) if lim == "auto" else lim == "true"
        if float(a.get("frictionloss", 0.0)) != 0.0:
            raise ConfigError("tendon frictionloss is not supported")
        sl = _floats(a.get("springlength"))
        if sl is None:
            sl = np.array([-1.0, -1.0])
        elif sl.size == 1:
            sl = np.array([sl[0], sl[0]])
        t = dict(name=a.get("name"), limited=bool(limited), range=rng if rng is not None else np.zeros(2),
                 margin=float(a.get("margin", 0.0)),
                 solref=_floats(a.get("solreflimit"), None, [0.02, 1.0]),
                 solimp=self._solimp(a.get("solimplimit")),
                 stiffness=float(a.get("stiffness", 0.0)), damping=float(a.get("damping", 0.0)),
                 springlength=sl, wraps=[(j.get("joint"), float(j.get("coef", 1.0))) for j in node.findall("joint")])
        self.tendons.append(t)


def _geom_volume_inertia(gtype: int, s: np.ndarray) -> tuple[float, np.ndarray]:
    """Volume and diagonal inertia per unit mass in the geom frame (MuJoCo mjCGeom::SetInertia)."""
    if gtype == GEOM_SPHERE:
        r = s[0]
        return 4.0 / 3.0 * math.pi * r ** 3, np.full(3, 0.4 * r * r)
    if gtype == GEOM_CAPSULE:
        r, h = s[0], 2 * s[1]
        vol = math.pi * (r * r * h + 4.0 / 3.0 * r ** 3)
        sm = 4 * r / (4 * r + 3 * h)  # mass fraction of the two hemispheres
        cm = 1.0 - sm
        ixx = cm * (3 * r * r + h * h) / 12 + 0.4 * sm * r * r + sm * h * (3 * r + 2 * h) / 8
        izz = cm * r * r / 2 + 0.4 * sm * r * r
        return vol, np.array([ixx, ixx, izz])
    if gtype == GEOM_CYLINDER:
        r, h = s[0], 2 * s[1]
        return math.pi * r * r * h, np.array([(3 * r * r + h * h) / 12, (3 * r * r + h * h) / 12, r * r / 2])
    if gtype == GEOM_ELLIPSOID:
        a, b, c = s
        return 4.0 / 3.0 * math.pi * a * b * c, np.array([(b * b + c * c) / 5, (a * a + c * c) / 5, (a * a + b * b) / 5])
    if gtype == GEOM_BOX:
        a, b, c = s
        return 8 * a * b * c, np.array([(b * b + c * c) / 3, (a * a + c * c) / 3, (a * a + b * b) / 3])
    return 0.0, np.zeros(3)


def _body_inertial(body: dict, geoms: list[dict]) -> tuple[float, np.ndarray, np.ndarray, np.ndarray]:
    """(mass, ipos, iquat, diag inertia) of a body from its geoms (MuJoCo mjCBody::GeomFrame)."""
    ex = body["explicit_inertial"]
    if ex is not None:
        return ex["mass"], ex["pos"], ex["quat"], ex["inertia"]
    gs = [geoms[g] for g in body["geoms"] if geoms[g]["mass"] > 0]
    mass = sum(g["mass"] for g in gs)
    if mass < mjMINVAL:
        return 0.0, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3)
    com = sum(g["mass"] * g["pos"] for g in gs) / mass
    I = np.zeros((3, 3))
    for g in gs:
        R = quat_to_mat(g["quat"])
        d = g["pos"] - com
        I += R @ np.diag(g["inertia"]) @ R.T + g["mass"] * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    offdiag = abs(I[0, 1]) + abs(I[0, 2]) + abs(I[1, 2])
    if len(gs) == 1:
        # single geom: inertial frame is the geom frame (no eigen-decomposition upstream)
        return mass, com, gs[0]["quat"].copy(), gs[0]["inertia"].copy()
    if offdiag < 1e-14 * max(1.0, np.trace(I)):
        w, V = np.diag(I).copy(), np.eye(3)
    else:
        w, V = np.linalg.eigh(I)
    order = np.argsort(-w, kind="stable")  # principal moments sorted descending
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return mass, com, mat_to_quat(V), w


def _mix_contact(g1: dict, g2: dict) -> dict:
    """mj_contactParam for a geom pair (equal priority mixes; higher priority wins)."""
    if g1["priority"] != g2["priority"]:
        g = g1 if g1["priority"] > g2["priority"] else g2
        condim, solref, solimp, fri = g["condim"], g["solref"].copy(), g["solimp"].copy(), g["friction"].copy()
    else:
        condim = max(g1["condim"], g2["condim"])
        m1, m2 = g1["solmix"], g2["solmix"]
        if m1 >= mjMINVAL and m2 >= mjMINVAL:
            mix = m1 / (m1 + m2)
        elif m1 < mjMINVAL and m2 < mjMINVAL:
            mix = 0.5
        elif m1 < mjMINVAL:
            mix = 0.0
        else:
            mix = 1.0
        if g1["solref"][0] > 0 and g2["solref"][0] > 0:
            solref = mix * g1["solref"] + (1 - mix) * g2["solref"]
        else:
            solref = np.minimum(g1["solref"], g2["solref"])
        solimp = mix * g1["solimp"] + (1 - mix) * g2["solimp"]
        fri = np.maximum(g1["friction"], g2["friction"])
    friction5 = np.array([fri[0], fri[0], fri[1], fri[2], fri[2]])
    return dict(dim=condim, solref=solref, solimp=solimp, friction=friction5,
                margin=max(g1["margin"], g2["margin"]), gap=max(g1["gap"], g2["gap"]))


def compile_root(root: ET.Element, basedir: str) -> dict:
    c = _Compiler()
    c.load(root, basedir)
    return _finalize(c)


def compile_xml_path(path: str) -> dict:
    try:
        root = ET.parse(path).getroot()
    except (OSError, ET.ParseError) as exc:
        raise ConfigError(f"cannot load MJCF {path}: {exc}") from exc
    return compile_root(root, os.path.dirname(os.path.abspath(path)))


def compile_xml_string(text: str, basedir: str = ".") -> dict:
    try:
        root = ET.fromstring(text)
    except ET.ParseError as exc:
        raise ConfigError(f"cannot parse MJCF string: {exc}") from exc
    return compile_root(root, basedir)


def _finalize(c: _Compiler) -> dict:
    nbody, njnt, ngeom, nsite = len(c.bodies), len(c.joints), len(c.geoms), len(c.sites)
    m: dict = dict(model_name=c.model_name)
    m.update(nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=nsite)
    # --- addresses
    qadr, dadr = [], []
    nq = nv = 0
    for j in c.joints:
        qadr.append(nq)
        dadr.append(nv)
        if j["type"] == JNT_FREE:
            nq, nv = nq + 7, nv + 6
        else:
            nq, nv = nq + 1, nv + 1
    m.update(nq=nq, nv=nv)
    parent = np.array([max(b["parent"], 0) for b in c.bodies], dtype=np.int32)
    body_jntnum = np.array([len(b["joints"]) for b in c.bodies], dtype=np.int32)
    body_jntadr = np.array([b["joints"][0] if b["joints"] else -1 for b in c.bodies], dtype=np.int32)
    body_dofnum = np.zeros(nbody, np.int32)
    body_dofadr = np.full(nbody, -1, np.int32)
    dof_bodyid = np.zeros(nv, np.int32)
    dof_jntid = np.zeros(nv, np.int32)
    for jid, j in enumerate(c.joints):
        n = 6 if j["type"] == JNT_FREE else 1
        b = j["body"]
        if body_dofadr[b] < 0:
            body_dofadr[b] = dadr[jid]
        body_dofnum[b] += n
        dof_bodyid[dadr[jid]: dadr[jid] + n] = b
        dof_jntid[dadr[jid]: dadr[jid] + n] = jid
    rootid = np.zeros(nbody, np.int32)
    weldid = np.zeros(nbody, np.int32)
    for i in range(1, nbody):
        rootid[i] = i if parent[i] == 0 else rootid[parent[i]]
        weldid[i] = i if body_jntnum[i] > 0 else weldid[parent[i]]
    dof_parentid = np.full(nv, -1, np.int32)
    for i in range(1, nbody):
        if body_dofnum[i] == 0:
            continue
        p = parent[i]
        while p > 0 and body_dofnum[p] == 0:
            p = parent[p]
        last = body_dofadr[p] + body_dofnum[p] - 1 if p > 0 else -1
        for d in range(body_dofadr[i], body_dofadr[i] + body_dofnum[i]):
            dof_parentid[d] = last
            last = d
    # --- inertial properties
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile([1.0, 0, 0, 0], (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    for i, b in enumerate(c.bodies):
        if i == 0:
            continue
        mass, ipos, iquat, inertia = _body_inertial(b, c.geoms)
        if body_dofnum[i] > 0 or weldid[i] > 0:
            pass
        body_mass[i], body_ipos[i], body_iquat[i], body_inertia[i] = mass, ipos, iquat, inertia
    for i in range(1, nbody):
        if weldid[i] > 0:
            # moving subtree must carry mass somewhere; checked per weld group below
            pass
    subtreemass = body_mass.copy()
    for i in range(nbody - 1, 0, -1):
        subtreemass[parent[i]] += subtreemass[i]
    for i in range(1, nbody):
        if body_jntnum[i] > 0 and subtreemass[i] < mjMINVAL:
            raise ConfigError(f"moving body {c.bodies[i]['name']!r} has no mass in its subtree")
    # --- qpos0 / springs
    qpos0 = np.zeros(nq)
    qpos_spring = np.zeros(nq)
    for jid, j in enumerate(c.joints):
        a = qadr[jid]
        if j["type"] == JNT_FREE:
            b = c.bodies[j["body"]]
            qpos0[a: a + 3] = b["pos"]
            qpos0[a + 3: a + 7] = b["quat"]
            qpos_spring[a: a + 7] = qpos0[a: a + 7]
        else:
            qpos0[a] = j["ref"]
            qpos_spring[a] = j["springref"]
    m.update(
        body_parentid=parent, body_rootid=rootid, body_weldid=weldid, body_jntnum=body_jntnum,
        body_jntadr=body_jntadr, body_dofnum=body_dofnum, body_dofadr=body_dofadr,
        body_pos=np.array([b["pos"] for b in c.bodies]), body_quat=np.array([b["quat"] for b in c.bodies]),
        body_ipos=body_ipos, body_iquat=body_iquat, body_mass=body_mass, body_subtreemass=subtreemass,
        body_inertia=body_inertia,
        jnt_type=np.array([j["type"] for j in c.joints], np.int32), jnt_qposadr=np.array(qadr, np.int32),
        jnt_dofadr=np.array(dadr, np.int32), jnt_bodyid=np.array([j["body"] for j in c.joints], np.int32),
        jnt_limited=np.array([j["limited"] for j in c.joints], np.int32),
        jnt_pos=np.array([j["pos"] for j in c.joints]).reshape(njnt, 3),
        jnt_axis=np.array([j["axis"] for j in c.joints]).reshape(njnt, 3),
        jnt_stiffness=np.array([j["stiffness"] for j in c.joints]),
        jnt_range=np.array([j["range"] for j in c.joints]).reshape(njnt, 2),
        jnt_margin=np.array([j["margin"] for j in c.joints]),
        jnt_solref=np.array([j["solref"] for j in c.joints]).reshape(njnt, 2),
        jnt_solimp=np.array([j["solimp"] for j in c.joints]).reshape(njnt, 5),
        qpos0=qpos0, qpos_spring=qpos_spring,
        dof_bodyid=dof_bodyid, dof_jntid=dof_jntid, dof_parentid=dof_parentid,
        dof_armature=np.array([c.joints[j]["armature"] for j in dof_jntid]),
        dof_damping=np.array([c.joints[j]["damping"] for j in dof_jntid]),
        geom_type=np.array([g["type"] for g in c.geoms], np.int32),
        geom_bodyid=np.array([g["body"] for g in c.geoms], np.int32),
        geom_size=np.array([g["size"] for g in c.geoms]).reshape(ngeom, 3),
        geom_rbound=np.array([g["rbound"] for g in c.geoms]),
        geom_pos=np.array([g["pos"] for g in c.geoms]).reshape(ngeom, 3),
        geom_quat=np.array([g["quat"] for g in c.geoms]).reshape(ngeom, 4),
        geom_contype=np.array([g["contype"] for g in c.geoms], np.int32),
        geom_conaffinity=np.array([g["conaffinity"] for g in c.geoms], np.int32),
        geom_condim=np.array([g["condim"] for g in c.geoms], np.int32),
        geom_friction=np.array([g["friction"] for g in c.geoms]).reshape(ngeom, 3),
        geom_group=np.array([g["group"] for g in c.geoms], np.int32),
        site_bodyid=np.array([s["body"] for s in c.sites], np.int32),
        site_pos=np.array([s["pos"] for s in c.sites]).reshape(nsite, 3),
        site_quat=np.array([s["quat"] for s in c.sites]).reshape(nsite, 4),
    )
    # --- names
    names = dict(
        body=[b["name"] for b in c.bodies], joint=[j["name"] for j in c.joints],
        geom=[g["name"] for g in c.geoms], site=[s["name"] for s in c.sites],
        actuator=[a["name"] for a in c.actuators], tendon=[t["name"] for t in c.tendons],
        key=[k.get("name") for k in c.keys],
    )
    m["names"] = names

    def lookup(kind: str, name: str) -> int:
        try:
            return names[kind].index(name)
        except ValueError:
            raise ConfigError(f"{kind} {name!r} referenced but not defined") from None

    # --- tendons
    nt = len(c.tendons)
    wraps = [(lookup("joint", jn), coef) for t in c.tendons for jn, coef in t["wraps"]]
    for jid, _ in wraps:
        if c.joints[jid]["type"] not in (JNT_HINGE, JNT_SLIDE):
            raise ConfigError("fixed tendons may only reference hinge/slide joints")
    tadr, cursor = [], 0
    for t in c.tendons:
        tadr.append(cursor)
        cursor += len(t["wraps"])
    lengthspring = np.array([t["springlength"] for t in c.tendons]).reshape(nt, 2)
    for k, t in enumerate(c.tendons):
        if lengthspring[k, 0] < 0:  # springlength=-1: use length at qpos0
            L0 = sum(coef * qpos0[qadr[lookup("joint", jn)]] for jn, coef in t["wraps"])
            lengthspring[k] = L0
    m.update(
        ntendon=nt, nwrap=len(wraps), tendon_adr=np.array(tadr, np.int32),
        tendon_num=np.array([len(t["wraps"]) for t in c.tendons], np.int32),
        tendon_limited=np.array([t["limited"] for t in c.tendons], np.int32),
        tendon_range=np.array([t["range"] for t in c.tendons]).reshape(nt, 2),
        tendon_margin=np.array([t["margin"] for t in c.tendons]),
        tendon_solref=np.array([t["solref"] for t in c.tendons]).reshape(nt, 2),
        tendon_solimp=np.array([t["solimp"] for t in c.tendons]).reshape(nt, 5),
        tendon_stiffness=np.array([t["stiffness"] for t in c.tendons]),
        tendon_damping=np.array([t["damping"] for t in c.tendons]),
        tendon_lengthspring=lengthspring,
        wrap_jntid=np.array([w[0] for w in wraps], np.int32), wrap_coef=np.array([w[1] for w in wraps]),
    )
    # --- actuators
    nu = len(c.actuators)
    trnid = []
    for a in c.actuators:
        if a["trntype"] == TRN_JOINT:
            jid = lookup("joint", a["target"])
            if c.joints[jid]["type"] not in (JNT_HINGE, JNT_SLIDE):
                raise ConfigError("joint transmission requires a hinge or slide joint")
            trnid.append(jid)
        else:
            trnid.append(lookup("site", a["target"]))
    m.update(
        nu=nu, actuator_trntype=np.array([a["trntype"] for a in c.actuators], np.int32),
        actuator_trnid=np.array(trnid, np.int32),
        actuator_ctrllimited=np.array([a["ctrllimited"] for a in c.actuators], np.int32),
        actuator_forcelimited=np.array([a["forcelimited"] for a in c.actuators], np.int32),
        actuator_disabled=np.zeros(nu, np.int32),
        actuator_gear=np.array([a["gear"] for a in c.actuators]).reshape(nu, 6),
        actuator_ctrlrange=np.array([a["ctrlrange"] for a in c.actuators]).reshape(nu, 2),
        actuator_forcerange=np.array([a["forcerange"] for a in c.actuators]).reshape(nu, 2),
        actuator_gainprm=np.array([a["gain"] for a in c.actuators]),
        actuator_biasprm=np.array([a["bias"] for a in c.actuators]).reshape(nu, 3),
        actuator_group=np.array([a["group"] for a in c.actuators], np.int32),
    )
    # --- keyframes
    nkey = len(c.keys)
    key_qpos = np.tile(qpos0, (nkey, 1)).reshape(nkey, nq)
    key_qvel = np.zeros((nkey, nv))
    key_ctrl = np.zeros((nkey, nu))
    key_time = np.zeros(nkey)
    for k, key in enumerate(c.keys):
        for attr, arr, n in (("qpos", key_qpos, nq), ("qvel", key_qvel, nv), ("ctrl", key_ctrl, nu)):
            if attr in key:
                arr[k] = _floats(key[attr], n)
        if "act" in key and key["act"].strip():
            raise ConfigError("keyframe act is not supported (na=0)")
        key_time[k] = float(key.get("time", 0.0))
    m.update(nkey=nkey, key_qpos=key_qpos, key_qvel=key_qvel, key_ctrl=key_ctrl, key_time=key_time)
    # --- options
    o = c.opt
    m.update(timestep=o["timestep"], gravity=o["gravity"], wind=o["wind"], density=o["density"],
             viscosity=o["viscosity"], tolerance=o["tolerance"], ls_tolerance=o["ls_tolerance"],
             impratio=o["impratio"], integrator=o["integrator"], iterations=o["iterations"],
             ls_iterations=o["ls_iterations"], disableflags=0, nmocap_unused=0,
             has_fluid=int(o["density"] > 0 or o["viscosity"] > 0),
             has_dofdamping=int(bool(np.any(m["dof_damping"] > 0))))
    if o["impratio"] != 1.0:
        raise ConfigError("impratio != 1 is not supported")
    # --- static collision filter + contact parameter mixing
    excl = set()
    for b1, b2 in c.excludes:
        i1, i2 = lookup("body", b1), lookup("body", b2)
        excl.add((min(i1, i2), max(i1, i2)))
    pairs = []
    for b1 in range(nbody):
        for b2 in range(b1 + 1, nbody):
            w1, w2 = weldid[b1], weldid[b2]
            if w1 == w2:
                continue
            wp1, wp2 = weldid[parent[w1]], weldid[parent[w2]]
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue
            if (b1, b2) in excl:
                continue
            for g1 in c.bodies[b1]["geoms"]:
                for g2 in c.bodies[b2]["geoms"]:
                    G1, G2 = c.geoms[g1], c.geoms[g2]
                    if not ((G1["contype"] & G2["conaffinity"]) or (G2["contype"] & G1["conaffinity"])):
                        continue
                    a, b = (g1, g2) if G1["type"] <= G2["type"] else (g2, g1)
                    tp = (c.geoms[a]["type"], c.geoms[b]["type"])
                    if tp not in SUPPORTED_PAIRS:
                        raise ConfigError(
                            f"collision between geom types {tp} (geoms {a},{b}) is not supported by the B200 path")
                    mix = _mix_contact(c.geoms[a], c.geoms[b])
                    pairs.append((a, b, mix))
    npair = len(pairs)
    m.update(
        npair=npair, pair_geom1=np.array([p[0] for p in pairs], np.int32),
        pair_geom2=np.array([p[1] for p in pairs], np.int32),
        pair_dim=np.array([p[2]["dim"] for p in pairs], np.int32),
        pair_margin=np.array([p[2]["margin"] for p in pairs]),
        pair_gap=np.array([p[2]["gap"] for p in pairs]),
        pair_friction=np.array([p[2]["friction"] for p in pairs]).reshape(npair, 5),
        pair_solref=np.array([p[2]["solref"] for p in pairs]).reshape(npair, 2),
        pair_solimp=np.array([p[2]["solimp"] for p in pairs]).reshape(npair, 5),
    )
    # --- sensors (subset: the kinds the reference's models and tests use, plus their velocity/frame siblings)
    sens = []
    adr = 0
    for sd in c.sensors:
        tag = sd["tag"]
        if tag not in SENSOR_TYPES:
            raise ConfigError(f"<sensor><{tag}> is not supported by the B200 path (supported: {sorted(SENSOR_TYPES)})")
        code, dim = SENSOR_TYPES[tag]
        if sd.get("reftype") or sd.get("refname"):
            raise ConfigError(f"sensor {sd.get('name')!r}: reference frames (reftype/refname) are not supported")
        if tag in ("jointpos", "jointvel"):
            jid = lookup("joint", sd.get("joint"))
            if int(m["jnt_type"][jid]) in (JNT_FREE, JNT_BALL):
                raise ConfigError(f"sensor {sd.get('name')!r}: {tag} needs a slide or hinge joint")
            objtype, objid = 3, jid
        elif tag in ("gyro", "accelerometer", "velocimeter"):
            objtype, objid = 6, lookup("site", sd.get("site"))
        else:  # framepos / framequat
            ot = sd.get("objtype")
            kinds = dict(body=(1, "body"), xbody=(2, "body"), geom=(5, "geom"), site=(6, "site"))
            if ot not in kinds:
                raise ConfigError(f"sensor {sd.get('name')!r}: objtype {ot!r} is not supported")
            objtype, objid = kinds[ot][0], lookup(kinds[ot][1], sd.get("objname"))
        sens.append(dict(name=sd.get("name"), type=code, objtype=objtype, objid=objid, adr=adr, dim=dim,
                         cutoff=float(sd.get("cutoff", 0.0))))
        adr += dim
    names["sensor"] = [x["name"] for x in sens]
    m.update(nsensor=len(sens), nsensordata=adr,
             sensor_type=np.array([x["type"] for x in sens], np.int32),
             sensor_objtype=np.array([x["objtype"] for x in sens], np.int32),
             sensor_objid=np.array([x["objid"] for x in sens], np.int32),
             sensor_adr=np.array([x["adr"] for x in sens], np.int32),
             sensor_dim=np.array([x["dim"] for x in sens], np.int32),
             sensor_cutoff=np.array([x["cutoff"] for x in sens], float))
    _set_const(m)
    return m


def _set_const(m: dict) -> None:
    """mj_setConst subset: M(qpos0) via body Jacobians, invweight0 family, meaninertia."""
    nbody, nv, njnt = m["nbody"], m["nv"], m["njnt"]
    parent = m["body_parentid"]
    xpos = np.zeros((nbody, 3))
    xquat = np.tile([1.0, 0, 0, 0], (nbody, 1))
    for i in range(1, nbody):
        R = quat_to_mat(xquat[parent[i]])
        xpos[i] = xpos[parent[i]] + R @ m["body_pos"][i]
        xquat[i] = quat_mul(xquat[parent[i]], m["body_quat"][i])
        xquat[i] /= np.linalg.norm(xquat[i])
    xmat = np.array([quat_to_mat(q) for q in xquat])
    xipos = np.array([xpos[i] + xmat[i] @ m["body_ipos"][i] for i in range(nbody)])
    ximat = np.array([quat_to_mat(quat_mul(xquat[i], m["body_iquat"][i])) for i in range(nbody)])
    # per-dof motion axes at qpos0
    dof_rot = np.zeros((nv, 3))  # angular part
    dof_lin_at = []  # function giving linear velocity at point p
    anchors = np.zeros((nv, 3))
    kinds = []
    for j in range(njnt):
        b = m["jnt_bodyid"][j]
        d = m["jnt_dofadr"][j]
        t = m["jnt_type"][j]
        if t == JNT_FREE:
            for k in range(3):
                kinds.append(("lin", np.eye(3)[k]))
            for k in range(3):
                dof_rot[d + 3 + k] = xmat[b][:, k]
                anchors[d + 3 + k] = xpos[b]
                kinds.append(("rot", xmat[b][:, k]))
        elif t == JNT_SLIDE:
            kinds.append(("lin", xmat[b] @ m["jnt_axis"][j]))
        else:
            ax = xmat[b] @ m["jnt_axis"][j]
            dof_rot[d] = ax
            anchors[d] = xpos[b] + xmat[b] @ m["jnt_pos"][j]
            kinds.append(("rot", ax))

    def chain(body: int) -> list[int]:
        out = []
        while body > 0 and m["body_dofnum"][body] == 0:
            body = parent[body]
        if body == 0:
            return out
        d = m["body_dofadr"][body] + m["body_dofnum"][body] - 1
        while d >= 0:
            out.append(d)
            d = m["dof_parentid"][d]
        return out

    def jac(body: int, point: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
        for d in chain(body):
            kind, ax = kinds[d]
            if kind == "lin":
                jp[:, d] = ax
            else:
                jr[:, d] = ax
                jp[:, d] = np.cross(ax, point - anchors[d])
        return jp, jr

    M = np.diag(np.asarray(m["dof_armature"], float)) if nv else np.zeros((0, 0))
    for i in range(1, nbody):
        if m["body_mass"][i] <= 0:
            continue
        jp, jr = jac(i, xipos[i])
        Iw = ximat[i] @ np.diag(m["body_inertia"][i]) @ ximat[i].T
        M = M + m["body_mass"][i] * jp.T @ jp + jr.T @ Iw @ jr
    m["meaninertia"] = float(np.mean(np.diag(M))) if nv else 1.0
    Minv = np.linalg.inv(M) if nv else M
    inv_body = np.zeros((nbody, 2))
    for i in range(1, nbody):
        if m["body_weldid"][i] == 0:
            continue
        jp, jr = jac(i, xipos[i])
        J = np.vstack([jp, jr])
        A = J @ Minv @ J.T
        inv_body[i, 0] = max(mjMINVAL, (A[0, 0] + A[1, 1] + A[2, 2]) / 3)
        inv_body[i, 1] = max(mjMINVAL, (A[3, 3] + A[4, 4] + A[5, 5]) / 3)
    inv_dof = np.zeros(nv)
    for j in range(njnt):
        d = m["jnt_dofadr"][j]
        if m["jnt_type"][j] == JNT_FREE:
            inv_dof[d: d + 3] = np.mean(np.diag(Minv)[d: d + 3])
            inv_dof[d + 3: d + 6] = np.mean(np.diag(Minv)[d + 3: d + 6])
        else:
            inv_dof[d] = Minv[d, d]
    inv_ten = np.zeros(m["ntendon"])
    for t in range(m["ntendon"]):
        J = np.zeros(nv)
        for w in range(m["tendon_adr"][t], m["tendon_adr"][t] + m["tendon_num"][t]):
            J[m["jnt_dofadr"][m["wrap_jntid"][w]]] += m["wrap_coef"][w]
        inv_ten[t] = J @ Minv @ J
    m.update(body_invweight0=inv_body, dof_invweight0=inv_dof, tendon_invweight0=inv_ten)
    m["_M0"] = M
