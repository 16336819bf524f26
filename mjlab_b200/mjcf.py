"""Minimal MJCF -> model-description compiler (host side, setup time).

The reference builds its model with MuJoCo's C compiler
(`scene.compile()` src/mjlab/scene/scene.py:47-48, `MjSpec.from_file`
src/mjlab/asset_zoo/robots/unitree_g1/g1_constants.py:35-38).  MuJoCo is not
installable here (no wheel, no network), so this module restates the subset of
the MJCF compiler the hot path's models need:

  * bodies / free + hinge (+slide) joints / explicit <inertial> or inertia
    inferred from primitive geoms
  * plane, sphere, capsule, box, cylinder, ellipsoid geoms (mesh geoms are kept
    as pose-only visual geoms; their mesh-frame recentring is NOT reproduced)
  * sites, <contact><exclude>, default classes / childclass
  * gyro / velocimeter / accelerometer / subtreeangmom / jointpos / jointvel /
    framepos / framequat sensors, and the mjSENS_CONTACT sensor the reference
    adds programmatically (src/mjlab/sensor/contact_sensor.py:474-533)
  * general actuators with FIXED gain / NONE|AFFINE bias and JOINT transmission
    (the <position>/<motor>/<velocity> shortcuts of
    src/mjlab/utils/spec.py:91-202)

`Spec` mirrors the handful of MjSpec editing operations mjlab performs
(attach with prefix, add actuator / sensor / key / body / geom / site), and
`Spec.compile()` returns a `ModelDesc`: plain numpy arrays named after the
`MjModel` fields (typings/mujoco/_structs.pyi:916).
"""

from __future__ import annotations

import copy
import math
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np

# ---------------------------------------------------------------------------
# Enums (values follow MuJoCo's public enums, typings/mujoco/_enums.pyi).
# ---------------------------------------------------------------------------
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
GEOM_TYPES = {
  "plane": GEOM_PLANE, "hfield": GEOM_HFIELD, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
  "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH,
}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
JNT_TYPES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
INT_EULER, INT_RK4, INT_IMPLICIT, INT_IMPLICITFAST = range(4)
CONE_PYRAMIDAL, CONE_ELLIPTIC = range(2)
SOL_PGS, SOL_CG, SOL_NEWTON = range(3)
GAIN_FIXED, GAIN_AFFINE = 0, 1
BIAS_NONE, BIAS_AFFINE = 0, 1
OBJ_UNKNOWN, OBJ_BODY, OBJ_XBODY, OBJ_JOINT, OBJ_DOF, OBJ_GEOM, OBJ_SITE = 0, 1, 2, 3, 4, 5, 6

# Sensor types we support (values follow mjtSensor ordering where it matters
# only internally; the engine consumes these codes, not MuJoCo's).
SENS_ACCELEROMETER = 1
SENS_VELOCIMETER = 2
SENS_GYRO = 3
SENS_JOINTPOS = 8
SENS_JOINTVEL = 9
SENS_FRAMEPOS = 25
SENS_FRAMEQUAT = 26
SENS_FRAMELINVEL = 30
SENS_FRAMEANGVEL = 31
SENS_SUBTREECOM = 34
SENS_SUBTREELINVEL = 35
SENS_SUBTREEANGMOM = 36
SENS_CONTACT = 42
SENSOR_TAGS = {
  "accelerometer": (SENS_ACCELEROMETER, 3, OBJ_SITE, "site"),
  "velocimeter": (SENS_VELOCIMETER, 3, OBJ_SITE, "site"),
  "gyro": (SENS_GYRO, 3, OBJ_SITE, "site"),
  "jointpos": (SENS_JOINTPOS, 1, OBJ_JOINT, "joint"),
  "jointvel": (SENS_JOINTVEL, 1, OBJ_JOINT, "joint"),
  "subtreecom": (SENS_SUBTREECOM, 3, OBJ_BODY, "body"),
  "subtreelinvel": (SENS_SUBTREELINVEL, 3, OBJ_BODY, "body"),
  "subtreeangmom": (SENS_SUBTREEANGMOM, 3, OBJ_BODY, "body"),
  "framepos": (SENS_FRAMEPOS, 3, None, None),
  "framequat": (SENS_FRAMEQUAT, 4, None, None),
  "framelinvel": (SENS_FRAMELINVEL, 3, None, None),
  "frameangvel": (SENS_FRAMEANGVEL, 3, None, None),
}
CONTACT_FIELD_DIMS = (1, 3, 3, 1, 3, 3, 3)  # found force torque dist pos normal tangent

MINVAL = 1e-15


# ---------------------------------------------------------------------------
# Small quaternion helpers (w, x, y, z).
# ---------------------------------------------------------------------------
def quat_mul(a, b):
  a = np.asarray(a, float)
  b = np.asarray(b, float)
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
  """Rotation matrix -> unit quaternion (w>=0 branch not enforced)."""
  m = np.asarray(m, float)
  tr = m[0, 0] + m[1, 1] + m[2, 2]
  if tr > 0:
    s = math.sqrt(tr + 1.0) * 2
    q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
  elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
    s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
    q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
  elif m[1, 1] > m[2, 2]:
    s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
    q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
  else:
    s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
    q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
  q = np.array(q)
  return q / np.linalg.norm(q)


def quat_z2vec(vec):
  """Quaternion rotating +z onto `vec` (MJCF `fromto` / `zaxis` convention)."""
  v = np.asarray(vec, float)
  n = np.linalg.norm(v)
  if n < 1e-10:
    return np.array([1.0, 0, 0, 0])
  v = v / n
  axis = np.cross([0.0, 0.0, 1.0], v)
  s = np.linalg.norm(axis)
  if s < 1e-10:
    axis = np.array([1.0, 0.0, 0.0])
  else:
    axis = axis / s
  ang = math.atan2(s, v[2])
  return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


def _floats(s: str | None, n: int | None = None, default=None):
  if s is None:
    return None if default is None else np.array(default, float)
  v = np.array([float(t) for t in s.split()], float)
  if n is not None and len(v) < n and default is not None:
    out = np.array(default, float)
    out[: len(v)] = v
    return out
  return v


# ---------------------------------------------------------------------------
# Spec element records.
# ---------------------------------------------------------------------------
@dataclass
class SGeom:
  name: str = ""
  type: int = GEOM_SPHERE
  pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
  quat: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
  size: np.ndarray = field(default_factory=lambda: np.zeros(3))
  contype: int = 1
  conaffinity: int = 1
  condim: int = 3
  priority: int = 0
  friction: np.ndarray = field(default_factory=lambda: np.array([1.0, 0.005, 0.0001]))
  solref: np.ndarray = field(default_factory=lambda: np.array([0.02, 1.0]))
  solimp: np.ndarray = field(default_factory=lambda: np.array([0.9, 0.95, 0.001, 0.5, 2.0]))
  solmix: float = 1.0
  margin: float = 0.0
  gap: float = 0.0
  density: float = 1000.0
  mass: float | None = None
  group: int = 0
  mesh: str = ""
  rgba: np.ndarray = field(default_factory=lambda: np.array([0.5, 0.5, 0.5, 1.0]))


@dataclass
class SSite:
  name: str = ""
  pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
  quat: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
  size: np.ndarray = field(default_factory=lambda: np.array([0.005, 0.005, 0.005]))
  group: int = 0


@dataclass
class SJoint:
  name: str = ""
  type: int = JNT_HINGE
  pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
  axis: np.ndarray = field(default_factory=lambda: np.array([0.0, 0, 1.0]))
  range: np.ndarray = field(default_factory=lambda: np.zeros(2))
  limited: int = 2  # 0 false, 1 true, 2 auto
  armature: float = 0.0
  damping: float = 0.0
  stiffness: float = 0.0
  frictionloss: float = 0.0
  margin: float = 0.0
  ref: float = 0.0
  springref: float = 0.0
  solref: np.ndarray = field(default_factory=lambda: np.array([0.02, 1.0]))
  solimp: np.ndarray = field(default_factory=lambda: np.array([0.9, 0.95, 0.001, 0.5, 2.0]))
  solreffriction: np.ndarray = field(default_factory=lambda: np.array([0.02, 1.0]))
  solimpfriction: np.ndarray = field(default_factory=lambda: np.array([0.9, 0.95, 0.001, 0.5, 2.0]))


@dataclass
class SBody:
  name: str = ""
  pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
  quat: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
  mocap: bool = False
  # explicit inertial (None -> infer from geoms)
  ipos: np.ndarray | None = None
  iquat: np.ndarray | None = None
  mass: float | None = None
  inertia: np.ndarray | None = None
  joints: list[SJoint] = field(default_factory=list)
  geoms: list[SGeom] = field(default_factory=list)
  sites: list[SSite] = field(default_factory=list)
  children: list["SBody"] = field(default_factory=list)

  def add_body(self, **kw) -> "SBody":
    b = SBody(**kw)
    self.children.append(b)
    return b

  def add_geom(self, **kw) -> SGeom:
    g = SGeom(**_np_kw(kw))
    self.geoms.append(g)
    return g

  def add_site(self, **kw) -> SSite:
    s = SSite(**_np_kw(kw))
    self.sites.append(s)
    return s

  def add_joint(self, **kw) -> SJoint:
    j = SJoint(**_np_kw(kw))
    self.joints.append(j)
    return j


def _np_kw(kw):
  return {k: (np.array(v, float) if isinstance(v, (tuple, list)) else v) for k, v in kw.items()}


@dataclass
class SActuator:
  name: str = ""
  target: str = ""  # joint name
  gaintype: int = GAIN_FIXED
  biastype: int = BIAS_NONE
  gainprm: np.ndarray = field(default_factory=lambda: np.array([1.0] + [0.0] * 9))
  biasprm: np.ndarray = field(default_factory=lambda: np.zeros(10))
  gear: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0, 0, 0]))
  forcelimited: bool = False
  forcerange: np.ndarray = field(default_factory=lambda: np.zeros(2))
  ctrllimited: bool = False
  ctrlrange: np.ndarray = field(default_factory=lambda: np.zeros(2))


@dataclass
class SSensor:
  name: str = ""
  type: int = 0
  objtype: int = OBJ_UNKNOWN
  objname: str = ""
  reftype: int = OBJ_UNKNOWN
  refname: str = ""
  intprm: tuple[int, int, int] = (0, 0, 0)
  dim: int = 0
  cutoff: float = 0.0


@dataclass
class SKey:
  name: str = ""
  qpos: np.ndarray | None = None
  qvel: np.ndarray | None = None
  ctrl: np.ndarray | None = None


@dataclass
class SOption:
  timestep: float = 0.002
  gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
  integrator: int = INT_EULER
  cone: int = CONE_PYRAMIDAL
  solver: int = SOL_NEWTON
  iterations: int = 100
  tolerance: float = 1e-8
  ls_iterations: int = 50
  ls_tolerance: float = 0.01
  impratio: float = 1.0
  disableflags: int = 0


class _OptView:
  """`model.opt` of a `ModelDesc`: MuJoCo's nested option struct as a live view of the flat `opt_*` attributes, so code written
  against `mujoco.MjModel` (`m.opt.timestep = 0.005`, tests/test_sim.py:62-72 of the reference) works unchanged."""

  _NAMES = ("timestep", "gravity", "integrator", "cone", "solver", "iterations", "tolerance", "ls_iterations", "ls_tolerance",
            "impratio", "disableflags", "ccd_iterations")

  def __init__(self, model):
    object.__setattr__(self, "_m", model)

  def __getattr__(self, name):
    if name in _OptView._NAMES:
      m = object.__getattribute__(self, "_m")
      if name == "ccd_iterations":
        return getattr(m, "opt_ccd_iterations", 50)
      return getattr(m, f"opt_{name}")
    raise AttributeError(name)

  def __setattr__(self, name, value):
    if name not in _OptView._NAMES:
      raise AttributeError(f"unknown option '{name}'")
    setattr(object.__getattribute__(self, "_m"), f"opt_{name}", np.asarray(value, float) if name == "gravity" else value)


class _Named:
  """`model.body(i)` / `model.geom("name")` style accessor result: id and name."""

  def __init__(self, kind, i, name):
    self.kind, self.id, self.name = kind, i, name


class ModelDesc:
  """Attribute bag of numpy arrays named after MjModel fields."""

  def __init__(self, **kw):
    self.__dict__.update(kw)

  # ---- the parts of mujoco.MjModel's object API that mjlab-side code touches ------------------------------------------
  @property
  def opt(self) -> _OptView:
    return _OptView(self)

  @property
  def stat(self):
    from types import SimpleNamespace
    return SimpleNamespace(meaninertia=getattr(self, "stat_meaninertia", 1.0))

  def _named(self, kind: str, names_attr: str, key):
    names = getattr(self, names_attr)
    i = names.index(key) if isinstance(key, str) else int(key)
    return _Named(kind, i, names[i])

  def body(self, key): return self._named("body", "body_names", key)          # noqa: E704
  def geom(self, key): return self._named("geom", "geom_names", key)          # noqa: E704
  def joint(self, key): return self._named("joint", "jnt_names", key)         # noqa: E704
  def jnt(self, key): return self._named("joint", "jnt_names", key)           # noqa: E704
  def site(self, key): return self._named("site", "site_names", key)          # noqa: E704
  def actuator(self, key): return self._named("actuator", "actuator_names", key)  # noqa: E704
  def sensor(self, key): return self._named("sensor", "sensor_names", key)    # noqa: E704

  # arrays copied verbatim from a mujoco.MjModel (same names, same layout)
  _MJ_INTS = ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensor", "nsensordata", "nmocap", "nkey")
  _MJ_ARRAYS = (
    "body_parentid", "body_rootid", "body_weldid", "body_mocapid", "body_jntnum", "body_jntadr", "body_dofnum",
    "body_dofadr", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_subtreemass",
    "body_invweight0", "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_pos", "jnt_axis", "jnt_range",
    "jnt_limited", "jnt_solref", "jnt_solimp", "jnt_margin", "jnt_stiffness", "dof_bodyid", "dof_jntid", "dof_parentid",
    "dof_armature", "dof_damping", "dof_frictionloss", "dof_solref", "dof_solimp", "dof_invweight0", "qpos0", "qpos_spring", "geom_type",
    "geom_bodyid", "geom_pos", "geom_quat", "geom_size", "geom_contype", "geom_conaffinity", "geom_condim",
    "geom_priority", "geom_friction", "geom_solref", "geom_solimp", "geom_solmix", "geom_margin", "geom_gap",
    "geom_rbound", "geom_group", "geom_rgba", "site_bodyid", "site_pos", "site_quat", "actuator_trnid",
    "actuator_gaintype", "actuator_biastype", "actuator_gainprm", "actuator_biasprm", "actuator_gear",
    "actuator_forcelimited", "actuator_forcerange", "actuator_ctrllimited", "actuator_ctrlrange", "sensor_type",
    "sensor_objtype", "sensor_objid", "sensor_reftype", "sensor_refid", "sensor_intprm", "sensor_dim", "sensor_adr",
    "sensor_cutoff", "key_qpos", "key_qvel", "key_ctrl",
  )

  @staticmethod
  def from_mjmodel(mj_model) -> "ModelDesc":
    """In-process conversion of a compiled `mujoco.MjModel` (or any object with the same attribute surface) into the array
    container the engine uploads - the `mjwarp.put_model(mjm)` half of sim/sim.py:139.  Arrays keep MuJoCo's names and layout;
    names come from the `model.body(i).name` accessors; `<exclude>` pairs from `exclude_signature`; the derived collision tables
    (candidate pairs, static-geom poses, geom field) are rebuilt by the same `collision_tables` the MJCF compiler uses."""
    m = ModelDesc()
    for k in ModelDesc._MJ_INTS:
      setattr(m, k, int(getattr(mj_model, k)))
    for k in ModelDesc._MJ_ARRAYS:
      if hasattr(mj_model, k):
        setattr(m, k, np.array(getattr(mj_model, k)))
    for kind, n in (("body", m.nbody), ("jnt", m.njnt), ("geom", m.ngeom), ("site", m.nsite), ("actuator", m.nu), ("sensor", m.nsensor)):
      acc = getattr(mj_model, {"jnt": "joint"}.get(kind, kind))
      setattr(m, f"{kind}_names", [acc(i).name for i in range(n)])
    o = mj_model.opt
    m.opt_timestep, m.opt_gravity = float(o.timestep), np.array(o.gravity, float)
    m.opt_integrator, m.opt_cone, m.opt_solver = int(o.integrator), int(o.cone), int(o.solver)
    m.opt_iterations, m.opt_ls_iterations = int(o.iterations), int(o.ls_iterations)
    m.opt_tolerance, m.opt_ls_tolerance, m.opt_impratio = float(o.tolerance), float(o.ls_tolerance), float(o.impratio)
    m.opt_disableflags = int(o.disableflags)
    m.opt_ccd_iterations = int(getattr(o, "ccd_iterations", 50))
    m.stat_meaninertia = float(mj_model.stat.meaninertia)
    # <exclude> body pairs: MuJoCo stores them as signatures (body1 << 16) + body2
    sig = np.asarray(getattr(mj_model, "exclude_signature", np.zeros(0, np.int64))).astype(np.int64).reshape(-1)
    excl = {(int(min(x >> 16, x & 0xFFFF)), int(max(x >> 16, x & 0xFFFF))) for x in sig}
    m.exclude_pairs = np.array(sorted(excl), np.int32).reshape(-1, 2)
    collision_tables(m, excl)
    return m.fill_defaults()

  def fill_defaults(self) -> "ModelDesc":
    """Arrays a source may lack (model files written before the field existed, duck-typed MjModels): MuJoCo's defaults."""
    nv = int(self.nv)
    if not hasattr(self, "dof_frictionloss"):
      self.dof_frictionloss = np.zeros(nv)
    if not hasattr(self, "dof_solref"):
      self.dof_solref = np.tile(np.array([0.02, 1.0]), (nv, 1))
    if not hasattr(self, "dof_solimp"):
      self.dof_solimp = np.tile(np.array([0.9, 0.95, 0.001, 0.5, 2.0]), (nv, 1))
    return self

  def fields(self):
    return {k: v for k, v in self.__dict__.items() if isinstance(v, np.ndarray)}

  def save(self, path) -> None:
    """Serialise to a single .npz (arrays as-is, everything else as JSON)."""
    import json

    arrays, meta = {}, {}
    for k, v in self.__dict__.items():
      if isinstance(v, np.ndarray):
        arrays[k] = v
      elif isinstance(v, (np.integer, np.floating)):
        meta[k] = v.item()
      else:
        meta[k] = v
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)

  @staticmethod
  def load(path) -> "ModelDesc":
    import json

    z = np.load(path)
    m = ModelDesc()
    for k in z.files:
      if k == "__meta__":
        m.__dict__.update(json.loads(bytes(z[k]).decode()))
      else:
        m.__dict__[k] = z[k]
    return m.fill_defaults()

  def name2id(self, kind: str, name: str) -> int:
    names = getattr(self, f"{kind}_names")
    try:
      return names.index(name)
    except ValueError as e:
      raise KeyError(f"{kind} '{name}' not found") from e


# ---------------------------------------------------------------------------
# Spec.
# ---------------------------------------------------------------------------
class Spec:
  def __init__(self):
    self.worldbody = SBody(name="world")
    self.actuators: list[SActuator] = []
    self.sensors: list[SSensor] = []
    self.excludes: list[tuple[str, str]] = []
    self.keys: list[SKey] = []
    self.option = SOption()
    self.autolimits = True
    self.modelname = ""

  # -- construction ---------------------------------------------------------
  @staticmethod
  def from_file(path: str | Path) -> "Spec":
    return Spec.from_string(Path(path).read_text())

  @staticmethod
  def from_string(xml: str) -> "Spec":
    return _Parser().parse(xml)

  # -- iteration helpers ----------------------------------------------------
  def bodies(self) -> list[SBody]:
    out: list[SBody] = []

    def rec(b):
      out.append(b)
      for c in b.children:
        rec(c)

    rec(self.worldbody)
    return out

  @property
  def geoms(self) -> list[SGeom]:
    return [g for b in self.bodies() for g in b.geoms]

  @property
  def joints(self) -> list[SJoint]:
    return [j for b in self.bodies() for j in b.joints]

  @property
  def sites(self) -> list[SSite]:
    return [s for b in self.bodies() for s in b.sites]

  def geom(self, name: str) -> SGeom:
    for g in self.geoms:
      if g.name == name:
        return g
    raise KeyError(name)

  def joint(self, name: str) -> SJoint:
    for j in self.joints:
      if j.name == name:
        return j
    raise KeyError(name)

  def body(self, name: str) -> SBody:
    for b in self.bodies():
      if b.name == name:
        return b
    raise KeyError(name)

  # -- editing (mirrors the MjSpec calls mjlab makes) -------------------------
  def add_actuator(self, **kw) -> SActuator:
    a = SActuator(**_np_kw(kw))
    self.actuators.append(a)
    return a

  def add_sensor(self, **kw) -> SSensor:
    s = SSensor(**kw)
    self.sensors.append(s)
    return s

  def add_key(self, name="", qpos=None, qvel=None, ctrl=None) -> SKey:
    k = SKey(name, None if qpos is None else np.array(qpos, float),
             None if qvel is None else np.array(qvel, float),
             None if ctrl is None else np.array(ctrl, float))
    self.keys.append(k)
    return k

  def attach(self, child: "Spec", prefix: str = "", parent: SBody | None = None) -> None:
    """Attach `child`'s world children under `parent` (default: worldbody),
    prefixing every name (scene/scene.py:176 `attach(ent.spec, prefix=...)`)."""
    parent = parent or self.worldbody
    child = copy.deepcopy(child)

    def rename(b: SBody):
      if b.name:
        b.name = prefix + b.name
      for lst in (b.joints, b.geoms, b.sites):
        for e in lst:
          if e.name:
            e.name = prefix + e.name
      for c in b.children:
        rename(c)

    for c in child.worldbody.children:
      rename(c)
      parent.children.append(c)
    # world-attached geoms/sites of the child go onto the parent body
    for g in child.worldbody.geoms:
      g.name = prefix + g.name if g.name else g.name
      parent.geoms.append(g)
    for s in child.worldbody.sites:
      s.name = prefix + s.name if s.name else s.name
      parent.sites.append(s)
    for a in child.actuators:
      a.name = prefix + a.name
      a.target = prefix + a.target
      self.actuators.append(a)
    for s in child.sensors:
      s.name = prefix + s.name
      if s.objname:
        s.objname = prefix + s.objname
      if s.refname:
        s.refname = prefix + s.refname
      self.sensors.append(s)
    for a, b in child.excludes:
      self.excludes.append((prefix + a, prefix + b))

  # -- compile ---------------------------------------------------------------
  def compile(self) -> ModelDesc:
    return _compile(self)


# ---------------------------------------------------------------------------
# XML parser with default classes.
# ---------------------------------------------------------------------------
class _Parser:
  def __init__(self):
    self.defaults: dict[str, dict[str, dict[str, str]]] = {"main": {}}
    self.parent_class: dict[str, str | None] = {"main": None}
    self.meshes: dict[str, str] = {}

  def parse(self, xml: str) -> Spec:
    root = ET.fromstring(xml)
    spec = Spec()
    spec.modelname = root.get("model", "")
    comp = root.find("compiler")
    if comp is not None:
      if comp.get("angle", "degree") != "radian":
        self.deg = True
      else:
        self.deg = False
      spec.autolimits = comp.get("autolimits", "true") == "true"
    else:
      self.deg = True
    for d in root.findall("default"):
      self._parse_defaults(d, "main", top=True)
    opt = root.find("option")
    if opt is not None:
      o = spec.option
      if opt.get("timestep"):
        o.timestep = float(opt.get("timestep"))
      if opt.get("gravity"):
        o.gravity = _floats(opt.get("gravity"))
      if opt.get("integrator"):
        o.integrator = {"Euler": INT_EULER, "RK4": INT_RK4, "implicit": INT_IMPLICIT,
                        "implicitfast": INT_IMPLICITFAST}[opt.get("integrator")]
      if opt.get("cone"):
        o.cone = {"pyramidal": CONE_PYRAMIDAL, "elliptic": CONE_ELLIPTIC}[opt.get("cone")]
      if opt.get("solver"):
        o.solver = {"PGS": SOL_PGS, "CG": SOL_CG, "Newton": SOL_NEWTON}[opt.get("solver")]
      for k in ("iterations", "ls_iterations"):
        if opt.get(k):
          setattr(o, k, int(opt.get(k)))
      for k in ("tolerance", "ls_tolerance", "impratio"):
        if opt.get(k):
          setattr(o, k, float(opt.get(k)))
    wb = root.find("worldbody")
    if wb is not None:
      self._parse_body_children(wb, spec.worldbody, None)
    con = root.find("contact")
    if con is not None:
      for e in con.findall("exclude"):
        spec.excludes.append((e.get("body1"), e.get("body2")))
    act = root.find("actuator")
    if act is not None:
      for e in act:
        spec.actuators.append(self._parse_actuator(e))
    sens = root.find("sensor")
    if sens is not None:
      for e in sens:
        spec.sensors.append(self._parse_sensor(e))
    kf = root.find("keyframe")
    if kf is not None:
      for e in kf.findall("key"):
        spec.add_key(e.get("name", ""), _floats(e.get("qpos")), _floats(e.get("qvel")), _floats(e.get("ctrl")))
    return spec

  # defaults ------------------------------------------------------------------
  def _parse_defaults(self, node, parent: str, top=False):
    cls = node.get("class", "main" if top else None)
    if cls is None:
      cls = "main"
    if cls not in self.defaults:
      self.defaults[cls] = {}
      self.parent_class[cls] = parent if cls != "main" else None
    for e in node:
      if e.tag == "default":
        self._parse_defaults(e, cls)
      else:
        self.defaults[cls].setdefault(e.tag, {}).update(e.attrib)

  def _resolved(self, tag: str, cls: str | None) -> dict[str, str]:
    chain = []
    c = cls or "main"
    while c is not None:
      chain.append(c)
      c = self.parent_class.get(c)
    out: dict[str, str] = {}
    for c in reversed(chain):
      out.update(self.defaults.get(c, {}).get(tag, {}))
    return out

  def _attrs(self, e, tag: str, childclass: str | None) -> dict[str, str]:
    cls = e.get("class", childclass)
    a = dict(self._resolved(tag, cls))
    own = dict(e.attrib)
    # an explicit fromto/pos on the element overrides a class-level one
    a.update(own)
    return a

  def _angle(self, v):
    return np.deg2rad(v) if self.deg else v

  def _orientation(self, a: dict[str, str]) -> np.ndarray:
    if "quat" in a:
      q = _floats(a["quat"])
      return q / np.linalg.norm(q)
    if "euler" in a:
      e = self._angle(_floats(a["euler"]))
      q = np.array([1.0, 0, 0, 0])
      for i, ang in enumerate(e):  # default eulerseq "xyz", intrinsic
        ax = np.zeros(3)
        ax[i] = 1
        q = quat_mul(q, np.concatenate([[math.cos(ang / 2)], ax * math.sin(ang / 2)]))
      return q
    if "axisangle" in a:
      v = _floats(a["axisangle"])
      ang = self._angle(v[3])
      ax = v[:3] / np.linalg.norm(v[:3])
      return np.concatenate([[math.cos(ang / 2)], ax * math.sin(ang / 2)])
    if "zaxis" in a:
      return quat_z2vec(_floats(a["zaxis"]))
    if "xyaxes" in a:
      v = _floats(a["xyaxes"])
      x = v[:3] / np.linalg.norm(v[:3])
      y = v[3:] - x * np.dot(x, v[3:])
      y /= np.linalg.norm(y)
      z = np.cross(x, y)
      return mat_to_quat(np.stack([x, y, z], axis=1))
    return np.array([1.0, 0, 0, 0])

  # bodies --------------------------------------------------------------------
  def _parse_body_children(self, node, body: SBody, childclass: str | None):
    for e in node:
      if e.tag == "body":
        cc = e.get("childclass", childclass)
        b = SBody(name=e.get("name", ""), pos=_floats(e.get("pos"), 3, [0, 0, 0]),
                  quat=self._orientation(e.attrib), mocap=e.get("mocap", "false") == "true")
        body.children.append(b)
        self._parse_body_children(e, b, cc)
      elif e.tag == "inertial":
        body.ipos = _floats(e.get("pos"), 3, [0, 0, 0])
        body.mass = float(e.get("mass"))
        if e.get("fullinertia"):
          f = _floats(e.get("fullinertia"))
          I = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
          w, v = np.linalg.eigh(I)
          order = np.argsort(-w)
          w, v = w[order], v[:, order]
          if np.linalg.det(v) < 0:
            v[:, 2] *= -1
          body.inertia = w
          body.iquat = quat_mul(self._orientation(e.attrib), mat_to_quat(v))
        else:
          body.inertia = _floats(e.get("diaginertia"), 3, [0, 0, 0])
          body.iquat = self._orientation(e.attrib)
      elif e.tag in ("joint", "freejoint"):
        a = self._attrs(e, "joint", childclass) if e.tag == "joint" else dict(e.attrib)
        j = SJoint(name=a.get("name", ""))
        j.type = JNT_FREE if e.tag == "freejoint" else JNT_TYPES[a.get("type", "hinge")]
        if "pos" in a:
          j.pos = _floats(a["pos"])
        if "axis" in a:
          ax = _floats(a["axis"])
          j.axis = ax / np.linalg.norm(ax)
        if "range" in a:
          r = _floats(a["range"])
          j.range = self._angle(r) if j.type == JNT_HINGE else r
        if "limited" in a:
          j.limited = {"false": 0, "true": 1, "auto": 2}[a["limited"]]
        for k in ("armature", "damping", "stiffness", "frictionloss", "margin"):
          if k in a:
            setattr(j, k, float(a[k]))
        for k in ("ref", "springref"):
          if k in a:
            v = float(a[k])
            setattr(j, k, self._angle(v) if j.type == JNT_HINGE else v)
        if "solreflimit" in a:
          j.solref = _floats(a["solreflimit"], 2, [0.02, 1.0])
        if "solimplimit" in a:
          j.solimp = _floats(a["solimplimit"], 5, [0.9, 0.95, 0.001, 0.5, 2.0])
        if "solreffriction" in a:
          j.solreffriction = _floats(a["solreffriction"], 2, [0.02, 1.0])
        if "solimpfriction" in a:
          j.solimpfriction = _floats(a["solimpfriction"], 5, [0.9, 0.95, 0.001, 0.5, 2.0])
        body.joints.append(j)
      elif e.tag == "geom":
        body.geoms.append(self._parse_geom(self._attrs(e, "geom", childclass)))
      elif e.tag == "site":
        a = self._attrs(e, "site", childclass)
        s = SSite(name=a.get("name", ""), pos=_floats(a.get("pos"), 3, [0, 0, 0]), quat=self._orientation(a))
        if "size" in a:
          s.size = _floats(a["size"], 3, [0.005, 0.005, 0.005])
        s.group = int(a.get("group", 0))
        body.sites.append(s)
      # light / camera / frame: irrelevant to the physics hot path

  def _parse_geom(self, a: dict[str, str]) -> SGeom:
    g = SGeom(name=a.get("name", ""))
    if "mesh" in a and "type" not in a:
      g.type = GEOM_MESH
    else:
      g.type = GEOM_TYPES[a.get("type", "sphere")]
    g.mesh = a.get("mesh", "")
    size = _floats(a.get("size"))
    if size is not None:
      g.size[: len(size)] = size
    g.pos = _floats(a.get("pos"), 3, [0, 0, 0])
    g.quat = self._orientation(a)
    if "fromto" in a and g.type in (GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_ELLIPSOID):
      ft = _floats(a["fromto"])
      vec = ft[0:3] - ft[3:6]  # z axis points from `to` towards `from`
      g.pos = 0.5 * (ft[0:3] + ft[3:6])
      g.quat = quat_z2vec(vec)
      half = 0.5 * np.linalg.norm(vec)
      if g.type in (GEOM_CAPSULE, GEOM_CYLINDER):
        g.size[1] = half
      else:
        g.size[2] = half
        g.size[1] = g.size[0]
    for k in ("contype", "conaffinity", "condim", "priority", "group"):
      if k in a:
        setattr(g, k, int(a[k]))
    if "friction" in a:
      f = _floats(a["friction"])
      g.friction[: len(f)] = f
    if "solref" in a:
      g.solref = _floats(a["solref"], 2, [0.02, 1.0])
    if "solimp" in a:
      g.solimp = _floats(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2.0])
    for k in ("solmix", "margin", "gap", "density"):
      if k in a:
        setattr(g, k, float(a[k]))
    if "mass" in a:
      g.mass = float(a["mass"])
    if "rgba" in a:
      g.rgba = _floats(a["rgba"], 4, [0.5, 0.5, 0.5, 1.0])
    return g

  def _parse_actuator(self, e) -> SActuator:
    a = dict(self._resolved(e.tag, e.get("class")))
    a.update(e.attrib)
    act = SActuator(name=a.get("name", ""), target=a.get("joint", ""))
    if "gear" in a:
      gv = _floats(a["gear"])
      act.gear[: len(gv)] = gv
    if "forcerange" in a:
      act.forcerange = _floats(a["forcerange"])
      act.forcelimited = True
    if "ctrlrange" in a:
      act.ctrlrange = _floats(a["ctrlrange"])
      act.ctrllimited = True
    if a.get("forcelimited") in ("true", "false"):
      act.forcelimited = a["forcelimited"] == "true"
    if a.get("ctrllimited") in ("true", "false"):
      act.ctrllimited = a["ctrllimited"] == "true"
    if e.tag == "motor":
      pass
    elif e.tag == "position":
      kp = float(a.get("kp", 1.0))
      kv = float(a.get("kv", 0.0))
      act.gainprm[0] = kp
      act.biastype = BIAS_AFFINE
      act.biasprm[1] = -kp
      act.biasprm[2] = -kv
    elif e.tag == "velocity":
      kv = float(a.get("kv", 1.0))
      act.gainprm[0] = kv
      act.biastype = BIAS_AFFINE
      act.biasprm[2] = -kv
    elif e.tag == "general":
      if "gainprm" in a:
        v = _floats(a["gainprm"])
        act.gainprm[:] = 0
        act.gainprm[: len(v)] = v
      if "biasprm" in a:
        v = _floats(a["biasprm"])
        act.biasprm[: len(v)] = v
      act.biastype = {"none": BIAS_NONE, "affine": BIAS_AFFINE}[a.get("biastype", "none")]
    else:
      raise ValueError(f"unsupported actuator <{e.tag}>")
    return act

  def _parse_sensor(self, e) -> SSensor:
    if e.tag not in SENSOR_TAGS:
      raise ValueError(f"unsupported sensor <{e.tag}>")
    stype, dim, objtype, attr = SENSOR_TAGS[e.tag]
    s = SSensor(name=e.get("name", ""), type=stype, dim=dim, cutoff=float(e.get("cutoff", 0.0)))
    if attr is not None:
      s.objtype, s.objname = objtype, e.get(attr)
    else:
      s.objtype = {"body": OBJ_BODY, "xbody": OBJ_XBODY, "geom": OBJ_GEOM, "site": OBJ_SITE}[e.get("objtype")]
      s.objname = e.get("objname")
      if e.get("reftype"):
        s.reftype = {"body": OBJ_BODY, "xbody": OBJ_XBODY, "geom": OBJ_GEOM, "site": OBJ_SITE}[e.get("reftype")]
        s.refname = e.get("refname")
    return s


# ---------------------------------------------------------------------------
# Inertia of primitive geoms (for bodies without <inertial>).
# ---------------------------------------------------------------------------
def _geom_volume_inertia(g: SGeom):
  """Returns (volume, unit-density diagonal inertia in the geom frame)."""
  s = g.size
  if g.type == GEOM_SPHERE:
    r = s[0]
    v = 4.0 / 3.0 * math.pi * r**3
    i = 0.4 * v * r * r
    return v, np.array([i, i, i])
  if g.type == GEOM_BOX:
    v = 8 * s[0] * s[1] * s[2]
    return v, v / 3.0 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
  if g.type == GEOM_CYLINDER:
    r, h = s[0], s[1]
    v = math.pi * r * r * 2 * h
    ix = v * (3 * r * r + 4 * h * h) / 12.0
    return v, np.array([ix, ix, v * r * r / 2])
  if g.type == GEOM_CAPSULE:
    r, h = s[0], s[1]
    vc = math.pi * r * r * 2 * h
    vs = 4.0 / 3.0 * math.pi * r**3
    v = vc + vs
    iz = vc * r * r / 2 + vs * 0.4 * r * r
    ix = vc * (3 * r * r + 4 * h * h) / 12.0 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
    return v, np.array([ix, ix, iz])
  if g.type == GEOM_ELLIPSOID:
    v = 4.0 / 3.0 * math.pi * s[0] * s[1] * s[2]
    return v, v / 5.0 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
  return 0.0, np.zeros(3)


def _infer_inertial(b: SBody):
  tot = 0.0
  com = np.zeros(3)
  parts = []
  for g in b.geoms:
    if g.type in (GEOM_PLANE, GEOM_MESH, GEOM_HFIELD):
      continue
    v, iu = _geom_volume_inertia(g)
    if v <= 0:
      continue
    m = g.mass if g.mass is not None else g.density * v
    if m <= 0:
      continue
    parts.append((m, iu * (m / v), g))
    tot += m
    com += m * g.pos
  if tot <= 0:
    return np.zeros(3), np.array([1.0, 0, 0, 0]), 0.0, np.zeros(3)
  com /= tot
  I = np.zeros((3, 3))
  for m, idiag, g in parts:
    R = quat_to_mat(g.quat)
    d = g.pos - com
    I += R @ np.diag(idiag) @ R.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
  w, v = np.linalg.eigh(I)
  order = np.argsort(-w)
  w, v = w[order], v[:, order]
  if np.linalg.det(v) < 0:
    v[:, 2] *= -1
  # keep axis-aligned frames exactly aligned (common case: single primitive)
  if np.allclose(I, np.diag(np.diag(I)), atol=1e-14 * max(1.0, np.abs(I).max())):
    return com, np.array([1.0, 0, 0, 0]), tot, np.diag(I).copy()
  return com, mat_to_quat(v), tot, w


# ---------------------------------------------------------------------------
# Compile: Spec -> ModelDesc.
# ---------------------------------------------------------------------------
def collision_tables(m: "ModelDesc", excl: set) -> None:
  """Static-geom poses, the geom field and the candidate pair list of a model whose body / geom arrays are filled in
  (`excl`: set of excluded (body, body) pairs, lower id first).  Shared by `_compile` and `tools/export_mjmodel.py`."""
  nbody, ngeom = m.nbody, m.ngeom
  body_weldid, body_mocapid, body_parentid = np.asarray(m.body_weldid), np.asarray(m.body_mocapid), np.asarray(m.body_parentid)
  body_pos, body_quat = np.asarray(m.body_pos, float), np.asarray(m.body_quat, float)
  # static geometry: geoms of bodies welded to the world never move, so their world pose is a model constant
  # (`sgeom_xpos/xmat`) and the kernels only run forward kinematics for `dyn_geom`
  gb, gw = m.geom_bodyid, body_weldid
  sxpos, sxquat = np.zeros((nbody, 3)), np.tile([1.0, 0, 0, 0], (nbody, 1))
  for b in range(1, nbody):
    if gw[b] == 0 and body_mocapid[b] < 0:
      pb = body_parentid[b]
      sxpos[b] = sxpos[pb] + quat_to_mat(sxquat[pb]) @ body_pos[b]
      sxquat[b] = quat_mul(sxquat[pb], body_quat[b])
  static = np.array([gw[gb[g]] == 0 and body_mocapid[gb[g]] < 0 for g in range(ngeom)], bool) if ngeom else np.zeros(0, bool)
  m.dyn_geom = np.flatnonzero(~static).astype(np.int32)
  m.ndyn = len(m.dyn_geom)
  m.sgeom_xpos, m.sgeom_xmat = np.zeros((ngeom, 3)), np.zeros((ngeom, 9))
  for g in np.flatnonzero(static):
    Rb = quat_to_mat(sxquat[gb[g]])
    m.sgeom_xpos[g] = sxpos[gb[g]] + Rb @ m.geom_pos[g]
    m.sgeom_xmat[g] = quat_to_mat(quat_mul(sxquat[gb[g]], m.geom_quat[g])).reshape(9)

  # geom field: many static, bounded, collidable geoms (terrain boxes) are not enumerated pair by pair; they go into a
  # uniform xy grid (cell -> field geoms whose AABB overlaps it) that each moving "query" geom looks up at run time
  collidable = (m.geom_contype != 0) | (m.geom_conaffinity != 0)
  bounded = np.isin(m.geom_type, (GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX))
  fld = np.flatnonzero(static & collidable & bounded)
  if len(fld) < FIELD_MIN_GEOMS:
    fld = fld[:0]
  in_field = np.zeros(ngeom, bool)
  in_field[fld] = True
  _build_geom_field(m, fld, body_parentid, excl)

  pairs = []
  active = [g for g in range(ngeom) if collidable[g] and not in_field[g]]
  for ia, g1 in enumerate(active):
    for g2 in active[ia + 1:]:
      b1, b2 = gb[g1], gb[g2]
      ct1, ca1, ct2, ca2 = m.geom_contype[g1], m.geom_conaffinity[g1], m.geom_contype[g2], m.geom_conaffinity[g2]
      if not ((ct1 & ca2) or (ct2 & ca1)):
        continue
      w1, w2 = gw[b1], gw[b2]
      if w1 == w2:
        continue
      wp1, wp2 = gw[body_parentid[w1]], gw[body_parentid[w2]]
      if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
        continue
      if (min(b1, b2), max(b1, b2)) in excl:
        continue
      t1, t2 = m.geom_type[g1], m.geom_type[g2]
      if t1 in (GEOM_MESH, GEOM_HFIELD) or t2 in (GEOM_MESH, GEOM_HFIELD):
        continue
      if t1 == GEOM_PLANE and t2 == GEOM_PLANE:
        continue
      # contact.geom1 carries the lower geom-type enum (MuJoCo convention)
      pairs.append((g1, g2) if t1 <= t2 else (g2, g1))
  m.pair_geom = np.array(pairs, np.int32).reshape(-1, 2)
  m.npair = len(pairs)


FIELD_MIN_GEOMS = 16     # fewer static geoms than this stay ordinary candidate pairs
FIELD_CELL = 1.0         # grid cell edge [m]


def _build_geom_field(m: "ModelDesc", fld: np.ndarray, body_parentid, excl) -> None:
  """Uniform-grid index over the static field geoms `fld` (see `_compile`).

  Writes: field_geom[nfield] (geom id), field_aabb[nfield, 6] (centre, half extents of the world AABB),
  field_cell0[nfield, 2] (lowest grid cell the AABB touches: a pair met in several cells is reported from the first
  common one), field_grid = [x0, y0, cell, 1 / cell], field_nx / field_ny, field_cell_adr[ncell + 1], field_cell_item[...]
  (field indices, ascending within a cell), fq_geom[nfq] (moving geoms that can touch the field, in geom order).
  """
  nf = len(fld)
  m.nfield = nf
  m.field_geom = fld.astype(np.int32)
  m.field_aabb = np.zeros((nf, 6))
  m.field_cell0 = np.zeros((nf, 2), np.int32)
  m.field_grid = np.array([0.0, 0.0, FIELD_CELL, 1.0 / FIELD_CELL])
  m.field_nx = m.field_ny = 0
  m.field_cell_adr = np.zeros(1, np.int32)
  m.field_cell_item = np.zeros(0, np.int32)
  m.fq_geom = np.zeros(0, np.int32)
  m.nfq = 0
  if nf == 0:
    return
  for i, g in enumerate(fld):
    R = m.sgeom_xmat[g].reshape(3, 3)
    t, sz = m.geom_type[g], m.geom_size[g]
    if t == GEOM_BOX:
      half = np.abs(R) @ sz
    elif t == GEOM_SPHERE:
      half = np.full(3, sz[0])
    else:
      half = np.abs(R[:, 2]) * sz[1] + sz[0]
    m.field_aabb[i] = np.concatenate([m.sgeom_xpos[g], half])
  lo = (m.field_aabb[:, :2] - m.field_aabb[:, 3:5]).min(0)
  hi = (m.field_aabb[:, :2] + m.field_aabb[:, 3:5]).max(0)
  cell = FIELD_CELL
  nx, ny = (int(math.ceil((hi[k] - lo[k]) / cell)) + 1 for k in range(2))
  m.field_grid = np.array([lo[0], lo[1], cell, 1.0 / cell])
  m.field_nx, m.field_ny = nx, ny
  c0 = np.floor((m.field_aabb[:, :2] - m.field_aabb[:, 3:5] - lo) / cell).astype(np.int64)
  c1 = np.floor((m.field_aabb[:, :2] + m.field_aabb[:, 3:5] - lo) / cell).astype(np.int64)
  c0[:, 0], c1[:, 0] = np.clip(c0[:, 0], 0, nx - 1), np.clip(c1[:, 0], 0, nx - 1)
  c0[:, 1], c1[:, 1] = np.clip(c0[:, 1], 0, ny - 1), np.clip(c1[:, 1], 0, ny - 1)
  m.field_cell0 = c0.astype(np.int32)
  counts = np.zeros(nx * ny + 1, np.int64)
  for i in range(nf):
    xs, ys = np.arange(c0[i, 0], c1[i, 0] + 1), np.arange(c0[i, 1], c1[i, 1] + 1)
    np.add.at(counts, (xs[:, None] * ny + ys[None, :]).ravel() + 1, 1)
  adr = np.cumsum(counts)
  items = np.zeros(adr[-1], np.int32)
  fill = adr[:-1].copy()
  for i in range(nf):     # ascending i => ascending field index inside every cell
    xs, ys = np.arange(c0[i, 0], c1[i, 0] + 1), np.arange(c0[i, 1], c1[i, 1] + 1)
    cells = (xs[:, None] * ny + ys[None, :]).ravel()
    items[fill[cells]] = i
    fill[cells] += 1
  m.field_cell_adr, m.field_cell_item = adr.astype(np.int32), items
  # query geoms: the collision filter of every (moving geom, field geom) pair, which must not depend on the field geom
  ct, ca = m.geom_contype, m.geom_conaffinity
  fct, fca = np.unique(ct[fld]), np.unique(ca[fld])
  fbody = np.unique(m.geom_bodyid[fld])
  if len(fct) != 1 or len(fca) != 1:
    raise NotImplementedError("field geoms with differing contype / conaffinity")
  q = []
  for g in m.dyn_geom:
    if not ((ct[g] & fca[0]) or (fct[0] & ca[g])):
      continue
    b = m.geom_bodyid[g]
    if any((min(b, fb), max(b, fb)) in excl for fb in fbody):
      raise NotImplementedError("<exclude> between a moving body and a field body")
    if m.geom_type[g] not in (GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX):
      raise NotImplementedError(f"geom '{m.geom_names[g]}': only spheres, capsules and boxes collide with a geom field "
                                "(mesh / ellipsoid / cylinder narrowphase is not implemented)")
    if m.geom_type[g] > GEOM_BOX or np.any(m.geom_type[fld] != GEOM_BOX):
      raise NotImplementedError("geom fields hold boxes only")
    q.append(g)
  m.fq_geom = np.array(q, np.int32)
  m.nfq = len(q)


def _compile(spec: Spec) -> ModelDesc:
  bodies: list[SBody] = []
  parent: list[int] = []

  def rec(b: SBody, p: int):
    idx = len(bodies)
    bodies.append(b)
    parent.append(p)
    for c in b.children:
      rec(c, idx)

  # MuJoCo orders bodies depth-first.
  rec(spec.worldbody, 0)
  nbody = len(bodies)
  body_names = [b.name for b in bodies]

  body_pos = np.array([b.pos for b in bodies], float).reshape(nbody, 3)
  body_quat = np.array([b.quat / np.linalg.norm(b.quat) for b in bodies], float).reshape(nbody, 4)
  body_ipos = np.zeros((nbody, 3))
  body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
  body_mass = np.zeros(nbody)
  body_inertia = np.zeros((nbody, 3))
  for i, b in enumerate(bodies):
    if i == 0:
      continue
    if b.mass is not None:
      body_ipos[i] = b.ipos
      body_iquat[i] = b.iquat / np.linalg.norm(b.iquat)
      body_mass[i] = b.mass
      body_inertia[i] = b.inertia
    else:
      body_ipos[i], body_iquat[i], body_mass[i], body_inertia[i] = _infer_inertial(b)

  # joints / dofs
  jnt_type, jnt_qposadr, jnt_dofadr, jnt_bodyid = [], [], [], []
  jnt_pos, jnt_axis, jnt_range, jnt_limited = [], [], [], []
  jnt_solref, jnt_solimp, jnt_margin, jnt_stiffness = [], [], [], []
  jnt_names: list[str] = []
  dof_bodyid, dof_jntid, dof_parentid = [], [], []
  dof_armature, dof_damping, dof_frictionloss, dof_solref, dof_solimp = [], [], [], [], []
  qpos0: list[float] = []
  qpos_spring: list[float] = []
  body_jntnum = np.zeros(nbody, np.int32)
  body_jntadr = -np.ones(nbody, np.int32)
  body_dofnum = np.zeros(nbody, np.int32)
  body_dofadr = -np.ones(nbody, np.int32)
  body_lastdof = -np.ones(nbody, np.int64)  # last dof on the path to root

  for i, b in enumerate(bodies):
    last = body_lastdof[parent[i]] if i > 0 else -1
    if b.joints:
      body_jntadr[i] = len(jnt_type)
      body_dofadr[i] = len(dof_bodyid)
    for j in b.joints:
      jid = len(jnt_type)
      jnt_names.append(j.name)
      jnt_type.append(j.type)
      jnt_qposadr.append(len(qpos0))
      jnt_dofadr.append(len(dof_bodyid))
      jnt_bodyid.append(i)
      jnt_pos.append(j.pos)
      jnt_axis.append(j.axis)
      lim = j.limited
      if lim == 2:
        lim = 1 if (spec.autolimits and j.range[0] < j.range[1]) else 0
      if j.type == JNT_FREE:
        lim = 0
      jnt_limited.append(lim)
      jnt_range.append(j.range)
      jnt_solref.append(j.solref)
      jnt_solimp.append(j.solimp)
      jnt_margin.append(j.margin)
      jnt_stiffness.append(j.stiffness)
      ndof = {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}[j.type]
      if j.type == JNT_FREE:
        q = body_quat[i]
        qpos0.extend([*body_pos[i], *q])
        qpos_spring.extend([*body_pos[i], *q])
      elif j.type == JNT_BALL:
        qpos0.extend([1, 0, 0, 0])
        qpos_spring.extend([1, 0, 0, 0])
      else:
        qpos0.append(j.ref)
        qpos_spring.append(j.springref)
      for _ in range(ndof):
        dof_bodyid.append(i)
        dof_jntid.append(jid)
        dof_parentid.append(last)
        last = len(dof_bodyid) - 1
        dof_armature.append(j.armature)
        dof_damping.append(j.damping)
        dof_frictionloss.append(j.frictionloss)
        dof_solref.append(j.solreffriction)
        dof_solimp.append(j.solimpfriction)
      body_jntnum[i] += 1
      body_dofnum[i] += ndof
    body_lastdof[i] = last

  njnt, nv, nq = len(jnt_type), len(dof_bodyid), len(qpos0)

  body_parentid = np.array(parent, np.int32)
  body_rootid = np.zeros(nbody, np.int32)
  body_weldid = np.zeros(nbody, np.int32)
  for i in range(1, nbody):
    body_rootid[i] = i if parent[i] == 0 else body_rootid[parent[i]]
    body_weldid[i] = i if body_jntnum[i] > 0 else body_weldid[parent[i]]
  body_mocapid = -np.ones(nbody, np.int32)
  nmocap = 0
  for i, b in enumerate(bodies):
    if b.mocap:
      body_mocapid[i] = nmocap
      nmocap += 1
  body_subtreemass = body_mass.copy()
  for i in range(nbody - 1, 0, -1):
    body_subtreemass[parent[i]] += body_subtreemass[i]

  # geoms / sites
  geoms, geom_bodyid = [], []
  sites, site_bodyid = [], []
  for i, b in enumerate(bodies):
    for g in b.geoms:
      geoms.append(g)
      geom_bodyid.append(i)
    for s in b.sites:
      sites.append(s)
      site_bodyid.append(i)
  ngeom, nsite = len(geoms), len(sites)

  def rbound(g: SGeom) -> float:
    s = g.size
    if g.type == GEOM_SPHERE:
      return s[0]
    if g.type == GEOM_CAPSULE:
      return s[0] + s[1]
    if g.type == GEOM_CYLINDER:
      return math.hypot(s[0], s[1])
    if g.type == GEOM_ELLIPSOID:
      return float(max(s))
    if g.type == GEOM_BOX:
      return float(np.linalg.norm(s))
    return 0.0

  def arr(lst, shape, dtype=float):
    return np.array(lst, dtype).reshape(shape)

  m = ModelDesc()
  m.nq, m.nv, m.nbody, m.njnt, m.ngeom, m.nsite, m.nmocap = nq, nv, nbody, njnt, ngeom, nsite, nmocap
  m.body_names, m.jnt_names = body_names, jnt_names
  m.geom_names = [g.name for g in geoms]
  m.site_names = [s.name for s in sites]
  m.body_parentid, m.body_rootid, m.body_weldid, m.body_mocapid = body_parentid, body_rootid, body_weldid, body_mocapid
  m.body_jntnum, m.body_jntadr, m.body_dofnum, m.body_dofadr = body_jntnum, body_jntadr, body_dofnum, body_dofadr
  m.body_pos, m.body_quat, m.body_ipos, m.body_iquat = body_pos, body_quat, body_ipos, body_iquat
  m.body_mass, m.body_inertia, m.body_subtreemass = body_mass, body_inertia, body_subtreemass
  m.jnt_type = arr(jnt_type, (njnt,), np.int32)
  m.jnt_qposadr = arr(jnt_qposadr, (njnt,), np.int32)
  m.jnt_dofadr = arr(jnt_dofadr, (njnt,), np.int32)
  m.jnt_bodyid = arr(jnt_bodyid, (njnt,), np.int32)
  m.jnt_pos = arr(jnt_pos, (njnt, 3))
  m.jnt_axis = arr(jnt_axis, (njnt, 3))
  m.jnt_range = arr(jnt_range, (njnt, 2))
  m.jnt_limited = arr(jnt_limited, (njnt,), np.int32)
  m.jnt_solref = arr(jnt_solref, (njnt, 2))
  m.jnt_solimp = arr(jnt_solimp, (njnt, 5))
  m.jnt_margin = arr(jnt_margin, (njnt,))
  m.jnt_stiffness = arr(jnt_stiffness, (njnt,))
  m.dof_bodyid = arr(dof_bodyid, (nv,), np.int32)
  m.dof_jntid = arr(dof_jntid, (nv,), np.int32)
  m.dof_parentid = arr(dof_parentid, (nv,), np.int32)
  m.dof_armature = arr(dof_armature, (nv,))
  m.dof_damping = arr(dof_damping, (nv,))
  m.dof_frictionloss = arr(dof_frictionloss, (nv,))
  m.dof_solref = arr(dof_solref, (nv, 2))
  m.dof_solimp = arr(dof_solimp, (nv, 5))
  m.qpos0 = arr(qpos0, (nq,))
  m.qpos_spring = arr(qpos_spring, (nq,))
  m.geom_type = arr([g.type for g in geoms], (ngeom,), np.int32)
  m.geom_bodyid = arr(geom_bodyid, (ngeom,), np.int32)
  m.geom_pos = arr([g.pos for g in geoms], (ngeom, 3))
  m.geom_quat = arr([g.quat / np.linalg.norm(g.quat) for g in geoms], (ngeom, 4))
  m.geom_size = arr([g.size for g in geoms], (ngeom, 3))
  m.geom_contype = arr([g.contype for g in geoms], (ngeom,), np.int32)
  m.geom_conaffinity = arr([g.conaffinity for g in geoms], (ngeom,), np.int32)
  m.geom_condim = arr([g.condim for g in geoms], (ngeom,), np.int32)
  m.geom_priority = arr([g.priority for g in geoms], (ngeom,), np.int32)
  m.geom_friction = arr([g.friction for g in geoms], (ngeom, 3))
  m.geom_solref = arr([g.solref for g in geoms], (ngeom, 2))
  m.geom_solimp = arr([g.solimp for g in geoms], (ngeom, 5))
  m.geom_solmix = arr([g.solmix for g in geoms], (ngeom,))
  m.geom_margin = arr([g.margin for g in geoms], (ngeom,))
  m.geom_gap = arr([g.gap for g in geoms], (ngeom,))
  m.geom_rbound = arr([rbound(g) for g in geoms], (ngeom,))
  m.geom_group = arr([g.group for g in geoms], (ngeom,), np.int32)
  m.geom_rgba = arr([g.rgba for g in geoms], (ngeom, 4))
  m.site_bodyid = arr(site_bodyid, (nsite,), np.int32)
  m.site_pos = arr([s.pos for s in sites], (nsite, 3))
  m.site_quat = arr([s.quat / np.linalg.norm(s.quat) for s in sites], (nsite, 4))

  # actuators
  nu = len(spec.actuators)
  m.nu = nu
  m.actuator_names = [a.name for a in spec.actuators]
  trnid = []
  for a in spec.actuators:
    if a.target not in jnt_names:
      raise ValueError(f"actuator '{a.name}' targets unknown joint '{a.target}'")
    trnid.append([jnt_names.index(a.target), -1])
  m.actuator_trnid = arr(trnid, (nu, 2), np.int32)
  m.actuator_gaintype = arr([a.gaintype for a in spec.actuators], (nu,), np.int32)
  m.actuator_biastype = arr([a.biastype for a in spec.actuators], (nu,), np.int32)
  m.actuator_gainprm = arr([a.gainprm for a in spec.actuators], (nu, 10))
  m.actuator_biasprm = arr([a.biasprm for a in spec.actuators], (nu, 10))
  m.actuator_gear = arr([a.gear for a in spec.actuators], (nu, 6))
  m.actuator_forcelimited = arr([int(a.forcelimited) for a in spec.actuators], (nu,), np.int32)
  m.actuator_forcerange = arr([a.forcerange for a in spec.actuators], (nu, 2))
  m.actuator_ctrllimited = arr([int(a.ctrllimited) for a in spec.actuators], (nu,), np.int32)
  m.actuator_ctrlrange = arr([a.ctrlrange for a in spec.actuators], (nu, 2))

  # sensors
  def resolve(objtype: int, name: str) -> int:
    if objtype in (OBJ_BODY, OBJ_XBODY):
      return body_names.index(name)
    if objtype == OBJ_GEOM:
      return m.geom_names.index(name)
    if objtype == OBJ_SITE:
      return m.site_names.index(name)
    if objtype == OBJ_JOINT:
      return jnt_names.index(name)
    return -1

  ns = len(spec.sensors)
  m.nsensor = ns
  m.sensor_names = [s.name for s in spec.sensors]
  s_type, s_objtype, s_objid, s_reftype, s_refid, s_intprm, s_dim, s_adr, s_cutoff = ([] for _ in range(9))
  adr = 0
  for s in spec.sensors:
    dim = s.dim
    if s.type == SENS_CONTACT:
      bits, _reduce, num = s.intprm
      dim = num * sum(d for k, d in enumerate(CONTACT_FIELD_DIMS) if bits & (1 << k))
    s_type.append(s.type)
    s_objtype.append(s.objtype)
    s_objid.append(resolve(s.objtype, s.objname) if s.objname else -1)
    s_reftype.append(s.reftype)
    s_refid.append(resolve(s.reftype, s.refname) if s.refname else -1)
    s_intprm.append(list(s.intprm))
    s_dim.append(dim)
    s_adr.append(adr)
    s_cutoff.append(s.cutoff)
    adr += dim
  m.nsensordata = adr
  m.sensor_type = arr(s_type, (ns,), np.int32)
  m.sensor_objtype = arr(s_objtype, (ns,), np.int32)
  m.sensor_objid = arr(s_objid, (ns,), np.int32)
  m.sensor_reftype = arr(s_reftype, (ns,), np.int32)
  m.sensor_refid = arr(s_refid, (ns,), np.int32)
  m.sensor_intprm = arr(s_intprm, (ns, 3), np.int32)
  m.sensor_dim = arr(s_dim, (ns,), np.int32)
  m.sensor_adr = arr(s_adr, (ns,), np.int32)
  m.sensor_cutoff = arr(s_cutoff, (ns,))

  # keyframes
  m.nkey = len(spec.keys)
  m.key_names = [k.name for k in spec.keys]
  m.key_qpos = arr([k.qpos if k.qpos is not None else m.qpos0 for k in spec.keys], (m.nkey, nq))
  m.key_qvel = arr([k.qvel if k.qvel is not None else np.zeros(nv) for k in spec.keys], (m.nkey, nv))
  m.key_ctrl = arr([k.ctrl if k.ctrl is not None else np.zeros(nu) for k in spec.keys], (m.nkey, nu))

  # options
  o = spec.option
  m.opt_timestep = float(o.timestep)
  m.opt_gravity = np.array(o.gravity, float)
  m.opt_integrator, m.opt_cone, m.opt_solver = o.integrator, o.cone, o.solver
  m.opt_iterations, m.opt_ls_iterations = o.iterations, o.ls_iterations
  m.opt_tolerance, m.opt_ls_tolerance, m.opt_impratio = o.tolerance, o.ls_tolerance, o.impratio
  m.opt_disableflags = o.disableflags

  # collision candidate pairs (static filter; MuJoCo's geom-pair filter)
  excl = set()
  for a, b in spec.excludes:
    ia, ib = body_names.index(a), body_names.index(b)
    excl.add((min(ia, ib), max(ia, ib)))
  m.exclude_pairs = np.array(sorted(excl), np.int32).reshape(-1, 2)
  collision_tables(m, excl)
  _setconst(m)
  return m


def _setconst(m: ModelDesc) -> None:
  """qpos0-dependent constants: body_invweight0, dof_invweight0, stat_meaninertia.

  Restates what MuJoCo's compiler derives at `qpos0` (the quantities the
  constraint rows use as their diagonal inverse-inertia approximation).  M is
  assembled here as sum_b J_b^T I_b J_b from world-frame body Jacobians, on
  purpose a different algorithm from the CRB pass of the device kernels and
  the oracle, so the three cross-check each other in tests.
  """
  nb, nv = m.nbody, m.nv
  xpos = np.zeros((nb, 3))
  xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
  xanchor = np.zeros((m.njnt, 3))
  xaxis = np.zeros((m.njnt, 3))
  for i in range(1, nb):
    p = m.body_parentid[i]
    pos = xpos[p] + quat_to_mat(xquat[p]) @ m.body_pos[i]
    quat = quat_mul(xquat[p], m.body_quat[i])
    for k in range(m.body_jntnum[i]):
      j = m.body_jntadr[i] + k
      if m.jnt_type[j] == JNT_FREE:
        qa = m.jnt_qposadr[j]
        pos = m.qpos0[qa:qa + 3].copy()
        quat = m.qpos0[qa + 3:qa + 7] / np.linalg.norm(m.qpos0[qa + 3:qa + 7])
        xanchor[j] = pos
        xaxis[j] = quat_to_mat(quat)[:, 2]
      else:
        R = quat_to_mat(quat)
        xanchor[j] = pos + R @ m.jnt_pos[j]
        xaxis[j] = R @ m.jnt_axis[j]
        # at qpos0 the joint displacement (qpos - qpos0) is zero
    xpos[i], xquat[i] = pos, quat / np.linalg.norm(quat)
  xipos = np.zeros((nb, 3))
  Iw = np.zeros((nb, 3, 3))
  for i in range(nb):
    R = quat_to_mat(xquat[i])
    xipos[i] = xpos[i] + R @ m.body_ipos[i]
    Ri = quat_to_mat(quat_mul(xquat[i], m.body_iquat[i]))
    Iw[i] = Ri @ np.diag(m.body_inertia[i]) @ Ri.T

  def jac(body: int, point: np.ndarray):
    jp = np.zeros((3, nv))
    jr = np.zeros((3, nv))
    b = body
    while b > 0:
      for k in range(m.body_jntnum[b]):
        j = m.body_jntadr[b] + k
        d = m.jnt_dofadr[j]
        t = m.jnt_type[j]
        if t == JNT_FREE:
          jp[:, d:d + 3] = np.eye(3)
          R = quat_to_mat(xquat[b])
          for a in range(3):
            jr[:, d + 3 + a] = R[:, a]
            jp[:, d + 3 + a] = np.cross(R[:, a], point - xpos[b])
        elif t == JNT_HINGE:
          jr[:, d] = xaxis[j]
          jp[:, d] = np.cross(xaxis[j], point - xanchor[j])
        elif t == JNT_SLIDE:
          jp[:, d] = xaxis[j]
        else:
          raise NotImplementedError("ball joints are outside the supported subset")
      b = m.body_parentid[b]
    return jp, jr

  M = np.diag(m.dof_armature.astype(float)) if nv else np.zeros((0, 0))
  jacs = []
  for i in range(nb):
    jp, jr = jac(i, xipos[i])
    jacs.append((jp, jr))
    if m.body_mass[i] > 0 or np.any(m.body_inertia[i] > 0):
      M = M + m.body_mass[i] * jp.T @ jp + jr.T @ Iw[i] @ jr
  m.body_invweight0 = np.zeros((nb, 2))
  m.dof_invweight0 = np.zeros(nv)
  m.stat_meaninertia = 1.0
  m.qM0 = M
  if nv == 0:
    return
  Minv = np.linalg.inv(M)
  m.stat_meaninertia = float(np.mean(np.diag(M)))
  for i in range(1, nb):
    if m.body_weldid[i] == 0:
      continue
    jp, jr = jacs[i]
    m.body_invweight0[i, 0] = np.trace(jp @ Minv @ jp.T) / 3.0
    m.body_invweight0[i, 1] = np.trace(jr @ Minv @ jr.T) / 3.0
  dg = np.diag(Minv)
  for j in range(m.njnt):
    d = m.jnt_dofadr[j]
    if m.jnt_type[j] == JNT_FREE:
      m.dof_invweight0[d:d + 3] = dg[d:d + 3].mean()
      m.dof_invweight0[d + 3:d + 6] = dg[d + 3:d + 6].mean()
    elif m.jnt_type[j] == JNT_BALL:
      m.dof_invweight0[d:d + 3] = dg[d:d + 3].mean()
    else:
      m.dof_invweight0[d] = dg[d]
