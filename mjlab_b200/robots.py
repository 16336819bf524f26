"""Scene assembly for the BASELINE configs (G1 / Go1 flat velocity).

Restates, on top of `mjcf.Spec`, the spec edits the reference performs before
`scene.compile()`:

  * robot constants (actuator groups, collision config, init keyframe):
    src/mjlab/asset_zoo/robots/unitree_g1/g1_constants.py:44-262,
    src/mjlab/asset_zoo/robots/unitree_go1/go1_constants.py:40-150
  * `create_position_actuator` src/mjlab/utils/spec.py:122-165
  * `CollisionCfg.edit_spec` src/mjlab/utils/spec_config.py:206-237
  * keyframe `Entity._add_initial_state_keyframe` src/mjlab/entity/entity.py:170-207
  * scene composition order terrain -> entities -> sensors
    src/mjlab/scene/scene.py:38-45,154-198; ground plane
    src/mjlab/terrains/terrain_importer.py:154-163
  * contact sensors src/mjlab/tasks/velocity/config/g1/env_cfgs.py:31-52,
    config/go1/env_cfgs.py:25-49 via sensor/contact_sensor.py:474-533
  * sim options src/mjlab/tasks/velocity/velocity_env_cfg.py:343-353

The robot MJCF itself is input data read from the reference asset zoo by
`tools/make_models.py`; the compiled result is committed as an .npz under
`mjlab_b200/assets/` so nothing reads /root/reference at run time.
"""

from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

from . import mjcf
from .mjcf import Spec


# -- name matching helpers (reference: utils/string.py:6-45, lab_api/string.py:178-260)
def filter_exp(exprs, names):
  pats = [re.compile(e) for e in exprs]
  return tuple(n for n in names if any(p.match(n) for p in pats))


def resolve_expr(pattern_map, names, default=None):
  pats = [(re.compile(p), v) for p, v in pattern_map.items()]
  out = []
  for n in names:
    for p, v in pats:
      if p.match(n):
        out.append(v)
        break
    else:
      out.append(default)
  return tuple(out)


def resolve_field(fieldval, names, default=None):
  if isinstance(fieldval, dict):
    return resolve_expr(fieldval, names, default)
  return tuple([fieldval] * len(names))


def resolve_matching_names(keys, names):
  if isinstance(keys, str):
    keys = [keys]
  idx, out = [], []
  for i, n in enumerate(names):
    if any(re.fullmatch(k, n) for k in keys):
      idx.append(i)
      out.append(n)
  return idx, out


@dataclass
class PositionActuatorCfg:
  joint_names_expr: tuple[str, ...]
  stiffness: float
  damping: float
  effort_limit: float | None
  armature: float
  frictionloss: float = 0.0


@dataclass
class CollisionCfg:
  geom_names_expr: tuple[str, ...]
  contype: int | dict = 1
  conaffinity: int | dict = 1
  condim: int | dict = 3
  priority: int | dict = 0
  friction: tuple | dict | None = None
  solref: tuple | dict | None = None
  solimp: tuple | dict | None = None
  disable_other_geoms: bool = True

  def edit_spec(self, spec: Spec) -> None:
    names = tuple(g.name for g in spec.geoms)
    subset = filter_exp(self.geom_names_expr, names)
    defaults = dict(condim=3, contype=1, conaffinity=1, priority=0, friction=None, solref=None, solimp=None)
    res = {k: resolve_field(getattr(self, k), subset, d) for k, d in defaults.items()}
    for i, n in enumerate(subset):
      g = spec.geom(n)
      g.condim, g.contype = res["condim"][i], res["contype"][i]
      g.conaffinity, g.priority = res["conaffinity"][i], res["priority"][i]
      for fld in ("friction", "solref", "solimp"):
        v = res[fld][i]
        if v is not None:
          arr = getattr(g, fld)
          for k, x in enumerate(v):
            arr[k] = x
    if self.disable_other_geoms:
      for n in set(names) - set(subset):
        g = spec.geom(n)
        g.contype = 0
        g.conaffinity = 0


@dataclass
class ContactSensorCfg:
  name: str
  primary_mode: str
  primary_names: tuple[str, ...]  # already resolved, un-prefixed
  secondary_mode: str | None
  secondary_name: str | None  # literal model name
  fields: tuple[str, ...] = ("found", "force")
  reduce: str = "maxforce"
  num_slots: int = 1
  entity: str = "robot"


_FIELD_BIT = {"found": 0, "force": 1, "torque": 2, "dist": 3, "pos": 4, "normal": 5, "tangent": 6}
_REDUCE = {"none": 0, "mindist": 1, "maxforce": 2, "netforce": 3}
_MODE = {"geom": mjcf.OBJ_GEOM, "body": mjcf.OBJ_BODY, "subtree": mjcf.OBJ_XBODY}


def add_contact_sensor(scene: Spec, cfg: ContactSensorCfg) -> None:
  for prim in cfg.primary_names:
    for fld in cfg.fields:
      kw = dict(
        name=f"{cfg.name}_{prim}_{fld}", type=mjcf.SENS_CONTACT, objtype=_MODE[cfg.primary_mode],
        objname=f"{cfg.entity}/{prim}", intprm=(1 << _FIELD_BIT[fld], _REDUCE[cfg.reduce], cfg.num_slots),
      )
      if cfg.secondary_name is not None:
        kw["reftype"] = _MODE[cfg.secondary_mode]
        kw["refname"] = cfg.secondary_name
      scene.add_sensor(**kw)


def _two_stage(rotor, gear):
  return rotor[0] * (gear[1] * gear[2]) ** 2 + rotor[1] * gear[2] ** 2 + rotor[2]


_NAT = 10 * 2.0 * 3.1415926535
_ZETA = 2.0


def _pd(arm):
  return arm * _NAT**2, 2.0 * _ZETA * arm * _NAT


def g1_cfg():
  a5020 = _two_stage((0.139e-4, 0.017e-4, 0.169e-4), (1, 1 + 46 / 18, 1 + 56 / 16))
  a7520_14 = _two_stage((0.489e-4, 0.098e-4, 0.533e-4), (1, 4.5, 1 + 48 / 22))
  a7520_22 = _two_stage((0.489e-4, 0.109e-4, 0.738e-4), (1, 4.5, 5))
  a4010 = _two_stage((0.068e-4, 0.0, 0.0), (1, 5, 5))
  k5020, d5020 = _pd(a5020)
  k14, d14 = _pd(a7520_14)
  k22, d22 = _pd(a7520_22)
  k4010, d4010 = _pd(a4010)
  acts = (
    PositionActuatorCfg((".*_elbow_joint", ".*_shoulder_pitch_joint", ".*_shoulder_roll_joint",
                         ".*_shoulder_yaw_joint", ".*_wrist_roll_joint"), k5020, d5020, 25.0, a5020),
    PositionActuatorCfg((".*_hip_pitch_joint", ".*_hip_yaw_joint", "waist_yaw_joint"), k14, d14, 88.0, a7520_14),
    PositionActuatorCfg((".*_hip_roll_joint", ".*_knee_joint"), k22, d22, 139.0, a7520_22),
    PositionActuatorCfg((".*_wrist_pitch_joint", ".*_wrist_yaw_joint"), k4010, d4010, 5.0, a4010),
    PositionActuatorCfg(("waist_pitch_joint", "waist_roll_joint"), k5020 * 2, d5020 * 2, 50.0, a5020 * 2),
    PositionActuatorCfg((".*_ankle_pitch_joint", ".*_ankle_roll_joint"), k5020 * 2, d5020 * 2, 50.0, a5020 * 2),
  )
  foot = r"^(left|right)_foot[1-7]_collision$"
  coll = CollisionCfg(
    geom_names_expr=(".*_collision",),
    condim={foot: 3, ".*_collision": 1}, priority={foot: 1}, friction={foot: (0.6,)},
  )
  init = dict(
    pos=(0, 0, 0.76), rot=(1.0, 0, 0, 0),
    joint_pos={".*_hip_pitch_joint": -0.312, ".*_knee_joint": 0.669, ".*_ankle_pitch_joint": -0.363,
               ".*_elbow_joint": 0.6, "left_shoulder_roll_joint": 0.2, "left_shoulder_pitch_joint": 0.2,
               "right_shoulder_roll_joint": -0.2, "right_shoulder_pitch_joint": 0.2},
  )
  sensors = (
    ContactSensorCfg("feet_ground_contact", "subtree", ("left_ankle_roll_link", "right_ankle_roll_link"),
                     "body", "terrain", ("found", "force"), "netforce", 1),
    ContactSensorCfg("self_collision", "subtree", ("pelvis",), "subtree", "robot/pelvis", ("found",), "none", 1),
  )
  return acts, coll, init, sensors


def go1_cfg():
  rotor = 0.000111842
  a_hip = rotor * 6**2
  a_knee = rotor * 9.0**2
  kh, dh = _pd(a_hip)
  kk, dk = _pd(a_knee)
  acts = (
    PositionActuatorCfg((".*_hip_joint", ".*_thigh_joint"), kh, dh, 23.7, a_hip),
    PositionActuatorCfg((".*_calf_joint",), kk, dk, 35.55, a_knee),
  )
  foot = "^[FR][LR]_foot_collision$"
  coll = CollisionCfg(
    geom_names_expr=(".*_collision",), condim={foot: 3, ".*_collision": 1}, priority={foot: 1},
    friction={foot: (0.6,)}, solimp={foot: (0.9, 0.95, 0.023)}, contype=1, conaffinity=0,
  )
  init = dict(
    pos=(0.0, 0.0, 0.278), rot=(1.0, 0, 0, 0),
    joint_pos={".*thigh_joint": 0.9, ".*calf_joint": -1.8, ".*R_hip_joint": 0.1, ".*L_hip_joint": -0.1},
  )
  feet = tuple(f"{n}_foot_collision" for n in ("FR", "FL", "RR", "RL"))
  sensors = (
    ContactSensorCfg("feet_ground_contact", "geom", feet, "body", "terrain", ("found", "force"), "netforce", 1),
    # primaries of nonfoot_ground_touch are resolved against the geom list in build_scene
    ContactSensorCfg("nonfoot_ground_touch", "geom", ("__NONFOOT__",), "body", "terrain", ("found",), "none", 1),
  )
  return acts, coll, init, sensors


def build_entity(robot: Spec, acts, coll: CollisionCfg, init) -> Spec:
  """Entity.__init__ spec edits (entity/entity.py:127-207)."""
  coll.edit_spec(robot)
  joint_names = [j.name for j in robot.joints if j.type != mjcf.JNT_FREE]
  for a in acts:
    _, names = resolve_matching_names(a.joint_names_expr, joint_names)
    if not names:
      raise ValueError(f"No joints found for actuator with expressions: {a.joint_names_expr}")
    for n in names:
      act = robot.add_actuator(name=n, target=n, gaintype=mjcf.GAIN_FIXED, biastype=mjcf.BIAS_AFFINE)
      act.gainprm[0] = a.stiffness
      act.biasprm[1] = -a.stiffness
      act.biasprm[2] = -a.damping
      if a.effort_limit is not None:
        act.forcelimited = True
        act.forcerange = np.array([-a.effort_limit, a.effort_limit])
      j = robot.joint(n)
      j.armature = a.armature
      j.frictionloss = a.frictionloss
  jp = resolve_expr(init["joint_pos"], tuple(joint_names), 0.0)
  has_free = any(j.type == mjcf.JNT_FREE for j in robot.joints)
  comps = []
  if has_free:
    comps += [np.array(init["pos"], float), np.array(init["rot"], float)]
  comps.append(np.array(jp, float))
  name_to_pos = dict(zip(joint_names, jp))
  ctrl = np.array([name_to_pos.get(a.target, 0.0) for a in robot.actuators], float)
  robot.keys.clear()
  robot.add_key(name="init_state", qpos=np.hstack(comps), ctrl=ctrl)
  return robot


def build_scene(robot_xml: str, which: str, *, terrain=None, timestep=0.005, iterations=10, ls_iterations=20) -> Spec:
  """terrain + robot + sensors, as Scene.__init__ (scene/scene.py:29-45).  `terrain=None`: the ground plane
  (terrain_importer.py:154-163); a `terrains.Terrain`: its boxes as geoms `terrain_<i>` of the body "terrain"
  (terrain_generator.py:93-116)."""
  acts, coll, init, sensors = {"g1": g1_cfg, "go1": go1_cfg}[which]()
  robot = build_entity(Spec.from_string(robot_xml), acts, coll, init)
  scene = Spec()
  tbody = scene.worldbody.add_body(name="terrain")
  if terrain is None:
    tbody.add_geom(name="terrain", type=mjcf.GEOM_PLANE, size=(0, 0, 0.01))
  else:
    for i, row in enumerate(terrain.boxes):
      tbody.add_geom(name=f"terrain_{i}", type=mjcf.GEOM_BOX, pos=row[:3], size=row[3:6])
  key = robot.keys[0]
  robot.keys.clear()
  scene.attach(robot, prefix="robot/")
  scene.add_key(name="init_state", qpos=key.qpos, ctrl=key.ctrl)
  for s in sensors:
    if s.primary_names == ("__NONFOOT__",):
      geoms = [g.name.removeprefix("robot/") for g in scene.geoms if g.name.startswith("robot/")]
      _, names = resolve_matching_names(r".*_collision\d*$", geoms)
      excl = set(f"{n}_foot_collision" for n in ("FR", "FL", "RR", "RL"))
      s = ContactSensorCfg(s.name, s.primary_mode, tuple(n for n in names if n not in excl), s.secondary_mode,
                           s.secondary_name, s.fields, s.reduce, s.num_slots)
    add_contact_sensor(scene, s)
  o = scene.option
  o.timestep, o.iterations, o.ls_iterations = timestep, iterations, ls_iterations
  o.integrator = mjcf.INT_IMPLICITFAST
  o.cone, o.solver, o.impratio = mjcf.CONE_PYRAMIDAL, mjcf.SOL_NEWTON, 1.0
  o.tolerance, o.ls_tolerance = 1e-8, 0.01
  o.gravity = np.array([0.0, 0.0, -9.81])
  return scene
