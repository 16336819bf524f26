"""Resolved configuration of the velocity task (G1 / Go1, flat) for the env kernels.

Restates, as plain arrays, what the reference resolves at `load_managers()` time from
`make_velocity_env_cfg()` (src/mjlab/tasks/velocity/velocity_env_cfg.py:33-353) and the per-robot overrides
(config/g1/env_cfgs.py:20-173, config/go1/env_cfgs.py:15-127): index maps of the robot entity
(entity/entity.py:754-813), default state and soft joint limits (entity/entity.py:409-462), action scale
(g1_constants.py:287-295), sensor addresses, reward weights / parameters, observation noise ranges, command,
push and reset ranges.
"""

from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np

def cfg_to_dict(obj, _depth: int = 0):
  """Plain-data view of a (reference) configuration object: dataclasses -> dicts of their fields, callables / classes -> their
  `__name__`, tuples -> lists, `slice(None)` -> None, tensors / arrays -> lists.  `from_manager_cfg` works on this view, so it takes
  the reference's live `ManagerBasedRlEnvCfg` and the JSON fixtures generated from it alike."""
  import dataclasses
  if _depth > 12:
    return repr(obj)
  if obj is None or isinstance(obj, (bool, int, float, str)):
    return obj
  if isinstance(obj, slice):
    return None if obj == slice(None) else [obj.start, obj.stop, obj.step]
  if isinstance(obj, dict):
    return {str(k): cfg_to_dict(v, _depth + 1) for k, v in obj.items()}
  if isinstance(obj, (list, tuple, set)):
    return [cfg_to_dict(v, _depth + 1) for v in obj]
  if isinstance(obj, np.ndarray):
    return obj.tolist()
  if hasattr(obj, "detach") and hasattr(obj, "tolist"):
    return obj.detach().cpu().tolist()
  if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
    d = {f.name: cfg_to_dict(getattr(obj, f.name), _depth + 1) for f in dataclasses.fields(obj)}
    d["__type__"] = type(obj).__name__
    return d
  if callable(obj):
    return getattr(obj, "__name__", type(obj).__name__)
  return repr(obj)


REWARD_NAMES = (
  "track_linear_velocity", "track_angular_velocity", "upright", "pose", "body_ang_vel", "angular_momentum",
  "dof_pos_limits", "action_rate_l2", "air_time", "foot_clearance", "foot_swing_height", "foot_slip", "soft_landing",
  "self_collisions",
)
TERMINATION_NAMES = ("time_out", "fell_over", "illegal_contact")

_G1_STD_WALK = {
  r".*hip_pitch.*": 0.3, r".*hip_roll.*": 0.15, r".*hip_yaw.*": 0.15, r".*knee.*": 0.35, r".*ankle_pitch.*": 0.25,
  r".*ankle_roll.*": 0.1, r".*waist_yaw.*": 0.2, r".*waist_roll.*": 0.08, r".*waist_pitch.*": 0.1,
  r".*shoulder_pitch.*": 0.15, r".*shoulder_roll.*": 0.15, r".*shoulder_yaw.*": 0.1, r".*elbow.*": 0.15, r".*wrist.*": 0.3,
}
_G1_STD_RUN = {
  r".*hip_pitch.*": 0.5, r".*hip_roll.*": 0.2, r".*hip_yaw.*": 0.2, r".*knee.*": 0.6, r".*ankle_pitch.*": 0.35,
  r".*ankle_roll.*": 0.15, r".*waist_yaw.*": 0.3, r".*waist_roll.*": 0.08, r".*waist_pitch.*": 0.2,
  r".*shoulder_pitch.*": 0.5, r".*shoulder_roll.*": 0.2, r".*shoulder_yaw.*": 0.15, r".*elbow.*": 0.35, r".*wrist.*": 0.3,
}
_GO1_HT = r".*(FR|FL|RR|RL)_(hip|thigh)_joint.*"
_GO1_C = r".*(FR|FL|RR|RL)_calf_joint.*"


@dataclass
class VelocityCfg:
  robot: str
  nj: int
  nfeet: int
  root_body: int
  body_ids: np.ndarray
  geom_ids: np.ndarray
  site_ids: np.ndarray
  joint_ids: np.ndarray
  joint_names: list
  joint_q_adr: np.ndarray
  joint_v_adr: np.ndarray
  ctrl_ids: np.ndarray
  default_root_state: np.ndarray
  default_joint_pos: np.ndarray
  joint_pos_limits: np.ndarray
  soft_joint_pos_limits: np.ndarray
  action_scale: np.ndarray
  foot_site_names: tuple
  foot_site_local_ids: tuple
  foot_site_ids: np.ndarray
  foot_found_adr: np.ndarray
  foot_force_adr: np.ndarray
  linvel_adr: int
  angvel_adr: int
  angmom_adr: int
  selfcol_adr: int
  illegal_adr: int
  illegal_count: int
  upright_body: int
  upright_body_local: int
  reward_weights: np.ndarray
  std_standing: np.ndarray
  std_walking: np.ndarray
  std_running: np.ndarray
  std_standing_map: dict
  std_walking_map: dict
  std_running_map: dict
  noise_lo: np.ndarray
  noise_hi: np.ndarray
  n_obs_policy: int
  n_obs_critic: int
  step_dt: float = 0.02
  decimation: int = 4
  max_episode_length: int = 1000
  max_episode_length_s: float = 20.0
  limit_angle: float = math.radians(70.0)
  # reward parameters (velocity_env_cfg.py:183-278)
  std_lin: float = math.sqrt(0.25)
  std_ang: float = math.sqrt(0.5)
  std_upright: float = math.sqrt(0.2)
  walking_threshold: float = 0.05
  running_threshold: float = 1.5
  air_thr_min: float = 0.05
  air_thr_max: float = 0.5
  air_cmd_threshold: float = 0.5
  target_height: float = 0.1
  cmd_threshold: float = 0.05
  # command (velocity_env_cfg.py:121-139)
  cmd_resample: tuple = (3.0, 8.0)
  cmd_lin_x: tuple = (-1.0, 1.0)
  cmd_lin_y: tuple = (-1.0, 1.0)
  cmd_ang_z: tuple = (-0.5, 0.5)
  cmd_heading: tuple = (-math.pi, math.pi)
  rel_standing: float = 0.1
  rel_heading: float = 0.3
  heading_stiffness: float = 0.5
  # events (velocity_env_cfg.py:145-181)
  reset_x: tuple = (-0.5, 0.5)
  reset_y: tuple = (-0.5, 0.5)
  reset_yaw: tuple = (-3.14, 3.14)
  reset_joint_pos: tuple = (0.0, 0.0)
  reset_joint_vel: tuple = (0.0, 0.0)
  push_interval: tuple = (1.0, 3.0)
  push_x: tuple = (-0.5, 0.5)
  push_y: tuple = (-0.5, 0.5)
  # full six-axis ranges {"x" | "y" | "z" | "roll" | "pitch" | "yaw": (lo, hi)} as the reference's event params spell them
  # (envs/mdp/events.py:104-107,141-146,217-221); when given they replace the x / y / yaw shorthands above
  reset_pose_range: dict = None
  reset_velocity_range: dict = None
  push_velocity_range: dict = None
  enable_corruption: bool = True
  push_enabled: bool = True
  heading_command: bool = True
  seed: int = 0
  foot_site_body: np.ndarray = None  # body of each foot site (site velocities, entity/data.py:323-331)
  # rough terrain (velocity_env_cfg.py:296-324): bit 0 = terrain_levels_vel curriculum, bit 1 = play-mode randomize_terrain
  terrain_mode: int = 0
  terrain_origins: np.ndarray = None  # [rows, cols, 3]
  terrain_half_size: float = 4.0
  max_init_terrain_level: int = 5
  extra: dict = field(default_factory=dict)

  @property
  def terrain_rows(self) -> int:
    return 0 if self.terrain_origins is None else int(self.terrain_origins.shape[0])

  @property
  def terrain_cols(self) -> int:
    return 0 if self.terrain_origins is None else int(self.terrain_origins.shape[1])

  @staticmethod
  def _ranges6(d: dict) -> np.ndarray:
    """{"x": (lo, hi), ...} -> [6, 2] in the order x y z roll pitch yaw (envs/mdp/events.py:104-107)."""
    return np.array([d.get(k, (0.0, 0.0)) for k in ("x", "y", "z", "roll", "pitch", "yaw")], float)

  @property
  def reset_pose(self) -> np.ndarray:
    if self.reset_pose_range is not None:
      return self._ranges6(self.reset_pose_range)
    return self._ranges6({"x": self.reset_x, "y": self.reset_y, "yaw": self.reset_yaw})

  @property
  def reset_vel(self) -> np.ndarray:
    return self._ranges6(self.reset_velocity_range or {})

  @property
  def push_vel(self) -> np.ndarray:
    if self.push_velocity_range is not None:
      return self._ranges6(self.push_velocity_range)
    return self._ranges6({"x": self.push_x, "y": self.push_y})


def _resolve_values(mapping: dict, names: list[str]) -> np.ndarray:
  """resolve_matching_names_values (utils/lab_api/string.py:274-340): each name takes the value of the key it fullmatches."""
  out = []
  for n in names:
    hit = [v for k, v in mapping.items() if re.fullmatch(k, n)]
    if len(hit) != 1:
      raise ValueError(f"'{n}' matches {len(hit)} keys of {list(mapping)}")
    out.append(hit[0])
  return np.array(out, float)


def velocity_cfg(m, robot: str, play: bool = False) -> VelocityCfg:
  pre = "robot/"
  body_ids = np.array([i for i, n in enumerate(m.body_names) if n.startswith(pre)])
  geom_ids = np.array([i for i, n in enumerate(m.geom_names) if n.startswith(pre)])
  site_ids = np.array([i for i, n in enumerate(m.site_names) if n.startswith(pre)])
  joint_ids = np.array([j for j in range(m.njnt) if m.jnt_type[j] != 0 and m.jnt_names[j].startswith(pre)])
  joint_names = [m.jnt_names[j] for j in joint_ids]
  local_names = [n.removeprefix(pre) for n in joint_names]
  nj = len(joint_ids)
  joint_q_adr = m.jnt_qposadr[joint_ids]
  joint_v_adr = m.jnt_dofadr[joint_ids]
  act_of_joint = {int(m.actuator_trnid[a, 0]): a for a in range(m.nu)}
  ctrl_ids = np.array([act_of_joint[int(j)] for j in joint_ids])
  kq = m.key_qpos[0]
  default_root_state = np.concatenate([kq[0:7], np.zeros(6)])
  default_joint_pos = kq[joint_q_adr]
  lim = m.jnt_range[joint_ids]
  mean, rng = 0.5 * (lim[:, 0] + lim[:, 1]), lim[:, 1] - lim[:, 0]
  soft = np.stack([mean - 0.5 * rng * 0.9, mean + 0.5 * rng * 0.9], 1)
  scale = 0.25 * m.actuator_forcerange[ctrl_ids, 1] / m.actuator_gainprm[ctrl_ids, 0]

  def sadr(name):
    return int(m.sensor_adr[m.sensor_names.index(name)])

  site_local = [n.removeprefix(pre) for n in (m.site_names[i] for i in site_ids)]
  body_local = [n.removeprefix(pre) for n in (m.body_names[i] for i in body_ids)]
  if robot == "g1":
    feet_sites = ("left_foot", "right_foot")
    prims = ("left_ankle_roll_link", "right_ankle_roll_link")
    upright = "torso_link"
    weights = dict(track_linear_velocity=2.0, track_angular_velocity=2.0, upright=1.0, pose=1.0, body_ang_vel=-0.05,
                   angular_momentum=-0.02, dof_pos_limits=-1.0, action_rate_l2=-0.1, air_time=0.0, foot_clearance=-2.0,
                   foot_swing_height=-0.25, foot_slip=-0.1, soft_landing=-1e-5, self_collisions=-1.0)
    std_stand, std_walk, std_run = {".*": 0.05}, _G1_STD_WALK, _G1_STD_RUN
    selfcol = sadr("self_collision_pelvis_found")
    illegal_adr, illegal_count = -1, 0
  elif robot == "go1":
    feet_sites = ("FR", "FL", "RR", "RL")
    prims = tuple(f"{n}_foot_collision" for n in feet_sites)
    upright = "trunk"
    weights = dict(track_linear_velocity=2.0, track_angular_velocity=2.0, upright=1.0, pose=1.0, body_ang_vel=0.0,
                   angular_momentum=0.0, dof_pos_limits=-1.0, action_rate_l2=-0.1, air_time=0.0, foot_clearance=-2.0,
                   foot_swing_height=-0.25, foot_slip=-0.1, soft_landing=-1e-5, self_collisions=0.0)
    std_stand = {_GO1_HT: 0.05, _GO1_C: 0.1}
    std_walk = {_GO1_HT: 0.3, _GO1_C: 0.6}
    std_run = {_GO1_HT: 0.3, _GO1_C: 0.6}
    selfcol = -1
    nf_names = [n for n in m.sensor_names if n.startswith("nonfoot_ground_touch_")]
    illegal_adr, illegal_count = sadr(nf_names[0]), len(nf_names)
    assert [sadr(n) for n in nf_names] == list(range(illegal_adr, illegal_adr + illegal_count))
  else:
    raise ValueError(robot)
  foot_local = tuple(site_local.index(s) for s in feet_sites)
  n_pol = 3 + 3 + 3 + nj + nj + nj + 3
  nfeet = len(feet_sites)
  n_cri = n_pol + nfeet + nfeet + nfeet + 3 * nfeet
  lo = np.zeros(n_pol)
  hi = np.zeros(n_pol)
  for a, b, (l, h) in ((0, 3, (-0.5, 0.5)), (3, 6, (-0.2, 0.2)), (6, 9, (-0.05, 0.05)), (9, 9 + nj, (-0.01, 0.01)),
                       (9 + nj, 9 + 2 * nj, (-1.5, 1.5))):
    lo[a:b], hi[a:b] = l, h
  return VelocityCfg(
    robot=robot, nj=nj, nfeet=nfeet, root_body=int(body_ids[0]), body_ids=body_ids, geom_ids=geom_ids, site_ids=site_ids,
    joint_ids=joint_ids, joint_names=joint_names, joint_q_adr=joint_q_adr, joint_v_adr=joint_v_adr, ctrl_ids=ctrl_ids,
    default_root_state=default_root_state, default_joint_pos=default_joint_pos, joint_pos_limits=lim,
    soft_joint_pos_limits=soft, action_scale=scale, foot_site_names=feet_sites, foot_site_local_ids=foot_local,
    foot_site_ids=site_ids[list(foot_local)],
    foot_found_adr=np.array([sadr(f"feet_ground_contact_{p}_found") for p in prims]),
    foot_force_adr=np.array([sadr(f"feet_ground_contact_{p}_force") for p in prims]),
    linvel_adr=sadr("robot/imu_lin_vel"), angvel_adr=sadr("robot/imu_ang_vel"), angmom_adr=sadr("robot/root_angmom"),
    selfcol_adr=selfcol, illegal_adr=illegal_adr, illegal_count=illegal_count,
    upright_body=int(body_ids[body_local.index(upright)]), upright_body_local=body_local.index(upright),
    reward_weights=np.array([weights[n] for n in REWARD_NAMES]),
    std_standing=_resolve_values(std_stand, local_names), std_walking=_resolve_values(std_walk, local_names),
    std_running=_resolve_values(std_run, local_names), std_standing_map=std_stand, std_walking_map=std_walk,
    std_running_map=std_run, noise_lo=lo, noise_hi=hi, n_obs_policy=n_pol, n_obs_critic=n_cri,
    max_episode_length=math.ceil(int(1e9) / 0.02) if play else 1000, max_episode_length_s=float(int(1e9)) if play else 20.0,
    enable_corruption=not play, push_enabled=not play,
    foot_site_body=np.asarray(m.site_bodyid)[site_ids[list(foot_local)]].astype(int),
    **_terrain_fields(m, play),
  )


def _terrain_fields(m, play: bool) -> dict:
  """Scenes compiled with a generated terrain carry its spawn origins (tools/make_models.py)."""
  org = getattr(m, "terrain_origins", None)
  if org is None:
    return {}
  return dict(terrain_mode=1 | (2 if play else 0), terrain_origins=np.asarray(org, float),
              terrain_half_size=float(np.asarray(m.terrain_size)[0]) / 2.0)


# ---------------------------------------------------------------------------------------------------------
# tracking task (tasks/tracking/tracking_env_cfg.py:44-317, config/g1/env_cfgs.py:15-100)
# ---------------------------------------------------------------------------------------------------------
TRACKING_REWARD_NAMES = ("motion_global_root_pos", "motion_global_root_ori", "motion_body_pos", "motion_body_ori", "motion_body_lin_vel",
                         "motion_body_ang_vel", "action_rate_l2", "joint_limit", "self_collisions")
TRACKING_TERMINATION_NAMES = ("time_out", "anchor_pos", "anchor_ori", "ee_body_pos")
G1_TRACKED_BODIES = ("pelvis", "left_hip_roll_link", "left_knee_link", "left_ankle_roll_link", "right_hip_roll_link", "right_knee_link",
                     "right_ankle_roll_link", "torso_link", "left_shoulder_roll_link", "left_elbow_link", "left_wrist_yaw_link",
                     "right_shoulder_roll_link", "right_elbow_link", "right_wrist_yaw_link")
G1_EE_BODIES = ("left_ankle_roll_link", "right_ankle_roll_link", "left_wrist_yaw_link", "right_wrist_yaw_link")
_TRACKING_VELOCITY_RANGE = {"x": (-0.5, 0.5), "y": (-0.5, 0.5), "z": (-0.2, 0.2), "roll": (-0.52, 0.52), "pitch": (-0.52, 0.52), "yaw": (-0.78, 0.78)}


@dataclass
class TrackingCfg:
  nj: int
  nbody: int               # tracked bodies
  nbody_motion: int        # robot bodies stored per motion frame
  root_body: int
  body_ids: np.ndarray     # global ids of all robot bodies (motion order)
  body_sel: np.ndarray     # tracked bodies as indices into body_ids
  anchor_sel: int          # anchor inside the tracked list
  ee_sel: np.ndarray       # end-effector bodies inside the tracked list
  joint_q_adr: np.ndarray
  joint_v_adr: np.ndarray
  ctrl_ids: np.ndarray
  default_joint_pos: np.ndarray
  soft_joint_pos_limits: np.ndarray
  action_scale: np.ndarray
  linvel_adr: int
  angvel_adr: int
  selfcol_adr: int
  noise_lo: np.ndarray
  noise_hi: np.ndarray
  n_obs_policy: int
  n_obs_critic: int
  time_step_total: int
  bin_count: int
  reward_weights: np.ndarray
  step_dt: float = 0.02
  max_episode_length: int = 500
  max_episode_length_s: float = 10.0
  anchor_pos_threshold: float = 0.25
  anchor_ori_threshold: float = 0.8
  ee_pos_threshold: float = 0.25
  std_anchor_pos: float = 0.3
  std_anchor_ori: float = 0.4
  std_body_pos: float = 0.3
  std_body_ori: float = 0.4
  std_body_lin_vel: float = 1.0
  std_body_ang_vel: float = 3.14
  sampling_mode: str = "adaptive"
  adaptive_kernel_size: int = 1
  adaptive_lambda: float = 0.8
  adaptive_uniform_ratio: float = 0.1
  adaptive_alpha: float = 0.001
  pose_range: np.ndarray = None
  velocity_range: np.ndarray = None
  joint_position_range: tuple = (-0.1, 0.1)
  push_interval: tuple = (1.0, 3.0)
  push_vel: np.ndarray = None
  enable_corruption: bool = True
  push_enabled: bool = True
  seed: int = 0


def tracking_cfg(m, time_step_total: int, robot: str = "g1", play: bool = False, tracked=G1_TRACKED_BODIES, anchor="torso_link", ee=G1_EE_BODIES) -> TrackingCfg:
  v = velocity_cfg(m, robot)   # index maps, default pose, soft limits, action scale, sensor addresses are shared
  body_local = [n.removeprefix("robot/") for n in (m.body_names[i] for i in v.body_ids)]
  body_sel = np.array([body_local.index(n) for n in tracked])
  nj, nb = v.nj, len(tracked)
  n_pol = 5 * nj + 15
  lo, hi = np.zeros(n_pol), np.zeros(n_pol)
  c = 2 * nj
  for a, b, (l, h) in ((c, c + 3, (-0.25, 0.25)), (c + 3, c + 9, (-0.05, 0.05)), (c + 9, c + 12, (-0.5, 0.5)), (c + 12, c + 15, (-0.2, 0.2)),
                       (c + 15, c + 15 + nj, (-0.01, 0.01)), (c + 15 + nj, c + 15 + 2 * nj, (-0.5, 0.5))):
    lo[a:b], hi[a:b] = l, h
  r6 = VelocityCfg._ranges6
  step_dt = 0.02
  return TrackingCfg(
    nj=nj, nbody=nb, nbody_motion=len(v.body_ids), root_body=v.root_body, body_ids=v.body_ids, body_sel=body_sel,
    anchor_sel=list(tracked).index(anchor), ee_sel=np.array([list(tracked).index(n) for n in ee]), joint_q_adr=v.joint_q_adr,
    joint_v_adr=v.joint_v_adr, ctrl_ids=v.ctrl_ids, default_joint_pos=v.default_joint_pos, soft_joint_pos_limits=v.soft_joint_pos_limits,
    action_scale=v.action_scale, linvel_adr=v.linvel_adr, angvel_adr=v.angvel_adr, selfcol_adr=v.selfcol_adr, noise_lo=lo, noise_hi=hi,
    n_obs_policy=n_pol, n_obs_critic=n_pol + 9 * nb, time_step_total=int(time_step_total),
    bin_count=int(time_step_total // (1 / step_dt)) + 1,
    reward_weights=np.array([0.5, 0.5, 1.0, 1.0, 1.0, 1.0, -0.1, -10.0, -10.0]), step_dt=step_dt,
    max_episode_length=math.ceil(int(1e9) / step_dt) if play else 500, max_episode_length_s=float(int(1e9)) if play else 10.0,
    sampling_mode="start" if play else "adaptive",
    pose_range=r6({} if play else {"x": (-0.05, 0.05), "y": (-0.05, 0.05), "z": (-0.01, 0.01), "roll": (-0.1, 0.1), "pitch": (-0.1, 0.1), "yaw": (-0.2, 0.2)}),
    velocity_range=r6({} if play else _TRACKING_VELOCITY_RANGE), push_vel=r6(_TRACKING_VELOCITY_RANGE),
    enable_corruption=not play, push_enabled=not play,
  )
