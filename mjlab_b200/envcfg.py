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
  # the generic library of envs/mdp/rewards.py (weight 0 in the shipped tasks; a consumed ManagerBasedRlEnvCfg may use them)
  "is_alive", "is_terminated", "joint_torques_l2", "joint_vel_l2", "joint_acc_l2", "action_acc_l2", "flat_orientation_l2",
  "electrical_power_cost",
)
TERMINATION_NAMES = ("time_out", "fell_over", "illegal_contact", "root_height_below_minimum", "nan_detection")

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
  min_root_height: float | None = None     # root_height_below_minimum termination (envs/mdp/terminations.py:35-42); None = off
  nan_termination: bool = False            # nan_detection termination (envs/mdp/terminations.py:45-47)
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
    reward_weights=np.array([weights.get(n, 0.0) for n in REWARD_NAMES]),
    std_standing=_resolve_values(std_stand, local_names), std_walking=_resolve_values(std_walk, local_names),
    std_running=_resolve_values(std_run, local_names), std_standing_map=std_stand, std_walking_map=std_walk,
    std_running_map=std_run, noise_lo=lo, noise_hi=hi, n_obs_policy=n_pol, n_obs_critic=n_cri,
    max_episode_length=math.ceil(int(1e9) / 0.02) if play else 1000, max_episode_length_s=float(int(1e9)) if play else 20.0,
    enable_corruption=not play, push_enabled=not play,
    foot_site_body=np.asarray(m.site_bodyid)[site_ids[list(foot_local)]].astype(int),
    **_terrain_fields(m, play and robot == "g1"),   # only the G1 play cfg adds the randomize_terrain reset event (config/g1/env_cfgs.py:136-140; go1: none)
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
  term_full3d: int = 0       # bit 0: bad_anchor_pos instead of bad_anchor_pos_z_only; bit 1: bad_motion_body_pos instead of ..._z_only
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
  extra: dict = field(default_factory=dict)   # what the kernels do not consume themselves (startup events, sim options)


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


# ---------------------------------------------------------------------------------------------------------
# consuming a reference ManagerBasedRlEnvCfg (envs/manager_based_rl_env.py:59-129)
# ---------------------------------------------------------------------------------------------------------
# kernel reward slot -> the reference term FUNCTION it implements (tasks/velocity/mdp/rewards.py, envs/mdp/rewards.py)
REWARD_FUNCS = {
  "track_linear_velocity": "track_linear_velocity", "track_angular_velocity": "track_angular_velocity", "upright": "flat_orientation",
  "pose": "variable_posture", "body_ang_vel": "body_angular_velocity_penalty", "angular_momentum": "angular_momentum_penalty",
  "dof_pos_limits": "joint_pos_limits", "action_rate_l2": "action_rate_l2", "air_time": "feet_air_time", "foot_clearance": "feet_clearance",
  "foot_swing_height": "feet_swing_height", "foot_slip": "feet_slip", "soft_landing": "soft_landing", "self_collisions": "self_collision_cost",
  "is_alive": "is_alive", "is_terminated": "is_terminated", "joint_torques_l2": "joint_torques_l2", "joint_vel_l2": "joint_vel_l2",
  "joint_acc_l2": "joint_acc_l2", "action_acc_l2": "action_acc_l2", "flat_orientation_l2": "flat_orientation_l2",
  "electrical_power_cost": "electrical_power_cost",
}
# observation terms the fused kernel writes, in its fixed order: (term function, sensor_name param or None)
POLICY_OBS_TERMS = (("builtin_sensor", "robot/imu_lin_vel"), ("builtin_sensor", "robot/imu_ang_vel"), ("projected_gravity", None),
                    ("joint_pos_rel", None), ("joint_vel_rel", None), ("last_action", None), ("generated_commands", None))
CRITIC_EXTRA_OBS_TERMS = (("foot_height", None), ("foot_air_time", None), ("foot_contact", None), ("foot_contact_forces", None))
_POST_KEYS = ("clip", "scale", "delay_min_lag", "delay_max_lag", "delay_per_env", "delay_hold_prob", "delay_update_period",
              "delay_per_env_phase", "history_length", "flatten_history_dim")


def _startup_actuator_event(name: str, f: str, p: dict) -> dict:
  """randomize_pd_gains / randomize_effort_limits / randomize_encoder_bias as startup events (envs/mdp/events.py:519-744): applied
  once by the env through `events.apply_startup_event` (model fields expanded per world first)."""
  ac = p.get("asset_cfg") or {"name": "robot"}
  keys = ("name", "joint_names", "actuator_names")
  params = {k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in p.items() if k != "asset_cfg"}
  return dict(name=name, kind=f, params=params,
              asset_cfg={k: (tuple(ac[k]) if isinstance(ac.get(k), (list, tuple)) else ac.get(k)) for k in keys})


class UnsupportedTermError(NotImplementedError):
  """A term / option of the configuration that the fused env kernels do not implement (never silently ignored)."""


def _range6(d) -> dict:
  return {k: (float(v[0]), float(v[1])) for k, v in (d or {}).items()}



# ---- parameters of a consumed term: every one is either read into the kernels' table, or a selector that must name what the
# kernels are wired to (the feet sites, the contact sensors, the "twist" command, all actuated joints); anything else raises
def _ac(v) -> dict:
  """The non-default selectors of a SceneEntityCfg dump."""
  v = v or {}
  out = {k: (tuple(x) if isinstance(x, (list, tuple)) else x) for k, x in v.items()
         if k.endswith(("_names", "_ids")) and x not in (None, False) and not (isinstance(x, dict))}
  if v.get("name", "robot") != "robot":
    out["name"] = v["name"]
  if v.get("preserve_order"):
    out["preserve_order"] = True
  return out


def _all_joints(v) -> bool:
  a = _ac(v)
  return a == {} or a == {"joint_names": (".*",)}


def _check_params(kind: str, name: str, func: str, params: dict, consumed: tuple, fixed: dict) -> None:
  for k, v in params.items():
    if k in consumed:
      continue
    if k not in fixed:
      raise UnsupportedTermError(f"{kind} '{name}' ({func}): parameter '{k}' is not consumed by the env kernels")
    want = fixed[k]
    ok = want(v) if callable(want) else v == want
    if not ok:
      raise UnsupportedTermError(f"{kind} '{name}' ({func}): {k}={v!r} - the env kernels are wired to "
                                 f"{want.__doc__ if callable(want) else repr(want)}")

def from_manager_cfg(cfg, m, robot: str) -> VelocityCfg:
  """Resolve a reference `ManagerBasedRlEnvCfg` of the velocity task (the live dataclass or its `cfg_to_dict` view) into the
  kernels' `VelocityCfg`: reward weights and parameters by term FUNCTION, termination set, observation noise / corruption flags,
  action scale, reset / push / friction events, command ranges, curricula, episode length, decimation.  The kernels implement the
  shipped term functions in a fixed observation layout; anything else raises `UnsupportedTermError` naming the term.  Per-term
  clip / scale / delay / history settings are returned in `extra["obs_post"]` for `obs_pipeline.ObservationPipeline`."""
  d = cfg if isinstance(cfg, dict) else cfg_to_dict(cfg)
  c = velocity_cfg(m, robot, play=False)
  pre = "robot/"
  local_names = [n.removeprefix(pre) for n in c.joint_names]
  # ---- timing ---------------------------------------------------------------------------------------------------------
  mj = d["sim"]["mujoco"]
  if mj["solver"] not in ("newton", "cg") or mj["cone"] not in ("pyramidal", "elliptic") or mj["integrator"] not in ("implicitfast", "euler"):
    raise UnsupportedTermError(f"sim options {mj['solver']} / {mj['cone']} / {mj['integrator']} are outside the supported subset")
  c.decimation = int(d["decimation"])
  c.step_dt = float(mj["timestep"]) * c.decimation
  c.max_episode_length_s = float(d["episode_length_s"])
  c.max_episode_length = math.ceil(c.max_episode_length_s / c.step_dt)
  c.extra["sim"] = dict(timestep=float(mj["timestep"]), iterations=int(mj["iterations"]), ls_iterations=int(mj["ls_iterations"]),
                        integrator=mj["integrator"], nconmax=d["sim"]["nconmax"], njmax=d["sim"]["njmax"])
  # ---- rewards --------------------------------------------------------------------------------------------------------
  by_func = {}
  for name, t in d["rewards"].items():
    if t is None:
      continue
    if t["func"] in by_func:
      raise UnsupportedTermError(f"reward function '{t['func']}' appears twice ('{by_func[t['func']][0]}', '{name}')")
    by_func[t["func"]] = (name, t)
  if "posture" in by_func:
    # envs/mdp/rewards.py:91-122: exp(-mean((q - q0)^2 / std^2)) with ONE std table = variable_posture (tasks/velocity/mdp/
    # rewards.py:291-359) whose three tables are equal (std s = s*1 + s*0 + s*0 whatever the commanded speed): same kernel slot
    if "variable_posture" in by_func:
      raise UnsupportedTermError(f"reward terms '{by_func['posture'][0]}' (posture) and '{by_func['variable_posture'][0]}' (variable_posture) share one kernel slot")
    pname, pt = by_func.pop("posture")
    pp = dict(pt["params"])
    std = pp.pop("std")
    by_func["variable_posture"] = (pname, dict(pt, func="variable_posture", params=dict(
      pp, std_standing=std, std_walking=std, std_running=std, walking_threshold=0.5, running_threshold=1.5, command_name="twist")))
  unknown = set(by_func) - set(REWARD_FUNCS.values())
  if unknown:
    raise UnsupportedTermError(f"reward terms {sorted(by_func[f][0] for f in unknown)} ({sorted(unknown)}) are not implemented by the env kernels")
  w = np.zeros(len(REWARD_NAMES))
  P = {}
  for k, slot in enumerate(REWARD_NAMES):
    if REWARD_FUNCS[slot] in by_func:
      _, t = by_func[REWARD_FUNCS[slot]]
      w[k] = float(t["weight"])
      P[slot] = t["params"]
  c.reward_weights = w
  feet = tuple(c.foot_site_names)

  def foot_sites(v):
    return _ac(v) == {"site_names": feet}
  foot_sites.__doc__ = f"the foot sites {feet}"

  def all_joints(v):
    return _all_joints(v)
  all_joints.__doc__ = "all actuated joints of 'robot'"

  def upright_body(v):
    return _ac(v) == _ac((P.get("upright") or {}).get("asset_cfg")) if "upright" in P else len(_ac(v).get("body_names", ())) == 1
  upright_body.__doc__ = "the body of the flat_orientation term (one body serves both)"
  twist, fgc = {"command_name": "twist"}, {"sensor_name": "feet_ground_contact"}
  rules = {
    "track_linear_velocity": (("std",), twist), "track_angular_velocity": (("std",), twist),
    "upright": (("std", "asset_cfg"), {}),
    "pose": (("std_standing", "std_walking", "std_running", "walking_threshold", "running_threshold"), {**twist, "asset_cfg": all_joints}),
    "body_ang_vel": ((), {"asset_cfg": upright_body}),
    "angular_momentum": ((), {"sensor_name": "robot/root_angmom"}),
    "dof_pos_limits": ((), {"asset_cfg": all_joints}), "action_rate_l2": ((), {}),
    "air_time": (("threshold_min", "threshold_max", "command_threshold"), {**twist, **fgc}),
    "foot_clearance": (("target_height", "command_threshold"), {**twist, "asset_cfg": foot_sites}),
    "foot_swing_height": (("target_height", "command_threshold"), {**twist, **fgc, "asset_cfg": foot_sites}),
    "foot_slip": (("command_threshold",), {**twist, **fgc, "asset_cfg": foot_sites}),
    "soft_landing": (("command_threshold",), {**twist, **fgc}),
    "self_collisions": ((), {"sensor_name": "self_collision"}),
    "is_alive": ((), {}), "is_terminated": ((), {}), "action_acc_l2": ((), {}),
    "joint_torques_l2": ((), {"asset_cfg": all_joints}), "joint_vel_l2": ((), {"asset_cfg": all_joints}),
    "joint_acc_l2": ((), {"asset_cfg": all_joints}), "flat_orientation_l2": ((), {"asset_cfg": all_joints}),
    "electrical_power_cost": ((), {"asset_cfg": all_joints}),
  }
  for slot, params in P.items():
    _check_params("reward", by_func[REWARD_FUNCS[slot]][0], REWARD_FUNCS[slot], params, *rules[slot])
  for slot in ("joint_torques_l2", "joint_vel_l2", "joint_acc_l2", "flat_orientation_l2", "electrical_power_cost"):
    ac = (P.get(slot) or {}).get("asset_cfg") or {}
    if any(ac.get(k) not in (None, (".*",), [".*"], ".*") for k in ("joint_names", "body_names")) or ac.get("name", "robot") != "robot":
      raise UnsupportedTermError(f"reward '{slot}': the kernel term covers all actuated joints of 'robot' (asset_cfg {ac})")
  if "track_linear_velocity" in P: c.std_lin = float(P["track_linear_velocity"]["std"])        # noqa: E701
  if "track_angular_velocity" in P: c.std_ang = float(P["track_angular_velocity"]["std"])      # noqa: E701
  if "upright" in P:
    c.std_upright = float(P["upright"]["std"])
    bn = P["upright"]["asset_cfg"].get("body_names")
    if bn:
      body_local = [m.body_names[i].removeprefix(pre) for i in c.body_ids]
      hit = [i for i, n in enumerate(body_local) if any(re.fullmatch(k, n) for k in bn)]
      if len(hit) != 1:
        raise UnsupportedTermError(f"flat_orientation: body_names {bn} must select exactly one body")
      c.upright_body_local, c.upright_body = hit[0], int(c.body_ids[hit[0]])
  if "pose" in P:
    p = P["pose"]
    c.walking_threshold, c.running_threshold = float(p["walking_threshold"]), float(p["running_threshold"])
    for key in ("standing", "walking", "running"):
      mp = {k: float(v) for k, v in p[f"std_{key}"].items()}
      setattr(c, f"std_{key}_map", mp)
      setattr(c, f"std_{key}", _resolve_values(mp, local_names))
  if "air_time" in P:
    p = P["air_time"]
    c.air_thr_min, c.air_thr_max, c.air_cmd_threshold = float(p["threshold_min"]), float(p["threshold_max"]), float(p["command_threshold"])
  th = {float(P[s]["command_threshold"]) for s in ("foot_clearance", "foot_swing_height", "foot_slip", "soft_landing") if s in P}
  hs = {float(P[s]["target_height"]) for s in ("foot_clearance", "foot_swing_height") if s in P}
  if len(th) > 1 or len(hs) > 1:
    raise UnsupportedTermError(f"the foot terms share one command_threshold / target_height in the kernels, got {sorted(th)} / {sorted(hs)}")
  if th: c.cmd_threshold = th.pop()          # noqa: E701
  if hs: c.target_height = hs.pop()          # noqa: E701
  # ---- terminations -----------------------------------------------------------------------------------------------------
  tf = {t["func"]: t for t in d["terminations"].values() if t is not None}
  bad = set(tf) - {"time_out", "bad_orientation", "illegal_contact", "root_height_below_minimum", "nan_detection"}
  if bad:
    raise UnsupportedTermError(f"termination functions {sorted(bad)} are not implemented by the env kernels")
  if "time_out" not in tf or not tf["time_out"]["time_out"]:
    raise UnsupportedTermError("the kernels always apply the time_out truncation")
  trules = {"time_out": ((), {}), "bad_orientation": (("limit_angle",), {"asset_cfg": all_joints}),
            "illegal_contact": ((), {"sensor_name": "nonfoot_ground_touch"}),
            "root_height_below_minimum": (("minimum_height",), {"asset_cfg": all_joints}), "nan_detection": ((), {})}
  for f, t in tf.items():
    _check_params("termination", f, f, t["params"], *trules[f])
  if "bad_orientation" in tf:
    c.limit_angle = float(tf["bad_orientation"]["params"]["limit_angle"])
  else:
    c.limit_angle = 4.0          # never reached: projected gravity angle <= pi
  c.min_root_height = float(tf["root_height_below_minimum"]["params"]["minimum_height"]) if "root_height_below_minimum" in tf else None
  c.nan_termination = "nan_detection" in tf
  for f in ("bad_orientation", "root_height_below_minimum", "nan_detection", "illegal_contact"):
    if f in tf and tf[f]["time_out"]:
      raise UnsupportedTermError(f"termination '{f}' with time_out=True: the kernels count it as a failure termination")
  if ("illegal_contact" in tf) != (c.illegal_count > 0):
    raise UnsupportedTermError("illegal_contact needs the scene's nonfoot_ground_touch sensor (and the compiled Go1 scene always has it)")
  # ---- observations -------------------------------------------------------------------------------------------------------
  # The kernel writes a fixed term bank: the seven policy terms (with the policy group's noise) and a clean copy of them followed
  # by the four foot terms (the critic group).  Any group of the cfg - any name, any subset, any order - is a gather of bank
  # segments (`extra["obs_layout"]`, applied by obs_pipeline.ObservationPipeline together with clip / scale / delay / history);
  # a term with noise comes from the noisy copy, one without from the clean copy.
  obs = d["observations"]
  nj, nf = c.nj, c.nfeet
  bank_p = {f: (o, w) for f, o, w in zip(POLICY_OBS_TERMS, (0, 3, 6, 9, 9 + nj, 9 + 2 * nj, 9 + 3 * nj), (3, 3, 3, nj, nj, nj, 3))}
  bank_e = {f: (o, w) for f, o, w in zip(CRITIC_EXTRA_OBS_TERMS, (12 + 3 * nj, 12 + 3 * nj + nf, 12 + 3 * nj + 2 * nf, 12 + 3 * nj + 3 * nf),
                                         (nf, nf, nf, 3 * nf))}
  bank_x = {("base_lin_vel", None): (0, 3), ("base_ang_vel", None): (3, 3)}
  lo, hi = np.zeros(c.n_obs_policy), np.zeros(c.n_obs_policy)
  noise_of: dict = {}
  ranges_of: dict = {}
  post, layout = {}, {}
  for group, g in obs.items():
    if g is None:
      continue
    if not g["concatenate_terms"]:
      raise UnsupportedTermError("concatenate_terms=False")
    post[group], layout[group] = [], []
    for n, t in g["terms"].items():
      if t is None:
        continue
      key = (t["func"], t["params"].get("sensor_name") if t["func"] == "builtin_sensor" else None)
      orules = {"builtin_sensor": (("sensor_name",), {}), "projected_gravity": ((), {"asset_cfg": all_joints}),
                "joint_pos_rel": (("biased",), {"asset_cfg": all_joints}), "joint_vel_rel": ((), {"asset_cfg": all_joints}),
                "last_action": ((), {"action_name": lambda v: v in (None, "joint_pos")}), "generated_commands": ((), twist),
                "foot_height": ((), {"asset_cfg": foot_sites}), "foot_air_time": ((), fgc), "foot_contact": ((), fgc),
                "foot_contact_forces": ((), fgc), "base_lin_vel": ((), {"asset_cfg": all_joints}), "base_ang_vel": ((), {"asset_cfg": all_joints})}
      if t["func"] in orules:
        _check_params("observation term", n, t["func"], t["params"], *orules[t["func"]])
      if t["func"] == "joint_pos_rel" and t["params"].get("biased"):
        raise UnsupportedTermError(f"observation term '{n}': joint_pos_rel(biased=True) is implemented by the tracking kernels only")
      if key in bank_x:   # outside the shipped groups: the kernel writes them to a side buffer, clean only
        if t["noise"] is not None and g["enable_corruption"]:
          raise UnsupportedTermError(f"observation term '{n}': {key[0]} carries no noise in the kernels (use the imu sensors' terms)")
        o, w = bank_x[key]
        layout[group].append((n, w, "extra", o))
        e = {k: t[k] for k in _POST_KEYS}
        if g.get("history_length") is not None:
          e["history_length"], e["flatten_history_dim"] = g["history_length"], g["flatten_history_dim"]
        e["name"] = n
        post[group].append(e)
        continue
      if t["func"] == "builtin_sensor" and key not in bank_p:
        # any other builtin sensor of the scene (envs/mdp/observations.py:102-108): its slice of sensordata, as the physics kernels
        # wrote it (the accelerometer `robot/imu_lin_acc`, `robot/root_angmom`, ...); no kernel-side noise for these
        sn = [str(x) for x in m.sensor_names]
        if key[1] not in sn or int(m.sensor_type[sn.index(key[1])]) == 42:
          raise UnsupportedTermError(f"observation term '{n}': builtin sensor '{key[1]}' is not in the scene (builtin sensors: "
                                     f"{[x for x, ty in zip(sn, m.sensor_type) if int(ty) != 42]})")
        if t["noise"] is not None and g["enable_corruption"]:
          raise UnsupportedTermError(f"observation term '{n}': sensor '{key[1]}' carries no noise in the kernels (only the imu velocity sensors do)")
        si = sn.index(key[1])
        layout[group].append((n, int(m.sensor_dim[si]), "sensordata", int(m.sensor_adr[si])))
        e = {k: t[k] for k in _POST_KEYS}
        if g.get("history_length") is not None:
          e["history_length"], e["flatten_history_dim"] = g["history_length"], g["flatten_history_dim"]
        e["name"] = n
        post[group].append(e)
        continue
      if key not in bank_p and key not in bank_e:
        raise UnsupportedTermError(f"observation term '{n}' ({key[0]}{', ' + str(key[1]) if key[1] else ''}) is not in the env kernels' term bank "
                                   f"{[k[1] or k[0] for k in POLICY_OBS_TERMS + CRITIC_EXTRA_OBS_TERMS] + ['base_lin_vel', 'base_ang_vel']}")
      nz = t["noise"]
      if nz is not None:
        if nz.get("__type__") != "UniformNoiseCfg" or nz.get("operation") != "add":
          raise UnsupportedTermError(f"observation term '{n}': noise {nz} (only additive UniformNoiseCfg is implemented)")
        if key not in bank_p:
          raise UnsupportedTermError(f"observation term '{n}': the foot terms of the kernels' bank carry no noise")
        rng = (float(nz["n_min"]), float(nz["n_max"]))
        if g["enable_corruption"]:
          if noise_of.setdefault(key, rng) != rng:
            raise UnsupportedTermError(f"observation term '{n}': one noise range per term function ({noise_of[key]} vs {rng})")
          ranges_of[key] = rng
        else:
          ranges_of.setdefault(key, rng)     # kept in the table while no corrupted group says otherwise (the reference keeps the term's cfg too)
      if nz is not None and g["enable_corruption"]:
        o, w = bank_p[key]
        src = ("policy", o, w)
      else:
        o, w = bank_p[key] if key in bank_p else bank_e[key]
        src = ("critic", o, w)
      layout[group].append((n, w, src[0], src[1]))
      e = {k: t[k] for k in _POST_KEYS}
      if g.get("history_length") is not None:
        e["history_length"], e["flatten_history_dim"] = g["history_length"], g["flatten_history_dim"]
      e["name"] = n
      post[group].append(e)
  for key, rng in ranges_of.items():
    o, w = bank_p[key]
    lo[o:o + w], hi[o:o + w] = noise_of.get(key, rng)
  if "policy" not in layout or "critic" not in layout:
    raise UnsupportedTermError(f"observation groups {sorted(layout)}: 'policy' and 'critic' are required (rsl_rl's actor / critic inputs)")
  c.enable_corruption = bool(noise_of)
  c.noise_lo, c.noise_hi = lo, hi
  c.extra["obs_post"] = post
  c.extra["obs_layout"] = layout
  default = {"policy": [(o, w, "policy") for o, w in bank_p.values()],
             "critic": [(o, w, "critic") for o, w in list(bank_p.values()) + list(bank_e.values())]}
  got = {g: [(o, w, sg) for _, w, sg, o in terms] for g, terms in layout.items()}
  # the kernel's own two groups, as they are: nothing to gather (a clean-copy source is the same data when the term has no noise)
  c.extra["obs_custom"] = set(got) != {"policy", "critic"} or any(
    [(o, w, sg in ("extra", "sensordata")) for o, w, sg in got[g]] != [(o, w, False) for o, w, _ in default[g]] for g in ("policy", "critic"))
  # ---- actions ----------------------------------------------------------------------------------------------------------------
  acts = [a for a in d["actions"].values() if a is not None]
  if len(acts) != 1 or acts[0]["class_type"] != "JointPositionAction" or not acts[0]["use_default_offset"] or acts[0]["clip"]:
    raise UnsupportedTermError("actions: one JointPositionAction with use_default_offset=True and no clip is implemented")
  sc = acts[0]["scale"]
  c.action_scale = _resolve_values({k: float(v) for k, v in sc.items()}, local_names) if isinstance(sc, dict) else np.full(nj, float(sc))
  # ---- events -------------------------------------------------------------------------------------------------------------------
  c.push_enabled = False
  c.extra["foot_friction"] = None
  terrain_mode = 0
  for name, e in d["events"].items():
    if e is None:
      continue
    f, p = e["func"], e["params"]
    if f in ("reset_root_state_uniform", "reset_joints_by_offset", "push_by_setting_velocity"):
      _check_params("event", name, f, p, ("pose_range", "velocity_range", "position_range"), {"asset_cfg": all_joints})
    if f == "reset_root_state_uniform" and e["mode"] == "reset":
      c.reset_pose_range, c.reset_velocity_range = _range6(p.get("pose_range")), _range6(p.get("velocity_range"))
    elif f == "reset_joints_by_offset" and e["mode"] == "reset":
      c.reset_joint_pos, c.reset_joint_vel = tuple(map(float, p["position_range"])), tuple(map(float, p["velocity_range"]))
    elif f == "push_by_setting_velocity" and e["mode"] == "interval":
      c.push_enabled, c.push_interval = True, tuple(map(float, e["interval_range_s"]))
      c.push_velocity_range = _range6(p["velocity_range"])
    elif f == "randomize_field" and e["mode"] == "startup":
      # applied once by the env after construction through events.randomize_field (any field of its FIELD_SPECS), in cfg order
      rg = p["ranges"]
      rg = {int(k): tuple(map(float, v)) for k, v in rg.items()} if isinstance(rg, dict) else tuple(map(float, rg))
      ac = p.get("asset_cfg") or {"name": "robot"}
      c.extra.setdefault("startup_dr", []).append(dict(
        name=name, field=p["field"], ranges=rg, distribution=p.get("distribution", "uniform"), operation=p.get("operation", "abs"),
        axes=list(p["axes"]) if p.get("axes") is not None else None,
        asset_cfg={k: (tuple(ac[k]) if isinstance(ac.get(k), (list, tuple)) else ac.get(k)) for k in ("name", "geom_names", "body_names", "joint_names", "site_names")}))
      if p["field"] == "geom_friction" and p.get("operation", "abs") == "abs" and not isinstance(rg, dict):
        c.extra["foot_friction"] = dict(ranges=rg, geom_names=tuple(ac["geom_names"]))   # (kept for the task-table tests)
    elif f in ("randomize_pd_gains", "randomize_effort_limits", "randomize_encoder_bias") and e["mode"] == "startup":
      c.extra.setdefault("startup_dr", []).append(_startup_actuator_event(name, f, p))
    elif f == "randomize_terrain" and e["mode"] == "reset":
      terrain_mode |= 2
    else:
      raise UnsupportedTermError(f"event '{name}' ({f}, mode {e['mode']}) is not implemented by the env kernels")
  # ---- commands -------------------------------------------------------------------------------------------------------------------
  cmds = [t for t in d["commands"].values() if t is not None]
  if len(cmds) != 1 or cmds[0]["class_type"] != "UniformVelocityCommand":
    raise UnsupportedTermError("commands: one UniformVelocityCommand is implemented")
  t = cmds[0]
  if t["init_velocity_prob"] != 0.0:
    raise UnsupportedTermError("UniformVelocityCommand.init_velocity_prob > 0")
  r = t["ranges"]
  c.cmd_resample = tuple(map(float, t["resampling_time_range"]))
  c.cmd_lin_x, c.cmd_lin_y, c.cmd_ang_z = tuple(map(float, r["lin_vel_x"])), tuple(map(float, r["lin_vel_y"])), tuple(map(float, r["ang_vel_z"]))
  c.heading_command = bool(t["heading_command"])
  if c.heading_command:
    c.cmd_heading = tuple(map(float, r["heading"]))
  c.rel_standing, c.rel_heading, c.heading_stiffness = float(t["rel_standing_envs"]), float(t["rel_heading_envs"]), float(t["heading_control_stiffness"])
  # ---- curriculum -------------------------------------------------------------------------------------------------------------------
  c.extra["velocity_stages"] = ()
  c.extra["reward_weight_stages"] = []   # [(kernel reward slot, ((step, weight), ...))], tasks/velocity/mdp/curriculums.py:95-107
  slot_of_term = {by_func[f][0]: k for k, s_ in enumerate(REWARD_NAMES) for f in (REWARD_FUNCS[s_],) if f in by_func}
  for name, t in d["curriculum"].items():
    if t is None:
      continue
    if t["func"] == "commands_vel":
      c.extra["velocity_stages"] = tuple({k: (tuple(v) if isinstance(v, list) else v) for k, v in st.items()} for st in t["params"]["velocity_stages"])
    elif t["func"] == "terrain_levels_vel":
      terrain_mode |= 1
    elif t["func"] == "reward_weight":
      rn = t["params"]["reward_name"]
      if rn not in slot_of_term:
        raise UnsupportedTermError(f"curriculum term '{name}' (reward_weight): no reward term named '{rn}'")
      c.extra["reward_weight_stages"].append((slot_of_term[rn], tuple((int(s_["step"]), float(s_["weight"])) for s_ in t["params"]["weight_stages"])))
    else:
      raise UnsupportedTermError(f"curriculum term '{name}' ({t['func']})")
  has_terrain = c.terrain_origins is not None
  gen = d["scene"]["terrain"].get("terrain_generator") if d["scene"].get("terrain") else None
  if (gen is not None) != has_terrain:
    raise UnsupportedTermError("the configuration's terrain (plane / generator) does not match the compiled scene")
  c.terrain_mode = terrain_mode if has_terrain else 0
  if gen is not None:
    c.terrain_half_size = float(gen["size"][0]) / 2.0
    c.max_init_terrain_level = int(d["scene"]["terrain"]["max_init_terrain_level"])
  return c


# ---------------------------------------------------------------------------------------------------------
# the tracking task from a reference ManagerBasedRlEnvCfg (tasks/tracking/tracking_env_cfg.py)
# ---------------------------------------------------------------------------------------------------------
TRACKING_REWARD_FUNCS = {
  "motion_global_root_pos": "motion_global_anchor_position_error_exp", "motion_global_root_ori": "motion_global_anchor_orientation_error_exp",
  "motion_body_pos": "motion_relative_body_position_error_exp", "motion_body_ori": "motion_relative_body_orientation_error_exp",
  "motion_body_lin_vel": "motion_global_body_linear_velocity_error_exp", "motion_body_ang_vel": "motion_global_body_angular_velocity_error_exp",
  "action_rate_l2": "action_rate_l2", "joint_limit": "joint_pos_limits", "self_collisions": "self_collision_cost",
}
_TRACKING_STD = {"motion_global_root_pos": "std_anchor_pos", "motion_global_root_ori": "std_anchor_ori", "motion_body_pos": "std_body_pos",
                 "motion_body_ori": "std_body_ori", "motion_body_lin_vel": "std_body_lin_vel", "motion_body_ang_vel": "std_body_ang_vel"}
_TRACKING_POLICY_OBS = (("generated_commands", None), ("motion_anchor_pos_b", None), ("motion_anchor_ori_b", None), ("builtin_sensor", "robot/imu_lin_vel"),
                        ("builtin_sensor", "robot/imu_ang_vel"), ("joint_pos_rel", None), ("joint_vel_rel", None), ("last_action", None))
_TRACKING_CRITIC_OBS = _TRACKING_POLICY_OBS[:3] + (("robot_body_pos_b", None), ("robot_body_ori_b", None)) + _TRACKING_POLICY_OBS[3:]


TRACKING_POLICY_TERM_NAMES = ("command", "motion_anchor_pos_b", "motion_anchor_ori_b", "base_lin_vel", "base_ang_vel", "joint_pos", "joint_vel", "actions")


def tracking_without_state_estimation(c: TrackingCfg, drop=("motion_anchor_pos_b", "base_lin_vel")) -> TrackingCfg:
  """The built-in tracking table as the registered "...-No-State-Estimation" task (tasks/tracking/config/g1/env_cfgs.py:73-84): the
  policy group loses `drop`; what remains is gathered from the kernels' term bank exactly as `tracking_from_manager_cfg` does for
  the task's own cfg (noisy terms from obs_policy, noise-free ones from the clean copy, joint_pos(biased) from obs_policy)."""
  nj, nb = c.nj, int(c.nbody)
  widths_p = (2 * nj, 3, 6, 3, 3, nj, nj, nj)
  off = np.concatenate([[0], np.cumsum(widths_p)])
  plain_post = {k: (None if k in ("clip", "scale") else 0) for k in _POST_KEYS}
  plain_post.update(delay_per_env=True, delay_per_env_phase=True, flatten_history_dim=True, delay_hold_prob=0.0)
  pol = []
  for i, (name, w) in enumerate(zip(TRACKING_POLICY_TERM_NAMES, widths_p)):
    o = int(off[i])
    if name in drop:
      c.noise_lo[o:o + w] = 0
      c.noise_hi[o:o + w] = 0
      continue
    noisy = c.enable_corruption and bool(np.any(c.noise_lo[o:o + w] != 0) or np.any(c.noise_hi[o:o + w] != 0))
    if noisy or name == "joint_pos":
      pol.append((name, w, "policy", o))
    else:
      pol.append((name, w, "critic", o + (9 * nb if i >= 3 else 0)))
  cri_names = TRACKING_POLICY_TERM_NAMES[:3] + ("body_pos", "body_ori") + TRACKING_POLICY_TERM_NAMES[3:]
  widths_c = widths_p[:3] + (3 * nb, 6 * nb) + widths_p[3:]
  offc = np.concatenate([[0], np.cumsum(widths_c)])
  cri = [(n, w, "critic", int(offc[i])) for i, (n, w) in enumerate(zip(cri_names, widths_c))]
  c.extra = dict(getattr(c, "extra", None) or {})
  c.extra["obs_layout"] = {"policy": pol, "critic": cri}
  c.extra["obs_post"] = {g: [dict(plain_post, name=t[0]) for t in ts] for g, ts in c.extra["obs_layout"].items()}
  c.extra["obs_custom"] = True
  return c


def tracking_from_manager_cfg(cfg, m, time_step_total: int, robot: str = "g1") -> TrackingCfg:
  """Resolve a reference `ManagerBasedRlEnvCfg` of the tracking task (live dataclass or `cfg_to_dict` view) into the tracking
  kernels' `TrackingCfg`: reward weights / stds by term FUNCTION, termination thresholds and end-effector bodies, the
  MotionCommand's bodies, anchor, reset ranges and sampling parameters, observation noise, action scale, push event, episode
  length.  Startup events come back in `extra["startup_dr"]` / `extra["encoder_bias_range"]` for the env to apply.  The
  observation layout is the kernel's fixed one; anything the kernels do not implement raises `UnsupportedTermError`."""
  d = cfg if isinstance(cfg, dict) else cfg_to_dict(cfg)
  cmds = [t for t in d["commands"].values() if t is not None]
  if len(cmds) != 1 or cmds[0]["class_type"] != "MotionCommand":
    raise UnsupportedTermError("commands: one MotionCommand is implemented")
  mc = cmds[0]
  ee_t = [t for t in d["terminations"].values() if t is not None and t["func"] in ("bad_motion_body_pos_z_only", "bad_motion_body_pos")]
  ee = tuple(ee_t[0]["params"]["body_names"]) if ee_t else G1_EE_BODIES
  c = tracking_cfg(m, time_step_total, robot, play=False, tracked=tuple(mc["body_names"]), anchor=mc["anchor_body_name"], ee=ee)
  pre = "robot/"
  local_names = [n.removeprefix(pre) for n in velocity_cfg(m, robot).joint_names]
  mj = d["sim"]["mujoco"]
  if mj["solver"] not in ("newton", "cg") or mj["cone"] not in ("pyramidal", "elliptic") or mj["integrator"] not in ("implicitfast", "euler"):
    raise UnsupportedTermError(f"sim options {mj['solver']} / {mj['cone']} / {mj['integrator']} are outside the supported subset")
  c.step_dt = float(mj["timestep"]) * int(d["decimation"])
  c.max_episode_length_s = float(d["episode_length_s"])
  c.max_episode_length = math.ceil(c.max_episode_length_s / c.step_dt)
  c.bin_count = int(time_step_total // (1 / c.step_dt)) + 1
  c.extra = {"decimation": int(d["decimation"]),
             "sim": dict(timestep=float(mj["timestep"]), iterations=int(mj["iterations"]), ls_iterations=int(mj["ls_iterations"]),
                         integrator=mj["integrator"], nconmax=d["sim"]["nconmax"], njmax=d["sim"]["njmax"])}
  # ---- rewards ----
  by_func = {}
  for name, t in d["rewards"].items():
    if t is None:
      continue
    if t["func"] in by_func:
      raise UnsupportedTermError(f"reward function '{t['func']}' appears twice")
    by_func[t["func"]] = (name, t)
  unknown = set(by_func) - set(TRACKING_REWARD_FUNCS.values())
  if unknown:
    raise UnsupportedTermError(f"reward terms {sorted(by_func[f][0] for f in unknown)} ({sorted(unknown)}) are not implemented by the tracking kernels")
  w = np.zeros(len(TRACKING_REWARD_NAMES))

  def all_joints(v):
    return _all_joints(v)
  all_joints.__doc__ = "all actuated joints of 'robot'"
  motion = {"command_name": "motion"}
  for k, slot in enumerate(TRACKING_REWARD_NAMES):
    if TRACKING_REWARD_FUNCS[slot] in by_func:
      rname, t = by_func[TRACKING_REWARD_FUNCS[slot]]
      _check_params("reward", rname, t["func"], t["params"], ("std",), {**motion, "asset_cfg": all_joints, "sensor_name": "self_collision"})
      w[k] = float(t["weight"])
      if slot in _TRACKING_STD:
        setattr(c, _TRACKING_STD[slot], float(t["params"]["std"]))
  if w[TRACKING_REWARD_NAMES.index("self_collisions")] != 0 and c.selfcol_adr < 0:
    raise UnsupportedTermError("self_collision_cost needs the scene's self_collision sensor")
  c.reward_weights = w
  # ---- terminations ----
  tf = {t["func"]: t for t in d["terminations"].values() if t is not None}
  bad = set(tf) - {"time_out", "bad_anchor_pos_z_only", "bad_anchor_pos", "bad_anchor_ori", "bad_motion_body_pos_z_only", "bad_motion_body_pos"}
  if bad:
    raise UnsupportedTermError(f"termination functions {sorted(bad)} are not implemented by the tracking kernels")
  for a_, b_ in (("bad_anchor_pos", "bad_anchor_pos_z_only"), ("bad_motion_body_pos", "bad_motion_body_pos_z_only")):
    if a_ in tf and b_ in tf:
      raise UnsupportedTermError(f"terminations {a_} and {b_} share one kernel slot (3-D distance or height difference)")
  if "time_out" not in tf or not tf["time_out"]["time_out"]:
    raise UnsupportedTermError("the kernels always apply the time_out truncation")
  for f, t in tf.items():
    _check_params("termination", f, f, t["params"], ("threshold", "body_names"), {**motion, "asset_cfg": all_joints})
  BIG = 1e9   # a threshold no error reaches = term absent
  apf = "bad_anchor_pos" if "bad_anchor_pos" in tf else "bad_anchor_pos_z_only"
  eef = "bad_motion_body_pos" if "bad_motion_body_pos" in tf else "bad_motion_body_pos_z_only"
  c.term_full3d = (1 if "bad_anchor_pos" in tf else 0) | (2 if "bad_motion_body_pos" in tf else 0)
  c.anchor_pos_threshold = float(tf[apf]["params"]["threshold"]) if apf in tf else BIG
  c.anchor_ori_threshold = float(tf["bad_anchor_ori"]["params"]["threshold"]) if "bad_anchor_ori" in tf else BIG
  c.ee_pos_threshold = float(tf[eef]["params"]["threshold"]) if eef in tf else BIG
  # ---- command ----
  r6 = VelocityCfg._ranges6
  c.pose_range, c.velocity_range = r6(_range6(mc.get("pose_range"))), r6(_range6(mc.get("velocity_range")))
  c.joint_position_range = tuple(map(float, mc["joint_position_range"]))
  c.sampling_mode = mc["sampling_mode"]
  c.adaptive_kernel_size, c.adaptive_lambda = int(mc["adaptive_kernel_size"]), float(mc["adaptive_lambda"])
  c.adaptive_uniform_ratio, c.adaptive_alpha = float(mc["adaptive_uniform_ratio"]), float(mc["adaptive_alpha"])
  # ---- observations ----
  # The kernel writes a fixed term bank: the eight policy terms (joint_pos_rel BIASED, each term with the policy group's noise) into
  # obs_policy and a clean copy of them (joint_pos_rel unbiased) plus robot_body_pos_b / robot_body_ori_b into obs_critic.  The
  # shipped layout is returned as it is; any other set of groups - a subset (the registered "...-No-State-Estimation" task drops
  # motion_anchor_pos_b and base_lin_vel from the policy, config/g1/env_cfgs.py:73-84), another order, further groups, clip /
  # scale / delay / history - is a gather of bank segments by obs_pipeline.ObservationPipeline (`extra["obs_layout"]`).
  obs = d["observations"]
  groups = {k: g for k, g in obs.items() if g is not None}
  if not {"policy", "critic"} <= set(groups):
    raise UnsupportedTermError(f"tracking observation groups {sorted(groups)}: 'policy' and 'critic' are required (rsl_rl's actor / critic inputs)")
  nj, nb = c.nj, int(c.nbody)
  widths_p = (2 * nj, 3, 6, 3, 3, nj, nj, nj)
  off = np.concatenate([[0], np.cumsum(widths_p)])
  bank_p = {k: (int(off[i]), w) for i, (k, w) in enumerate(zip(_TRACKING_POLICY_OBS, widths_p))}
  widths_c = widths_p[:3] + (3 * nb, 6 * nb) + widths_p[3:]
  offc = np.concatenate([[0], np.cumsum(widths_c)])
  bank_c = {k: (int(offc[i]), w) for i, (k, w) in enumerate(zip(_TRACKING_CRITIC_OBS, widths_c))}
  lo, hi = np.zeros(c.n_obs_policy), np.zeros(c.n_obs_policy)
  def _key(t):
    return (t["func"], t["params"].get("sensor_name") if t["func"] == "builtin_sensor" else None)

  def _noise(g, t):
    return t["noise"] if g["enable_corruption"] else None

  # pass 1: which bank terms carry noise (the kernel corrupts a segment of obs_policy for everyone who reads it)
  noise_of: dict = {}
  kept: dict = {}     # ranges of terms whose group is not corrupted: kept in the table (the reference keeps the term's cfg too), not applied
  for group, g in groups.items():
    for n, t in g["terms"].items():
      nz = t["noise"] if t is not None else None
      if nz is None:
        continue
      if nz.get("__type__") != "UniformNoiseCfg" or nz.get("operation") != "add":
        raise UnsupportedTermError(f"observation term '{n}': noise {nz} (only additive UniformNoiseCfg is implemented)")
      key, rng = _key(t), (float(nz["n_min"]), float(nz["n_max"]))
      if not g["enable_corruption"]:
        if key in bank_p:
          kept.setdefault(key, rng)
        continue
      if key not in bank_c:
        continue      # named as unknown by the second pass
      if key not in bank_p:
        raise UnsupportedTermError(f"tracking observation term '{n}': {key[0]} carries no noise in the kernels")
      if noise_of.setdefault(key, rng) != rng:
        raise UnsupportedTermError(f"observation term '{n}': one noise range per term function ({noise_of[key]} vs {rng})")
  for key, rng in (noise_of if noise_of else kept).items():   # with any corrupted group the kernel's noise is live: only asked-for ranges
    o, w = bank_p[key]
    lo[o:o + w], hi[o:o + w] = rng
  # pass 2: every term as a bank segment - noisy terms from obs_policy, clean ones from the clean copy in obs_critic; joint_pos_rel exists
  # as (biased, obs_policy) and (unbiased, obs_critic)
  layout, post, shape = {}, {}, {}
  for group, g in groups.items():
    if not g["concatenate_terms"]:
      raise UnsupportedTermError("concatenate_terms=False")
    layout[group], post[group], shape[group] = [], [], []
    for n, t in g["terms"].items():
      if t is None:
        continue
      key = _key(t)
      if key not in bank_c:
        raise UnsupportedTermError(f"tracking observation term '{n}' ({key[0]}{', ' + str(key[1]) if key[1] else ''}) is not in the tracking kernels' "
                                   f"term bank {[k[1] or k[0] for k in _TRACKING_CRITIC_OBS]}")
      _check_params("observation term", n, t["func"], t["params"], ("sensor_name", "biased"), {**motion, "asset_cfg": all_joints})
      biased = bool(t["params"].get("biased")) if t["func"] == "joint_pos_rel" else False
      noisy = _noise(g, t) is not None
      if t["func"] == "joint_pos_rel" and not biased and noisy:
        raise UnsupportedTermError(f"tracking observation term '{n}': the kernels corrupt joint_pos_rel(biased=True) only (the encoder reading of the policy group)")
      if t["func"] == "joint_pos_rel" and biased and not noisy and key in noise_of:
        raise UnsupportedTermError(f"tracking observation term '{n}': joint_pos_rel(biased=True) without noise next to a group that corrupts it")
      if noisy or biased:
        o, w = bank_p[key]
        layout[group].append((n, w, "policy", o))
      else:
        o, w = bank_c[key]
        layout[group].append((n, w, "critic", o))
      shape[group].append((key, biased, noisy or key not in noise_of))    # last: obs_policy's copy of the term is what this term asks for
      e = {k: t[k] for k in _POST_KEYS}
      if g.get("history_length") is not None:
        e["history_length"], e["flatten_history_dim"] = g["history_length"], g["flatten_history_dim"]
      e["name"] = n
      post[group].append(e)
  c.enable_corruption = bool(noise_of)
  c.noise_lo, c.noise_hi = lo, hi
  # the shipped layout (train or play): the kernel's two buffers are the two groups as they are, nothing to gather
  plain = (set(layout) == {"policy", "critic"}
           and shape["policy"] == [(k, k[0] == "joint_pos_rel", True) for k in _TRACKING_POLICY_OBS]
           and [(k, b) for k, b, _ in shape["critic"]] == [(k, False) for k in _TRACKING_CRITIC_OBS]
           and all(lay[2] == "critic" for lay in layout["critic"]))
  needs_post = any(t["clip"] or t["scale"] is not None or t["delay_max_lag"] > 0 or t["history_length"] > 0 for ts in post.values() for t in ts)
  c.extra = dict(getattr(c, "extra", {}) or {})
  c.extra["obs_layout"], c.extra["obs_post"], c.extra["obs_custom"] = layout, post, (not plain) or needs_post
  # ---- actions ----
  acts = [a for a in d["actions"].values() if a is not None]
  if len(acts) != 1 or acts[0]["class_type"] != "JointPositionAction" or not acts[0]["use_default_offset"] or acts[0]["clip"]:
    raise UnsupportedTermError("actions: one JointPositionAction with use_default_offset=True and no clip is implemented")
  sc = acts[0]["scale"]
  c.action_scale = _resolve_values({k: float(v) for k, v in sc.items()}, local_names) if isinstance(sc, dict) else np.full(nj, float(sc))
  # ---- events ----
  c.push_enabled = False
  c.extra["startup_dr"], c.extra["encoder_bias_range"] = [], None
  for name, e in d["events"].items():
    if e is None:
      continue
    f, p = e["func"], e["params"]
    if f == "push_by_setting_velocity" and e["mode"] == "interval":
      c.push_enabled, c.push_interval = True, tuple(map(float, e["interval_range_s"]))
      c.push_vel = r6(_range6(p["velocity_range"]))
    elif f == "randomize_field" and e["mode"] == "startup":
      rg = p["ranges"]
      rg = {int(k): tuple(map(float, v)) for k, v in rg.items()} if isinstance(rg, dict) else tuple(map(float, rg))
      ac = p.get("asset_cfg") or {"name": "robot"}
      c.extra["startup_dr"].append(dict(
        name=name, field=p["field"], ranges=rg, distribution=p.get("distribution", "uniform"), operation=p.get("operation", "abs"),
        axes=list(p["axes"]) if p.get("axes") is not None else None,
        asset_cfg={k: (tuple(ac[k]) if isinstance(ac.get(k), (list, tuple)) else ac.get(k)) for k in ("name", "geom_names", "body_names", "joint_names", "site_names")}))
    elif f == "randomize_encoder_bias" and e["mode"] == "startup":
      c.extra["encoder_bias_range"] = tuple(map(float, p["bias_range"]))
    elif f in ("randomize_pd_gains", "randomize_effort_limits") and e["mode"] == "startup":
      c.extra["startup_dr"].append(_startup_actuator_event(name, f, p))
    else:
      raise UnsupportedTermError(f"event '{name}' ({f}, mode {e['mode']}) is not implemented by the tracking kernels")
  if any(t is not None for t in (d.get("curriculum") or {}).values()):
    raise UnsupportedTermError("tracking curriculum terms are not implemented")
  return c
