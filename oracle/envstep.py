"""CPU oracle of the manager-based env step of the velocity task (TEST INFRASTRUCTURE, never on the product path).

A numpy restatement, batched over envs, of what mjlab's managers compute around the physics step.  Every function
cites the reference lines it follows (paths relative to /root/reference/src/mjlab).  It is PINNED by
tests/golden/mdp_velocity_{g1,go1}.npz, which hold the outputs of the reference's OWN functions (imported with the
native dependencies stubbed, tools/make_mdp_golden.py) for rewards, terminations, observations, the contact-sensor
air-time state machine, the swing-height state and the heading controller (tests/test_env_oracle.py).  The parts the
golden file does not cover (reset events, command resampling, push events, observation noise, logging sums) follow the
cited lines and use the same counter-based Philox4x32-10 draws as the CUDA kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""

from __future__ import annotations

import math
from types import SimpleNamespace as NS

import numpy as np

from mjlab_b200 import envabi as A

K = A.K
F = np.float32


# ---------------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) - identical to b2_philox in mjlab_b200/csrc/b2_env.cuh
# ---------------------------------------------------------------------------------------------------------
def philox4x32(k0, k1, c0, c1, c2, c3):
  c0, c1, c2, c3 = (np.asarray(c, np.uint64) & np.uint64(0xFFFFFFFF) for c in (c0, c1, c2, c3))
  c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
  k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
  M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
  for _ in range(10):
    p0, p1 = M0 * c0, M1 * c2
    c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & MASK, p0 & MASK
    k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
  return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def draws(seed: int, nworld: int, step: int) -> np.ndarray:
  """[nworld, B2E_NDRAW] uniforms in [0, 1): draw `slot` of env w at env-step `step` (env_draw in b2_env.cuh)."""
  nblk = (A.NDRAW + 3) // 4
  w = np.arange(nworld, dtype=np.uint64)[:, None]
  blk = np.arange(nblk, dtype=np.uint64)[None, :]
  o = philox4x32(seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, w, step & 0xFFFFFFFF, blk, (step >> 32) & 0xFFFFFFFF)
  u = (o >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)
  return u.reshape(nworld, nblk * 4)[:, : A.NDRAW]


# ---------------------------------------------------------------------------------------------------------
# small math (utils/lab_api/math.py)
# ---------------------------------------------------------------------------------------------------------
def quat_apply(q, v):
  """utils/lab_api/math.py:629-648"""
  xyz = q[..., 1:]
  t = np.cross(xyz, v) * 2
  return v + q[..., 0:1] * t + np.cross(xyz, t)


def quat_apply_inverse(q, v):
  """utils/lab_api/math.py:651-670"""
  xyz = q[..., 1:]
  t = np.cross(xyz, v) * 2
  return v - q[..., 0:1] * t + np.cross(xyz, t)


def quat_mul(a, b):
  """utils/lab_api/math.py:527-563"""
  w1, x1, y1, z1 = (a[..., i] for i in range(4))
  w2, x2, y2, z2 = (b[..., i] for i in range(4))
  return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                   w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)


def quat_from_euler_xyz(roll, pitch, yaw):
  """utils/lab_api/math.py:275-302"""
  cy, sy, cr, sr, cp, sp = np.cos(yaw * 0.5), np.sin(yaw * 0.5), np.cos(roll * 0.5), np.sin(roll * 0.5), np.cos(pitch * 0.5), np.sin(pitch * 0.5)
  return np.stack([cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp], -1)


def wrap_to_pi(a):
  """utils/lab_api/math.py:102-124"""
  w = np.mod(a + F(np.pi), F(2 * np.pi))
  return np.where((w == 0) & (a > 0), F(np.pi), w - F(np.pi)).astype(a.dtype)


def velocity_from_cvel(pos, subtree_com, cvel):
  """compute_velocity_from_cvel, entity/data.py:20-31 -> (lin_w, ang_w)"""
  ang = cvel[..., 0:3]
  return cvel[..., 3:6] - np.cross(ang, subtree_com - pos), ang


# ---------------------------------------------------------------------------------------------------------
# state
# ---------------------------------------------------------------------------------------------------------
def make_state(cfg, nworld: int, env_origins=None) -> NS:
  nj, nf = cfg.nj, cfg.nfeet
  z = lambda *s: np.zeros((nworld,) + s, F)  # noqa: E731
  return NS(
    env_origins=z(3) if env_origins is None else np.asarray(env_origins, F),
    terrain_level=np.zeros(nworld, np.int32), terrain_type=np.zeros(nworld, np.int32),
    action=z(nj), prev_action=z(nj), prev_prev_action=z(nj), raw_action=z(nj), processed_action=z(nj), encoder_bias=z(nj),
    joint_pos_target=z(nj), command=z(3), heading_target=z(), heading_error=z(), cmd_time_left=z(), cmd_metrics=z(2),
    cmd_counter=np.zeros(nworld, np.int32), is_heading=np.zeros(nworld, bool), is_standing=np.zeros(nworld, bool),
    air_state=z(4, nf), air_last_time=z(), peak_heights=z(nf), push_time_left=z(), episode_length=np.zeros(nworld, np.int64),
    reward=z(), step_reward=z(A.NREW), episode_sums=z(A.NREW), term_dones=np.zeros((nworld, A.NTERM), bool),
    terminated=np.zeros(nworld, bool), time_outs=np.zeros(nworld, bool), reset_buf=np.zeros(nworld, bool),
    obs_policy=z(cfg.n_obs_policy), obs_critic=z(cfg.n_obs_critic), obs_extra=z(6), log_env=z(A.NLOG), log=np.zeros(A.NLOG, F), step_counter=0,
  )


def root_kinematics(cfg, d) -> NS:
  """EntityData root properties (entity/data.py:212-229,498-533)."""
  rb = cfg.root_body
  quat = d.xquat[:, rb].astype(F)
  lin_w, ang_w = velocity_from_cvel(d.xpos[:, rb].astype(F), d.subtree_com[:, rb].astype(F), d.cvel[:, rb].astype(F))
  g = np.array([0, 0, -1], F)
  return NS(quat=quat, lin_w=lin_w, ang_w=ang_w, lin_b=quat_apply_inverse(quat, lin_w), ang_b=quat_apply_inverse(quat, ang_w),
            pg=quat_apply_inverse(quat, np.broadcast_to(g, lin_w.shape)),
            heading=None)


def heading_w(quat):
  """entity/data.py:504-508"""
  fw = quat_apply(quat, np.broadcast_to(np.array([1, 0, 0], F), quat.shape[:-1] + (3,)))
  return np.arctan2(fw[:, 1], fw[:, 0])


# ---------------------------------------------------------------------------------------------------------
# action path (managers/action_manager.py:113-130, envs/mdp/actions/joint_actions.py:84-104, actuator/builtin_group.py:68-81)
# ---------------------------------------------------------------------------------------------------------
def apply_action(cfg, s, d, action_in) -> None:
  a = np.asarray(action_in, F)
  s.prev_prev_action[:] = s.prev_action
  s.prev_action[:] = s.action
  s.action[:] = a
  s.raw_action[:] = a
  s.processed_action[:] = a * cfg.action_scale.astype(F) + cfg.default_joint_pos.astype(F)
  s.joint_pos_target[:] = s.processed_action - s.encoder_bias
  d.ctrl[:, cfg.ctrl_ids] = s.joint_pos_target


# ---------------------------------------------------------------------------------------------------------
# contact-sensor air time (sensor/contact_sensor.py:327-367) and first contact (:260-269)
# ---------------------------------------------------------------------------------------------------------
def update_air_time(cfg, s, sensordata, time) -> None:
  found = sensordata[:, cfg.foot_found_adr].astype(F)
  now = np.asarray(time, F)
  elapsed = (now - s.air_last_time)[:, None]
  contact = found > 0
  cur_air, last_air, cur_con, last_con = (s.air_state[:, i] for i in range(4))
  first_contact = (cur_air > 0) & contact
  first_detached = (cur_con > 0) & ~contact
  s.air_state[:, 1] = np.where(first_contact, cur_air + elapsed, last_air)
  s.air_state[:, 0] = np.where(~contact, cur_air + elapsed, 0)
  s.air_state[:, 3] = np.where(first_detached, cur_con + elapsed, last_con)
  s.air_state[:, 2] = np.where(contact, cur_con + elapsed, 0)
  s.air_last_time[:] = now


def first_contact(cfg, s):
  cur_con = s.air_state[:, 2]
  return (cur_con > 0) & (cur_con < F(cfg.step_dt) + F(1e-8))


# ---------------------------------------------------------------------------------------------------------
# terminations (envs/mdp/terminations.py:19-32, tasks/velocity/mdp/terminations.py:13-16)
# ---------------------------------------------------------------------------------------------------------
def terminations(cfg, s, d, R):
  time_out = s.episode_length >= cfg.max_episode_length
  fell = np.abs(np.arccos(-R.pg[:, 2])) > F(cfg.limit_angle)
  if cfg.illegal_count > 0:
    illegal = np.any(d.sensordata[:, cfg.illegal_adr: cfg.illegal_adr + cfg.illegal_count] != 0, axis=1)
  else:
    illegal = np.zeros_like(fell)
  # generic library (envs/mdp/terminations.py:35-47), off unless configured
  if getattr(cfg, "min_root_height", None) is not None:
    root_low = d.xpos[:, cfg.root_body, 2].astype(F) < F(cfg.min_root_height)
  else:
    root_low = np.zeros_like(fell)
  if getattr(cfg, "nan_termination", False):
    has_nan = np.zeros_like(fell)
    for a in (d.qpos, d.qvel, d.qacc, d.qacc_warmstart):
      has_nan |= ~np.all(np.isfinite(a), axis=1)
  else:
    has_nan = np.zeros_like(fell)
  return time_out, fell, illegal, root_low, has_nan


# ---------------------------------------------------------------------------------------------------------
# rewards (tasks/velocity/mdp/rewards.py:23-359, envs/mdp/rewards.py:56-88); returns [B, NREW] raw term values
# ---------------------------------------------------------------------------------------------------------
def reward_terms(cfg, s, d, R, log=None, terminated=None):
  B = s.command.shape[0]
  w = cfg.reward_weights
  cmd = s.command
  total_cmd = np.sqrt(cmd[:, 0] ** 2 + cmd[:, 1] ** 2) + np.abs(cmd[:, 2])
  sd = d.sensordata.astype(F)
  out = np.zeros((B, A.NREW), F)
  # track_linear_velocity :23-40 / track_angular_velocity :43-60
  xy = np.sum(np.square(cmd[:, :2] - R.lin_b[:, :2]), 1)
  out[:, K["B2E_REW_TRACK_LIN"]] = np.exp(-(xy + np.square(R.lin_b[:, 2])) / F(cfg.std_lin) ** 2)
  out[:, K["B2E_REW_TRACK_ANG"]] = np.exp(-(np.square(cmd[:, 2] - R.ang_b[:, 2]) + np.sum(np.square(R.ang_b[:, :2]), 1)) / F(cfg.std_ang) ** 2)
  # flat_orientation with body_ids :63-85, body_angular_velocity_penalty :98-107
  ub = cfg.upright_body
  pgu = quat_apply_inverse(d.xquat[:, ub].astype(F), np.broadcast_to(np.array([0, 0, -1], F), (B, 3)))
  out[:, K["B2E_REW_UPRIGHT"]] = np.exp(-np.sum(np.square(pgu[:, :2]), 1) / F(cfg.std_upright) ** 2)
  _, ang_u = velocity_from_cvel(d.xpos[:, ub].astype(F), d.subtree_com[:, cfg.root_body].astype(F), d.cvel[:, ub].astype(F))
  out[:, K["B2E_REW_BODY_ANG_VEL"]] = np.sum(np.square(ang_u[:, :2]), 1)
  # variable_posture :291-359
  ms = (total_cmd < F(cfg.walking_threshold)).astype(F)[:, None]
  mw = ((total_cmd >= F(cfg.walking_threshold)) & (total_cmd < F(cfg.running_threshold))).astype(F)[:, None]
  mr = (total_cmd >= F(cfg.running_threshold)).astype(F)[:, None]
  std = cfg.std_standing.astype(F) * ms + cfg.std_walking.astype(F) * mw + cfg.std_running.astype(F) * mr
  jp = d.qpos[:, cfg.joint_q_adr].astype(F)
  out[:, K["B2E_REW_POSE"]] = np.exp(-np.mean(np.square(jp - cfg.default_joint_pos.astype(F)) / std ** 2, 1))
  # angular_momentum_penalty :110-122
  am = sd[:, cfg.angmom_adr: cfg.angmom_adr + 3]
  am_sq = np.sum(np.square(am), 1)
  out[:, K["B2E_REW_ANGMOM"]] = am_sq
  # joint_pos_limits envs/mdp/rewards.py:73-88, action_rate_l2 :56-60
  soft = cfg.soft_joint_pos_limits.astype(F)
  out[:, K["B2E_REW_DOF_POS_LIMITS"]] = np.sum(-np.minimum(jp - soft[:, 0], 0) + np.maximum(jp - soft[:, 1], 0), 1)
  out[:, K["B2E_REW_ACTION_RATE"]] = np.sum(np.square(s.action - s.prev_action), 1)
  # feet terms
  cur_air = s.air_state[:, 0]
  found = sd[:, cfg.foot_found_adr]
  force = np.stack([sd[:, a: a + 3] for a in cfg.foot_force_adr], 1)
  site_pos = d.site_xpos[:, cfg.foot_site_ids].astype(F)
  site_body = cfg.foot_site_body
  lin_s, _ = velocity_from_cvel(site_pos, d.subtree_com[:, cfg.root_body].astype(F)[:, None], d.cvel[:, site_body].astype(F))
  vel_norm = np.sqrt(np.sum(np.square(lin_s[..., :2]), -1))
  fc = first_contact(cfg, s)
  active_air = (total_cmd > F(cfg.air_cmd_threshold)).astype(F)
  active = (total_cmd > F(cfg.cmd_threshold)).astype(F)
  out[:, K["B2E_REW_AIR_TIME"]] = np.sum(((cur_air > F(cfg.air_thr_min)) & (cur_air < F(cfg.air_thr_max))).astype(F), 1) * active_air  # :125-153
  out[:, K["B2E_REW_FOOT_CLEARANCE"]] = np.sum(np.abs(site_pos[..., 2] - F(cfg.target_height)) * vel_norm, 1) * active  # :156-177
  if w[K["B2E_REW_FOOT_SWING"]] != 0:  # :180-229 (stateful; only evaluated when the term is active, reward_manager.py:82-84)
    s.peak_heights[:] = np.where(found == 0, np.maximum(s.peak_heights, site_pos[..., 2]), s.peak_heights)
    err = s.peak_heights / F(cfg.target_height) - F(1)
    out[:, K["B2E_REW_FOOT_SWING"]] = np.sum(np.square(err) * fc.astype(F), 1) * active
    if log is not None:
      log[:, K["B2E_LOG_PEAK_SUM"]] = np.sum(s.peak_heights * fc, 1)
    s.peak_heights[:] = np.where(fc, 0, s.peak_heights)
  in_contact = (found > 0).astype(F)
  out[:, K["B2E_REW_FOOT_SLIP"]] = np.sum(np.square(vel_norm) * in_contact, 1) * active  # :232-259
  fmag = np.sqrt(np.sum(np.square(force), -1))
  out[:, K["B2E_REW_SOFT_LANDING"]] = np.sum(fmag * fc.astype(F), 1) * active  # :262-288
  if cfg.selfcol_adr >= 0:
    out[:, K["B2E_REW_SELF_COLLISIONS"]] = sd[:, cfg.selfcol_adr]  # :88-95
  # generic library, envs/mdp/rewards.py:22-70,149-155
  if terminated is not None:
    out[:, K["B2E_REW_IS_ALIVE"]] = (~terminated).astype(F)
    out[:, K["B2E_REW_IS_TERMINATED"]] = terminated.astype(F)
  if getattr(d, "actuator_force", None) is not None:   # (data views of callers that never weight these terms may lack the arrays)
    out[:, K["B2E_REW_JOINT_TORQUES_L2"]] = np.sum(np.square(d.actuator_force[:, cfg.ctrl_ids].astype(F)), 1)
    mech = d.actuator_force[:, cfg.ctrl_ids].astype(F) * d.qvel[:, cfg.joint_v_adr].astype(F)       # electrical_power_cost, :125-146
    out[:, K["B2E_REW_ELECTRICAL_POWER"]] = np.sum(np.maximum(mech, F(0)), 1)
  out[:, K["B2E_REW_JOINT_VEL_L2"]] = np.sum(np.square(d.qvel[:, cfg.joint_v_adr].astype(F)), 1)
  if getattr(d, "qacc", None) is not None:
    out[:, K["B2E_REW_JOINT_ACC_L2"]] = np.sum(np.square(d.qacc[:, cfg.joint_v_adr].astype(F)), 1)
  out[:, K["B2E_REW_ACTION_ACC_L2"]] = np.sum(np.square(s.action - F(2) * s.prev_action + s.prev_prev_action), 1)
  out[:, K["B2E_REW_FLAT_ORIENTATION_L2"]] = np.sum(np.square(R.pg[:, :2]), 1)
  if log is not None:
    if w[K["B2E_REW_ANGMOM"]] != 0:
      log[:, K["B2E_LOG_ANGMOM"]] = np.sqrt(am_sq)
    if w[K["B2E_REW_AIR_TIME"]] != 0:
      log[:, K["B2E_LOG_AIR_TIME_SUM"]] = np.sum(cur_air * (cur_air > 0), 1)
      log[:, K["B2E_LOG_AIR_COUNT"]] = np.sum(cur_air > 0, 1)
    log[:, K["B2E_LOG_LANDINGS"]] = np.sum(fc, 1)
    if w[K["B2E_REW_FOOT_SLIP"]] != 0:
      log[:, K["B2E_LOG_SLIP_SUM"]] = np.sum(vel_norm * in_contact, 1)
      log[:, K["B2E_LOG_NCONTACT"]] = np.sum(in_contact, 1)
    if w[K["B2E_REW_SOFT_LANDING"]] != 0:
      log[:, K["B2E_LOG_LANDING_FORCE_SUM"]] = np.sum(fmag * fc, 1)
  return out


def reward_total(cfg, s, terms):
  """RewardManager.compute (managers/reward_manager.py:77-91)."""
  dt = F(cfg.step_dt)
  s.reward[:] = 0
  for k in range(A.NREW):
    wk = F(cfg.reward_weights[k])
    if wk == 0:
      s.step_reward[:, k] = 0
      continue
    v = np.nan_to_num(terms[:, k] * wk * dt, nan=0.0, posinf=0.0, neginf=0.0).astype(F)
    s.reward += v
    s.episode_sums[:, k] += v
    s.step_reward[:, k] = v / dt


# ---------------------------------------------------------------------------------------------------------
# command (managers/command_manager.py:40-67, tasks/velocity/mdp/velocity_command.py:51-101)
# ---------------------------------------------------------------------------------------------------------
def _uni(u, lo, hi):
  return (u * F(F(hi) - F(lo)) + F(lo)).astype(F)


def resample_command(cfg, s, ids, u, slot0) -> None:
  s.cmd_time_left[ids] = _uni(u[ids, slot0], *cfg.cmd_resample)
  s.command[ids, 0] = _uni(u[ids, slot0 + 1], *cfg.cmd_lin_x)
  s.command[ids, 1] = _uni(u[ids, slot0 + 2], *cfg.cmd_lin_y)
  s.command[ids, 2] = _uni(u[ids, slot0 + 3], *cfg.cmd_ang_z)
  s.heading_target[ids] = _uni(u[ids, slot0 + 4], *cfg.cmd_heading)
  s.is_heading[ids] = u[ids, slot0 + 5] <= F(cfg.rel_heading)
  s.is_standing[ids] = u[ids, slot0 + 6] <= F(cfg.rel_standing)
  s.cmd_counter[ids] += 1


def update_command(cfg, s, quat) -> None:
  """_update_command, velocity_command.py:91-101"""
  s.heading_error[:] = wrap_to_pi(s.heading_target - heading_w(quat))
  ih = s.is_heading
  s.command[ih, 2] = np.clip(F(cfg.heading_stiffness) * s.heading_error[ih], F(cfg.cmd_ang_z[0]), F(cfg.cmd_ang_z[1]))
  s.command[s.is_standing] = 0


# ---------------------------------------------------------------------------------------------------------
# observations (tasks/velocity/velocity_env_cfg.py:40-101, envs/mdp/observations.py, tasks/velocity/mdp/observations.py)
# ---------------------------------------------------------------------------------------------------------
def observations(cfg, s, d, R, u=None):
  sd = d.sensordata.astype(F)
  pol = np.concatenate([
    sd[:, cfg.linvel_adr: cfg.linvel_adr + 3], sd[:, cfg.angvel_adr: cfg.angvel_adr + 3], R.pg,
    d.qpos[:, cfg.joint_q_adr].astype(F) - cfg.default_joint_pos.astype(F), d.qvel[:, cfg.joint_v_adr].astype(F), s.action, s.command], 1)
  force = np.concatenate([sd[:, a: a + 3] for a in cfg.foot_force_adr], 1)
  cri = np.concatenate([pol, d.site_xpos[:, cfg.foot_site_ids, 2].astype(F), s.air_state[:, 0], (sd[:, cfg.foot_found_adr] > 0).astype(F),
                        np.sign(force) * np.log1p(np.abs(force))], 1)
  if cfg.enable_corruption and u is not None:  # UniformNoiseCfg.apply, utils/noise/noise_cfg.py:61-76
    lo, hi = cfg.noise_lo.astype(F), cfg.noise_hi.astype(F)
    n = u[:, K["B2E_DRAW_NOISE"]: K["B2E_DRAW_NOISE"] + cfg.n_obs_policy] * (hi - lo) + lo
    pol = np.where(lo != hi, pol + n, pol).astype(F)
  s.obs_policy[:] = pol
  s.obs_critic[:] = cri
  s.obs_extra[:] = np.concatenate([R.lin_b, R.ang_b], 1)   # base_lin_vel, base_ang_vel (envs/mdp/observations.py:24-35)


# ---------------------------------------------------------------------------------------------------------
# the two halves of ManagerBasedRlEnv.step after the physics loop (envs/manager_based_rl_env.py:282-313,381-416)
# ---------------------------------------------------------------------------------------------------------
def post(cfg, m, s, d, u=None, do_reset=True) -> None:
  B = s.command.shape[0]
  if u is None:
    u = draws(cfg.seed, B, s.step_counter)
  s.log_env[:] = 0
  R = root_kinematics(cfg, d)
  s.episode_length += 1
  time_out, fell, illegal, root_low, has_nan = terminations(cfg, s, d, R)
  s.terminated[:] = fell | illegal | root_low | has_nan
  terms = reward_terms(cfg, s, d, R, log=s.log_env, terminated=s.terminated)
  s.last_terms = terms
  reward_total(cfg, s, terms)
  s.term_dones[:, 0], s.term_dones[:, 1], s.term_dones[:, 2], s.term_dones[:, 3], s.term_dones[:, 4] = time_out, fell, illegal, root_low, has_nan
  s.time_outs[:] = time_out
  s.reset_buf[:] = s.terminated | s.time_outs
  if not do_reset:
    return
  reset_idx(cfg, m, s, d, np.nonzero(s.reset_buf)[0], u)
  if cfg.terrain_mode:
    s.log_env[:, K["B2E_LOG_TERRAIN_LEVEL"]] = s.terrain_level


def reset_idx(cfg, m, s, d, ids, u) -> None:
  """ManagerBasedRlEnv._reset_idx (envs/manager_based_rl_env.py:381-416) for the envs `ids`."""
  if len(ids) == 0:
    return
  # log sums of the finished episodes (reward_manager.py:59-75, termination_manager.py:74-85, command_manager.py:40-51)
  s.log_env[ids, K["B2E_LOG_NRESET"]] = 1
  s.log_env[ids, K["B2E_LOG_EPISODE_SUM"]: K["B2E_LOG_EPISODE_SUM"] + A.NREW] = s.episode_sums[ids]
  s.log_env[ids, K["B2E_LOG_TERM_COUNT"]: K["B2E_LOG_TERM_COUNT"] + A.NTERM] = s.term_dones[ids]
  s.log_env[ids, K["B2E_LOG_CMD_ERR_XY"]] = s.cmd_metrics[ids, 0]
  s.log_env[ids, K["B2E_LOG_CMD_ERR_YAW"]] = s.cmd_metrics[ids, 1]
  # curriculum_manager.compute comes first (manager_based_rl_env.py:382): terrain_levels_vel on the pre-reset state
  # (tasks/velocity/mdp/curriculums.py:30-64) + TerrainImporter.update_env_origins (terrains/terrain_importer.py:196-209)
  if cfg.terrain_mode & 1:
    rows = cfg.terrain_rows
    dist = np.sqrt(np.sum(np.square(d.xpos[ids, cfg.root_body, :2].astype(F) - s.env_origins[ids, :2]), 1))
    up = dist > F(cfg.terrain_half_size)
    down = (dist < np.sqrt(np.sum(np.square(s.command[ids, :2]), 1)) * F(cfg.max_episode_length_s) * F(0.5)) & ~up
    lv = s.terrain_level[ids] + up.astype(np.int32) - down.astype(np.int32)
    restart = np.minimum((u[ids, K["B2E_DRAW_TERRAIN"]] * F(rows)).astype(np.int32), rows - 1)
    lv = np.where(lv >= rows, restart, np.maximum(lv, 0))
    s.terrain_level[ids] = lv
    s.env_origins[ids] = cfg.terrain_origins.astype(F)[lv, s.terrain_type[ids]]
  # sim.reset (mjwarp.reset_data, sim/sim.py:275-286)
  d.qpos[ids] = m.qpos0
  for f in ("qvel", "qacc", "qacc_warmstart", "ctrl", "time"):
    getattr(d, f)[ids] = 0
  d.sensordata[ids] = 0
  # scene.reset: Entity.clear_state (entity/data.py:193-198) + ContactSensor.reset (sensor/contact_sensor.py:240-253)
  s.joint_pos_target[ids] = 0
  s.air_state[ids] = 0
  s.air_last_time[ids] = d.time[ids]
  # reset_root_state_uniform (envs/mdp/events.py:81-168)
  rp, rv = cfg.reset_pose, cfg.reset_vel
  ps = np.stack([_uni(u[ids, K["B2E_DRAW_RESET_POSE"] + k], *rp[k]) for k in range(6)], 1)
  vs = np.stack([_uni(u[ids, K["B2E_DRAW_RESET_VEL"] + k], *rv[k]) for k in range(6)], 1)
  drs = cfg.default_root_state.astype(F)
  pos = drs[None, 0:3] + ps[:, 0:3] + s.env_origins[ids]
  quat = quat_mul(np.broadcast_to(drs[3:7], (len(ids), 4)), quat_from_euler_xyz(ps[:, 3], ps[:, 4], ps[:, 5])).astype(F)
  vel = drs[None, 7:13] + vs
  d.qpos[ids, 0:3], d.qpos[ids, 3:7] = pos, quat
  d.qvel[ids, 0:3], d.qvel[ids, 3:6] = vel[:, 0:3], quat_apply_inverse(quat, vel[:, 3:6])  # entity/data.py:99-110
  if cfg.terrain_mode & 2:
    # play mode: randomize_terrain (envs/mdp/events.py:26-37) is the last reset event; the pose above used the old origin
    rows, cols = cfg.terrain_rows, cfg.terrain_cols
    s.terrain_level[ids] = np.minimum((u[ids, K["B2E_DRAW_TERRAIN"] + 1] * F(rows)).astype(np.int32), rows - 1)
    s.terrain_type[ids] = np.minimum((u[ids, K["B2E_DRAW_TERRAIN"] + 2] * F(cols)).astype(np.int32), cols - 1)
    s.env_origins[ids] = cfg.terrain_origins.astype(F)[s.terrain_level[ids], s.terrain_type[ids]]
  # reset_joints_by_offset (envs/mdp/events.py:171-206)
  nj = cfg.nj
  jp = cfg.default_joint_pos.astype(F)[None] + _uni(u[ids, K["B2E_DRAW_RESET_JOINT_POS"]: K["B2E_DRAW_RESET_JOINT_POS"] + nj], *cfg.reset_joint_pos)
  jp = np.clip(jp, cfg.soft_joint_pos_limits[:, 0].astype(F), cfg.soft_joint_pos_limits[:, 1].astype(F))
  jv = _uni(u[ids, K["B2E_DRAW_RESET_JOINT_VEL"]: K["B2E_DRAW_RESET_JOINT_VEL"] + nj], *cfg.reset_joint_vel)
  d.qpos[np.ix_(ids, cfg.joint_q_adr)] = jp
  d.qvel[np.ix_(ids, cfg.joint_v_adr)] = jv
  # manager resets (manager_based_rl_env.py:392-416)
  for a in (s.action, s.prev_action, s.prev_prev_action, s.raw_action):
    a[ids] = 0
  s.episode_sums[ids] = 0
  s.cmd_metrics[ids] = 0
  s.cmd_counter[ids] = 0
  resample_command(cfg, s, ids, u, K["B2E_DRAW_RESET_CMD"])
  s.episode_length[ids] = 0


def obs_stage(cfg, m, s, d, u=None) -> None:
  B = s.command.shape[0]
  if u is None:
    u = draws(cfg.seed, B, s.step_counter)
  R = root_kinematics(cfg, d)
  dt = F(cfg.step_dt)
  # CommandManager.compute (command_manager.py:53-59): metrics, timers, resample, update
  max_step = F(cfg.cmd_resample[1]) / dt
  s.cmd_metrics[:, 0] += np.sqrt(np.sum(np.square(s.command[:, :2] - R.lin_b[:, :2]), 1)) / max_step
  s.cmd_metrics[:, 1] += np.abs(s.command[:, 2] - R.ang_b[:, 2]) / max_step
  s.cmd_time_left -= dt
  ids = np.nonzero(s.cmd_time_left <= 0)[0]
  if len(ids):
    resample_command(cfg, s, ids, u, K["B2E_DRAW_CMD"])
  update_command(cfg, s, R.quat)
  # interval event push_by_setting_velocity (managers/event_manager.py:124-146, envs/mdp/events.py:209-223)
  if cfg.push_enabled:
    s.push_time_left -= dt
    ids = np.nonzero(s.push_time_left < F(1e-6))[0]
    if len(ids):
      s.push_time_left[ids] = _uni(u[ids, K["B2E_DRAW_PUSH"]], *cfg.push_interval)
      pr = cfg.push_vel
      dv = np.stack([_uni(u[ids, K["B2E_DRAW_PUSH"] + 1 + k], *pr[k]) for k in range(6)], 1)
      vel = np.concatenate([R.lin_w[ids], R.ang_w[ids]], 1) + dv
      d.qvel[ids, 0:3] = vel[:, 0:3]
      d.qvel[ids, 3:6] = quat_apply_inverse(d.qpos[ids, 3:7].astype(F), vel[:, 3:6])
  observations(cfg, s, d, R, u)
  s.log[:] = s.log_env.sum(0)
  s.step_counter += 1
