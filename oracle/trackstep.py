"""CPU oracle of the tracking-task env step (TEST INFRASTRUCTURE, never on the product path).

Numpy restatement, batched over envs, of mjlab's motion-tracking MDP (paths relative to /root/reference/src/mjlab):
`MotionCommand` tasks/tracking/mdp/commands.py:70-412, rewards tasks/tracking/mdp/rewards.py:26-120 +
envs/mdp/rewards.py:56-88, observations tasks/tracking/mdp/observations.py:17-69 + envs/mdp/observations.py, terminations
tasks/tracking/mdp/terminations.py:17-86, configured as tasks/tracking/tracking_env_cfg.py:44-317 +
config/g1/env_cfgs.py:15-100.  PINNED by tests/golden/mdp_tracking_g1.npz, which holds the outputs of the reference's OWN
classes / functions on oracle physics states (tools/make_trk_golden.py): frame lookup, anchor-relative body targets,
metrics, all nine reward terms, the three motion terminations, both observation groups.  Reset sampling (`_resample_command`)
follows the cited lines with the kernels' Philox draws; the reference's `torch.multinomial` over bins is restated as
inverse-CDF sampling with one uniform draw per env (same distribution, different stream).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""

from __future__ import annotations

from types import SimpleNamespace as NS

import numpy as np

from mjlab_b200 import trkabi as A
from oracle.envstep import (draws, quat_apply, quat_apply_inverse, quat_from_euler_xyz, quat_mul, velocity_from_cvel,  # noqa: F401
                            _uni)

K = A.K
F = np.float32


def quat_conjugate(q):
  return np.concatenate([q[..., 0:1], -q[..., 1:]], -1)


def quat_inv(q):
  """utils/lab_api/math.py:261-272"""
  return quat_conjugate(q) / np.maximum(np.sum(q * q, -1, keepdims=True), F(1e-9))


def yaw_quat(q):
  """utils/lab_api/math.py:566-587"""
  qw, qx, qy, qz = (q[..., i] for i in range(4))
  yaw = np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))
  out = np.zeros_like(q)
  out[..., 0], out[..., 3] = np.cos(yaw / 2), np.sin(yaw / 2)
  return out / np.maximum(np.linalg.norm(out, axis=-1, keepdims=True), F(1e-9))


def axis_angle_from_quat(q, eps=1e-6):
  """utils/lab_api/math.py:478-507"""
  q = q * (1.0 - 2.0 * (q[..., 0:1] < 0.0)).astype(q.dtype)
  mag = np.linalg.norm(q[..., 1:], axis=-1)
  half = np.arctan2(mag, q[..., 0])
  ang = 2 * half
  with np.errstate(divide="ignore", invalid="ignore"):
    s = np.where(np.abs(ang) > eps, np.sin(half) / ang, 0.5 - ang * ang / 48)
  return q[..., 1:4] / s[..., None]


def quat_error_magnitude(q1, q2):
  """utils/lab_api/math.py:688-700 (quat_box_minus :590-605)"""
  return np.linalg.norm(axis_angle_from_quat(quat_mul(q1, quat_conjugate(q2))), axis=-1)


def matrix_from_quat(q):
  """utils/lab_api/math.py:166-196"""
  r, i, j, k = (q[..., n] for n in range(4))
  two_s = 2.0 / np.sum(q * q, -1)
  o = np.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r), two_s * (i * j + k * r),
                1 - two_s * (i * i + k * k), two_s * (j * k - i * r), two_s * (i * k - j * r), two_s * (j * k + i * r),
                1 - two_s * (i * i + j * j)], -1)
  return o.reshape(q.shape[:-1] + (3, 3))


def subtract_frame_transforms(t01, q01, t02, q02):
  """utils/lab_api/math.py:832-864"""
  q10 = quat_inv(q01)
  return quat_apply(q10, t02 - t01), quat_mul(q10, q02)


# ---------------------------------------------------------------------------------------------------------
def make_state(cfg, nworld: int, env_origins=None) -> NS:
  nj, nb = cfg.nj, len(cfg.body_sel)
  z = lambda *s: np.zeros((nworld,) + s, F)  # noqa: E731
  rq = z(nb, 4)
  rq[..., 0] = 1
  return NS(
    env_origins=z(3) if env_origins is None else np.asarray(env_origins, F),
    action=z(nj), prev_action=z(nj), prev_prev_action=z(nj), raw_action=z(nj), processed_action=z(nj), encoder_bias=z(nj),
    joint_pos_target=z(nj), time_steps=np.zeros(nworld, np.int64), rel_pos=z(nb, 3), rel_quat=rq, metrics=z(A.NMETRIC),
    push_time_left=z(), episode_length=np.zeros(nworld, np.int64), reward=z(), step_reward=z(A.NREW), episode_sums=z(A.NREW),
    term_dones=np.zeros((nworld, A.NTERM), bool), terminated=np.zeros(nworld, bool), time_outs=np.zeros(nworld, bool),
    reset_buf=np.zeros(nworld, bool), obs_policy=z(cfg.n_obs_policy), obs_critic=z(cfg.n_obs_critic),
    bin_failed_count=np.zeros(cfg.bin_count, F), current_bin_failed=np.zeros(cfg.bin_count, F), step_counter=0,
  )


def motion_frame(cfg, mo, s) -> NS:
  """MotionCommand frame properties (commands.py:131-177) at the envs' current time steps; tracked bodies only."""
  t = s.time_steps
  return NS(joint_pos=mo["joint_pos"][t], joint_vel=mo["joint_vel"][t],
            body_pos_w=mo["body_pos_w"][t][:, cfg.body_sel] + s.env_origins[:, None, :], body_quat_w=mo["body_quat_w"][t][:, cfg.body_sel],
            body_lin_vel_w=mo["body_lin_vel_w"][t][:, cfg.body_sel], body_ang_vel_w=mo["body_ang_vel_w"][t][:, cfg.body_sel])


def robot_bodies(cfg, d) -> NS:
  """robot_body_* properties (commands.py:187-217): tracked bodies from the engine's derived data."""
  g = cfg.body_ids[cfg.body_sel]
  pos = d.xpos[:, g].astype(F)
  lin, ang = velocity_from_cvel(pos, d.subtree_com[:, cfg.root_body].astype(F)[:, None], d.cvel[:, g].astype(F))
  return NS(pos=pos, quat=d.xquat[:, g].astype(F), lin=lin, ang=ang)


def terminations(cfg, s, fr, rb):
  a = cfg.anchor_sel
  time_out = s.episode_length >= cfg.max_episode_length
  full = int(getattr(cfg, "term_full3d", 0))
  if full & 1:
    anchor_pos = np.sqrt(np.sum(np.square(fr.body_pos_w[:, a] - rb.pos[:, a]), -1)) > F(cfg.anchor_pos_threshold)   # terminations.py:18-25
  else:
    anchor_pos = np.abs(fr.body_pos_w[:, a, 2] - rb.pos[:, a, 2]) > F(cfg.anchor_pos_threshold)                   # terminations.py:28-35
  g = np.broadcast_to(np.array([0, 0, -1], F), (len(time_out), 3))
  anchor_ori = np.abs(quat_apply_inverse(fr.body_quat_w[:, a], g)[:, 2] - quat_apply_inverse(rb.quat[:, a], g)[:, 2]) > F(cfg.anchor_ori_threshold)  # :38-52
  if full & 2:
    ee = np.any(np.sqrt(np.sum(np.square(s.rel_pos[:, cfg.ee_sel] - rb.pos[:, cfg.ee_sel]), -1)) > F(cfg.ee_pos_threshold), -1)   # :55-70
  else:
    ee = np.any(np.abs(s.rel_pos[:, cfg.ee_sel, 2] - rb.pos[:, cfg.ee_sel, 2]) > F(cfg.ee_pos_threshold), -1)   # :73-86
  return time_out, anchor_pos, anchor_ori, ee


def reward_terms(cfg, s, d, fr, rb):
  B, a = len(s.time_steps), cfg.anchor_sel
  out = np.zeros((B, A.NREW), F)
  out[:, 0] = np.exp(-np.sum(np.square(fr.body_pos_w[:, a] - rb.pos[:, a]), -1) / F(cfg.std_anchor_pos) ** 2)       # rewards.py:26-35
  out[:, 1] = np.exp(-quat_error_magnitude(fr.body_quat_w[:, a], rb.quat[:, a]) ** 2 / F(cfg.std_anchor_ori) ** 2)  # :38-44
  out[:, 2] = np.exp(-np.sum(np.square(s.rel_pos - rb.pos), -1).mean(-1) / F(cfg.std_body_pos) ** 2)                # :47-63
  out[:, 3] = np.exp(-(quat_error_magnitude(s.rel_quat, rb.quat) ** 2).mean(-1) / F(cfg.std_body_ori) ** 2)         # :66-82
  out[:, 4] = np.exp(-np.sum(np.square(fr.body_lin_vel_w - rb.lin), -1).mean(-1) / F(cfg.std_body_lin_vel) ** 2)    # :85-101
  out[:, 5] = np.exp(-np.sum(np.square(fr.body_ang_vel_w - rb.ang), -1).mean(-1) / F(cfg.std_body_ang_vel) ** 2)    # :104-120
  out[:, 6] = np.sum(np.square(s.action - s.prev_action), 1)                                                         # envs/mdp/rewards.py:56-60
  jp = d.qpos[:, cfg.joint_q_adr].astype(F)
  soft = cfg.soft_joint_pos_limits.astype(F)
  out[:, 7] = np.sum(-np.minimum(jp - soft[:, 0], 0) + np.maximum(jp - soft[:, 1], 0), 1)                            # :73-88
  out[:, 8] = d.sensordata[:, cfg.selfcol_adr].astype(F)
  return out


def reward_total(cfg, s, terms):
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


def update_metrics(cfg, s, d, fr, rb):
  """_update_metrics, commands.py:219-255"""
  a = cfg.anchor_sel
  m = s.metrics
  m[:, 0] = np.linalg.norm(fr.body_pos_w[:, a] - rb.pos[:, a], axis=-1)
  m[:, 1] = quat_error_magnitude(fr.body_quat_w[:, a], rb.quat[:, a])
  m[:, 2] = np.linalg.norm(fr.body_lin_vel_w[:, a] - rb.lin[:, a], axis=-1)
  m[:, 3] = np.linalg.norm(fr.body_ang_vel_w[:, a] - rb.ang[:, a], axis=-1)
  m[:, 4] = np.linalg.norm(s.rel_pos - rb.pos, axis=-1).mean(-1)
  m[:, 5] = quat_error_magnitude(s.rel_quat, rb.quat).mean(-1)
  m[:, 6] = np.linalg.norm(fr.joint_pos - d.qpos[:, cfg.joint_q_adr].astype(F), axis=-1)
  m[:, 7] = np.linalg.norm(fr.joint_vel - d.qvel[:, cfg.joint_v_adr].astype(F), axis=-1)


def sampling_probabilities(cfg, s):
  """_adaptive_sampling, commands.py:270-282: uniform floor + non-causal exponential kernel + normalisation."""
  p = s.bin_failed_count + F(cfg.adaptive_uniform_ratio) / F(cfg.bin_count)
  ks = cfg.adaptive_kernel_size
  kern = np.array([cfg.adaptive_lambda ** i for i in range(ks)], F)
  kern = kern / kern.sum()
  padded = np.concatenate([p, np.repeat(p[-1:], ks - 1)])
  sm = np.array([np.dot(padded[i:i + ks], kern) for i in range(cfg.bin_count)], F)
  return sm / sm.sum()


def resample_command(cfg, mo, s, d, ids, u, slot0, terminated) -> None:
  """_resample_command, commands.py:311-377: pick the start frame, then write the (perturbed) motion state to the sim."""
  if len(ids) == 0:
    return
  T = cfg.time_step_total
  if cfg.sampling_mode == "start":
    s.time_steps[ids] = 0
  elif cfg.sampling_mode == "uniform":
    s.time_steps[ids] = np.minimum((u[ids, slot0] * F(T)).astype(np.int64), T - 1)
  else:
    failed = terminated[ids]
    if failed.any():
      cur = np.clip((s.time_steps * cfg.bin_count) // max(T, 1), 0, cfg.bin_count - 1)
      s.current_bin_failed += np.bincount(cur[ids][failed], minlength=cfg.bin_count).astype(F)
    cdf = np.cumsum(sampling_probabilities(cfg, s).astype(np.float64)).astype(F)
    bins = np.minimum(np.searchsorted(cdf, u[ids, slot0], side="right"), cfg.bin_count - 1)
    s.time_steps[ids] = ((bins.astype(F) + u[ids, slot0 + 1]) / F(cfg.bin_count) * F(T - 1)).astype(np.int64)
  fr = motion_frame(cfg, mo, s)
  ps = np.stack([_uni(u[ids, slot0 + 2 + k], *cfg.pose_range[k]) for k in range(6)], 1)
  vs = np.stack([_uni(u[ids, slot0 + 8 + k], *cfg.velocity_range[k]) for k in range(6)], 1)
  root_pos = fr.body_pos_w[ids, 0] + ps[:, 0:3]
  root_ori = quat_mul(quat_from_euler_xyz(ps[:, 3], ps[:, 4], ps[:, 5]), fr.body_quat_w[ids, 0]).astype(F)
  lin = fr.body_lin_vel_w[ids, 0] + vs[:, :3]
  ang = fr.body_ang_vel_w[ids, 0] + vs[:, 3:]
  nj = cfg.nj
  jp = fr.joint_pos[ids] + _uni(u[ids, slot0 + 14: slot0 + 14 + nj], *cfg.joint_position_range)
  jp = np.clip(jp, cfg.soft_joint_pos_limits[:, 0].astype(F), cfg.soft_joint_pos_limits[:, 1].astype(F))
  d.qpos[np.ix_(ids, cfg.joint_q_adr)] = jp
  d.qvel[np.ix_(ids, cfg.joint_v_adr)] = fr.joint_vel[ids]
  d.qpos[ids, 0:3], d.qpos[ids, 3:7] = root_pos, root_ori
  d.qvel[ids, 0:3], d.qvel[ids, 3:6] = lin, quat_apply_inverse(root_ori, ang)   # entity/data.py:99-110
  s.joint_pos_target[ids] = 0                                                     # robot.clear_state


def update_command(cfg, mo, s, d, rb, u) -> None:
  """_update_command, commands.py:379-412"""
  s.time_steps += 1
  ids = np.nonzero(s.time_steps >= cfg.time_step_total)[0]
  resample_command(cfg, mo, s, d, ids, u, K["B2T_DRAW_CMD"], s.terminated)
  fr = motion_frame(cfg, mo, s)
  a = cfg.anchor_sel
  nb = len(cfg.body_sel)
  anchor_pos = np.repeat(fr.body_pos_w[:, a][:, None], nb, 1)
  anchor_quat = np.repeat(fr.body_quat_w[:, a][:, None], nb, 1)
  delta_pos = np.repeat(rb.pos[:, a][:, None], nb, 1).copy()
  delta_pos[..., 2] = anchor_pos[..., 2]
  delta_ori = yaw_quat(quat_mul(np.repeat(rb.quat[:, a][:, None], nb, 1), quat_inv(anchor_quat))).astype(F)
  s.rel_quat[:] = quat_mul(delta_ori, fr.body_quat_w)
  s.rel_pos[:] = delta_pos + quat_apply(delta_ori, fr.body_pos_w - anchor_pos)
  if cfg.sampling_mode == "adaptive":
    s.bin_failed_count[:] = F(cfg.adaptive_alpha) * s.current_bin_failed + F(1 - cfg.adaptive_alpha) * s.bin_failed_count
    s.current_bin_failed[:] = 0
  return fr


def observations(cfg, s, d, fr, rb, u=None):
  a = cfg.anchor_sel
  sd = d.sensordata.astype(F)
  nb = len(cfg.body_sel)
  B = len(s.time_steps)
  pos_b, ori_b = subtract_frame_transforms(rb.pos[:, a], rb.quat[:, a], fr.body_pos_w[:, a], fr.body_quat_w[:, a])
  ori6 = matrix_from_quat(ori_b)[..., :2].reshape(B, -1)
  bp, bo = subtract_frame_transforms(np.repeat(rb.pos[:, a][:, None], nb, 1), np.repeat(rb.quat[:, a][:, None], nb, 1), rb.pos, rb.quat)
  bo6 = matrix_from_quat(bo)[..., :2].reshape(B, -1)
  cmd = np.concatenate([fr.joint_pos, fr.joint_vel], 1)
  lin, ang = sd[:, cfg.linvel_adr: cfg.linvel_adr + 3], sd[:, cfg.angvel_adr: cfg.angvel_adr + 3]
  jp = d.qpos[:, cfg.joint_q_adr].astype(F)
  jv = d.qvel[:, cfg.joint_v_adr].astype(F)
  dflt = cfg.default_joint_pos.astype(F)
  pol = np.concatenate([cmd, pos_b, ori6, lin, ang, jp + s.encoder_bias - dflt, jv, s.action], 1)
  cri = np.concatenate([cmd, pos_b, ori6, bp.reshape(B, -1), bo6, lin, ang, jp - dflt, jv, s.action], 1)
  if cfg.enable_corruption and u is not None:
    lo, hi = cfg.noise_lo.astype(F), cfg.noise_hi.astype(F)
    n = u[:, K["B2T_DRAW_NOISE"]: K["B2T_DRAW_NOISE"] + cfg.n_obs_policy] * (hi - lo) + lo
    pol = np.where(lo != hi, pol + n, pol).astype(F)
  s.obs_policy[:] = pol
  s.obs_critic[:] = cri


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


def post(cfg, m, mo, s, d, u=None, do_reset=True) -> None:
  """terminations, rewards, _reset_idx (manager_based_rl_env.py:282-298,381-416 with the tracking managers)."""
  B = len(s.time_steps)
  if u is None:
    u = draws_trk(cfg.seed, B, s.step_counter)
  fr, rb = motion_frame(cfg, mo, s), robot_bodies(cfg, d)
  s.episode_length += 1
  time_out, apos, aori, ee = terminations(cfg, s, fr, rb)
  terms = reward_terms(cfg, s, d, fr, rb)
  s.last_terms = terms
  reward_total(cfg, s, terms)
  s.term_dones[:, 0], s.term_dones[:, 1], s.term_dones[:, 2], s.term_dones[:, 3] = time_out, apos, aori, ee
  s.terminated[:] = apos | aori | ee
  s.time_outs[:] = time_out
  s.reset_buf[:] = s.terminated | s.time_outs
  if not do_reset:
    return
  reset_idx(cfg, m, mo, s, d, np.nonzero(s.reset_buf)[0], u)


def reset_idx(cfg, m, mo, s, d, ids, u) -> None:
  if len(ids) == 0:
    return
  d.qpos[ids] = m.qpos0
  for f in ("qvel", "qacc", "qacc_warmstart", "ctrl", "time"):
    getattr(d, f)[ids] = 0
  d.sensordata[ids] = 0
  s.joint_pos_target[ids] = 0
  for a in (s.action, s.prev_action, s.prev_prev_action, s.raw_action):
    a[ids] = 0
  s.episode_sums[ids] = 0
  s.metrics[ids] = 0
  resample_command(cfg, mo, s, d, ids, u, K["B2T_DRAW_RESET"], s.terminated)
  s.episode_length[ids] = 0


def obs_stage(cfg, m, mo, s, d, u=None) -> None:
  B = len(s.time_steps)
  if u is None:
    u = draws_trk(cfg.seed, B, s.step_counter)
  rb = robot_bodies(cfg, d)
  update_metrics(cfg, s, d, motion_frame(cfg, mo, s), rb)
  fr = update_command(cfg, mo, s, d, rb, u)
  dt = F(cfg.step_dt)
  if cfg.push_enabled:  # push_by_setting_velocity with the tracking velocity range (tracking_env_cfg.py:173-178)
    s.push_time_left -= dt
    ids = np.nonzero(s.push_time_left < F(1e-6))[0]
    if len(ids):
      s.push_time_left[ids] = _uni(u[ids, K["B2T_DRAW_PUSH"]], *cfg.push_interval)
      rbv = cfg.root_body
      lin, ang = velocity_from_cvel(d.xpos[ids, rbv].astype(F), d.subtree_com[ids, rbv].astype(F), d.cvel[ids, rbv].astype(F))
      dv = np.stack([_uni(u[ids, K["B2T_DRAW_PUSH"] + 1 + k], *cfg.push_vel[k]) for k in range(6)], 1)
      vel = np.concatenate([lin, ang], 1) + dv
      d.qvel[ids, 0:3] = vel[:, 0:3]
      d.qvel[ids, 3:6] = quat_apply_inverse(d.qpos[ids, 3:7].astype(F), vel[:, 3:6])
  observations(cfg, s, d, fr, rb, u)
  s.step_counter += 1


def draws_trk(seed: int, nworld: int, step: int) -> np.ndarray:
  """[nworld, B2T_NDRAW] uniforms: same generator and indexing as the velocity env (env_draw in b2_env.cuh)."""
  from oracle.envstep import philox4x32
  nblk = (A.NDRAW + 3) // 4
  w = np.arange(nworld, dtype=np.uint64)[:, None]
  blk = np.arange(nblk, dtype=np.uint64)[None, :]
  o = philox4x32(seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, w, step & 0xFFFFFFFF, blk, (step >> 32) & 0xFFFFFFFF)
  u = (o >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)
  return u.reshape(nworld, nblk * 4)[:, : A.NDRAW]
