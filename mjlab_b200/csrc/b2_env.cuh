// b2_env.cuh - fused manager-based env step of the velocity task: one warp per env.
//
// Each *_world function is the per-env body of one kernel of include/b200env.h.  The 32 lanes span the joints /
// observation entries of their env, reductions are warp shuffles.  Like b2_smooth / b2_solve the file also
// compiles single-lane for the host (-DB2_HOSTEMU, tests only).
//
// Reference semantics restated here (file:line relative to /root/reference/src/mjlab):
//   envs/manager_based_rl_env.py:272-313,381-416      step / _reset_idx ordering
//   managers/action_manager.py:113-130, envs/mdp/actions/joint_actions.py:84-104
//   sensor/contact_sensor.py:240-269,327-367           air-time state machine, compute_first_contact
//   tasks/velocity/mdp/rewards.py:23-359, envs/mdp/rewards.py:56-88
//   envs/mdp/terminations.py:19-32, tasks/velocity/mdp/terminations.py:13-16
//   managers/reward_manager.py:59-91, managers/termination_manager.py:74-97
//   managers/command_manager.py:40-67, tasks/velocity/mdp/velocity_command.py:51-101
//   managers/event_manager.py:124-146, envs/mdp/events.py:81-223
//   managers/observation_manager.py:154-208, utils/noise/noise_cfg.py:61-76
//   entity/data.py:20-31,99-110,212-533
#pragma once

#include "b2_common.cuh"
#include "../../include/b200env.h"

struct KEnv {
  int ip[B2E_IP_COUNT];
  float fp[B2E_FP_COUNT];
  unsigned long long seed;
  void* buf[B2E_B_COUNT];
};

#define EI(id) (e.ip[id])
#define EF(id) ((real)e.fp[id])
#define EBR(id) (reinterpret_cast<real*>(e.buf[id]))
#define EBI(id) (reinterpret_cast<int*>(e.buf[id]))
#define EBU(id) (reinterpret_cast<unsigned char*>(e.buf[id]))
#define EBL(id) (reinterpret_cast<long long*>(e.buf[id]))

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a draw is a pure function of
// (seed, env, step, slot) and the numpy oracle reproduces it bit for bit.
// ---------------------------------------------------------------------------
B2_DEV void b2_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1) for draw slot `slot` of env w at env-step `step`
B2_DEV real env_draw(const KEnv& e, int nworld, int w, long long step, int slot) {
  const real* inj = EBR(B2E_B_INJECT_U);
  if (inj) return inj[(size_t)w * B2E_NDRAW + slot];
  uint32_t o[4];
  b2_philox((uint32_t)e.seed, (uint32_t)(e.seed >> 32), (uint32_t)w, (uint32_t)step, (uint32_t)(slot >> 2), (uint32_t)((unsigned long long)step >> 32), o);
  return (real)(o[slot & 3] >> 8) * (real)(1.0 / 16777216.0);
}
B2_DEV real env_uniform(const KEnv& e, int nworld, int w, long long step, int slot, real lo, real hi) {
  return env_draw(e, nworld, w, step, slot) * (hi - lo) + lo;
}

// quat_apply / quat_apply_inverse (utils/lab_api/math.py:629-670): v +- w*t + xyz x t, t = 2 xyz x v
B2_DEV void env_quat_apply(real* r, const real* q, const real* v) {
  real t[3], c[3];
  cross3(t, q + 1, v);
  t[0] *= 2; t[1] *= 2; t[2] *= 2;
  cross3(c, q + 1, t);
  r[0] = v[0] + q[0] * t[0] + c[0]; r[1] = v[1] + q[0] * t[1] + c[1]; r[2] = v[2] + q[0] * t[2] + c[2];
}
B2_DEV void env_quat_apply_inv(real* r, const real* q, const real* v) {
  real t[3], c[3];
  cross3(t, q + 1, v);
  t[0] *= 2; t[1] *= 2; t[2] *= 2;
  cross3(c, q + 1, t);
  r[0] = v[0] - q[0] * t[0] + c[0]; r[1] = v[1] - q[0] * t[1] + c[1]; r[2] = v[2] - q[0] * t[2] + c[2];
}
// compute_velocity_from_cvel (entity/data.py:20-31): lin = cvel_lin - cvel_ang x (subtree_com - pos)
B2_DEV void env_vel_from_cvel(real* lin, real* ang, const real* pos, const real* com, const real* cvel) {
  real off[3] = {com[0] - pos[0], com[1] - pos[1], com[2] - pos[2]}, c[3];
  cross3(c, cvel, off);
  lin[0] = cvel[3] - c[0]; lin[1] = cvel[4] - c[1]; lin[2] = cvel[5] - c[2];
  ang[0] = cvel[0]; ang[1] = cvel[1]; ang[2] = cvel[2];
}
B2_DEV real env_wrap_to_pi(real a) {
  const real PI = (real)3.14159265358979323846, TWO_PI = 2 * PI;
  real r = fmod(a + PI, TWO_PI);
  if (r < 0) r += TWO_PI;
  return (r == 0 && a > 0) ? PI : r - PI;
}
B2_DEV bool env_finite(real x) { return x == x && x - x == 0; }   // neither NaN nor +-Inf
B2_DEV real env_nan_to_num(real v) { return (v == v && v - v == 0) ? v : (real)0; }

// DcMotorActuator._clip_effort (actuator/dc_actuator.py:136-162): torque-speed line, then the continuous limit
B2_DEV float b2_dc_clip(float effort, float vel, float limit, float sat, float vlim) {
  float corner = vlim * (1.f + limit / sat);
  float v = vel < -corner ? -corner : (vel > corner ? corner : vel);
  float top = sat * (1.f - v / vlim), bottom = sat * (-1.f - v / vlim);
  float hi = top < limit ? top : limit, lo = bottom > -limit ? bottom : -limit;
  return effort < lo ? lo : (effort > hi ? hi : effort);  // torch.clamp(x, min, max) = min(max(x, min), max)
}

struct EnvRoot {
  real quat[4], pos[3], lin_w[3], ang_w[3], lin_b[3], ang_b[3], pg[3];
};
B2_DEV void env_root(const KModel& m, const KData& d, const KEnv& e, int w, EnvRoot& r) {
  const int rb = EI(B2E_IP_ROOT_BODY);
  const real* xq = d.df[B2_DF_XQUAT] + ((size_t)w * m.nbody + rb) * 4;
  const real* xp = d.df[B2_DF_XPOS] + ((size_t)w * m.nbody + rb) * 3;
  const real* com = d.df[B2_DF_SUBTREE_COM] + ((size_t)w * m.nbody + rb) * 3;
  const real* cv = d.df[B2_DF_CVEL] + ((size_t)w * m.nbody + rb) * 6;
  for (int c = 0; c < 4; ++c) r.quat[c] = xq[c];
  for (int c = 0; c < 3; ++c) r.pos[c] = xp[c];
  env_vel_from_cvel(r.lin_w, r.ang_w, xp, com, cv);
  env_quat_apply_inv(r.lin_b, r.quat, r.lin_w);
  env_quat_apply_inv(r.ang_b, r.quat, r.ang_w);
  const real g[3] = {0, 0, -1};
  env_quat_apply_inv(r.pg, r.quat, g);
}

// UniformVelocityCommand._resample (command_manager.py:61-67, velocity_command.py:65-76); draws at slot0..slot0+6
B2_DEV void env_resample_command(const KEnv& e, int nworld, int w, long long step, int slot0) {
  EBR(B2E_B_CMD_TIME_LEFT)[w] = env_uniform(e, nworld, w, step, slot0, EF(B2E_FP_CMD_RESAMPLE), EF(B2E_FP_CMD_RESAMPLE + 1));
  real* cmd = EBR(B2E_B_COMMAND) + (size_t)w * 3;
  cmd[0] = env_uniform(e, nworld, w, step, slot0 + 1, EF(B2E_FP_CMD_LIN_X), EF(B2E_FP_CMD_LIN_X + 1));
  cmd[1] = env_uniform(e, nworld, w, step, slot0 + 2, EF(B2E_FP_CMD_LIN_Y), EF(B2E_FP_CMD_LIN_Y + 1));
  cmd[2] = env_uniform(e, nworld, w, step, slot0 + 3, EF(B2E_FP_CMD_ANG_Z), EF(B2E_FP_CMD_ANG_Z + 1));
  if (EI(B2E_IP_HEADING_COMMAND)) {
    EBR(B2E_B_HEADING_TARGET)[w] = env_uniform(e, nworld, w, step, slot0 + 4, EF(B2E_FP_CMD_HEADING), EF(B2E_FP_CMD_HEADING + 1));
    EBU(B2E_B_IS_HEADING)[w] = env_draw(e, nworld, w, step, slot0 + 5) <= EF(B2E_FP_REL_HEADING);
  }
  EBU(B2E_B_IS_STANDING)[w] = env_draw(e, nworld, w, step, slot0 + 6) <= EF(B2E_FP_REL_STANDING);
  EBI(B2E_B_CMD_COUNTER)[w] += 1;
}

// ---------------------------------------------------------------------------
// b2_env_action
// ---------------------------------------------------------------------------
B2_DEV void env_action_world(const KModel& m, const KData& d, const KEnv& e, int w) {
  const int nj = EI(B2E_IP_NJ);
  const int* ctrl_ids = EBI(B2E_B_CTRL_IDS);
  for (int j = B2_LANE; j < nj; j += B2_NLANE) {
    size_t o = (size_t)w * nj + j;
    real a_in = EBR(B2E_B_ACTION_IN)[o];
    EBR(B2E_B_PREV_PREV_ACTION)[o] = EBR(B2E_B_PREV_ACTION)[o];
    EBR(B2E_B_PREV_ACTION)[o] = EBR(B2E_B_ACTION)[o];
    EBR(B2E_B_ACTION)[o] = a_in;
    EBR(B2E_B_RAW_ACTION)[o] = a_in;
    real p = a_in * EBR(B2E_B_ACTION_SCALE)[j] + EBR(B2E_B_ACTION_OFFSET)[j];
    EBR(B2E_B_PROCESSED_ACTION)[o] = p;
    real target = p - EBR(B2E_B_ENCODER_BIAS)[o];
    EBR(B2E_B_JOINT_POS_TARGET)[o] = target;
    d.df[B2_DF_CTRL][(size_t)w * m.nu + ctrl_ids[j]] = target;
  }
}

// ---------------------------------------------------------------------------
// b2_env_airtime (one lane per foot)
// ---------------------------------------------------------------------------
B2_DEV void env_airtime_world(const KModel& m, const KData& d, const KEnv& e, int w) {
  const int nf = EI(B2E_IP_NFEET);
  const real now = d.df[B2_DF_TIME][w];
  const real elapsed = now - EBR(B2E_B_AIR_LAST_TIME)[w];
  real* st = EBR(B2E_B_AIR_STATE) + (size_t)w * 4 * nf;
  const real* sd = d.df[B2_DF_SENSORDATA] + (size_t)w * m.nsensordata;
  for (int f = B2_LANE; f < nf; f += B2_NLANE) {
    bool contact = sd[EI(B2E_IP_FOOT_FOUND_ADR + f)] > 0;
    real cur_air = st[f], cur_con = st[2 * nf + f];
    bool first_contact = cur_air > 0 && contact, first_detached = cur_con > 0 && !contact;
    if (first_contact) st[nf + f] = cur_air + elapsed;
    st[f] = contact ? (real)0 : cur_air + elapsed;
    if (first_detached) st[3 * nf + f] = cur_con + elapsed;
    st[2 * nf + f] = contact ? cur_con + elapsed : (real)0;
  }
  B2_WSYNC();
  if (B2_LANE == 0) EBR(B2E_B_AIR_LAST_TIME)[w] = now;
}

// mjwarp.reset_data for one world (also the body of k_reset)
B2_DEV void reset_world(const KModel& m, const KData& d, int w) {
  const real* qpos0 = MF(B2_MF_QPOS0);
  for (int i = B2_LANE; i < m.nq; i += B2_NLANE) d.df[B2_DF_QPOS][(size_t)w * m.nq + i] = qpos0[i];
  for (int i = B2_LANE; i < m.nv; i += B2_NLANE) {
    size_t o = (size_t)w * m.nv + i;
    d.df[B2_DF_QVEL][o] = 0; d.df[B2_DF_QACC][o] = 0; d.df[B2_DF_QACC_WARMSTART][o] = 0; d.df[B2_DF_QFRC_APPLIED][o] = 0;
  }
  for (int i = B2_LANE; i < m.nu; i += B2_NLANE) d.df[B2_DF_CTRL][(size_t)w * m.nu + i] = 0;
  for (int i = B2_LANE; i < m.nbody * 6; i += B2_NLANE) d.df[B2_DF_XFRC_APPLIED][(size_t)w * m.nbody * 6 + i] = 0;
  for (int i = B2_LANE; i < m.nsensordata; i += B2_NLANE) d.df[B2_DF_SENSORDATA][(size_t)w * m.nsensordata + i] = 0;
  const int* mocap = MI(B2_MI_BODY_MOCAPID);
  const real* bpos = MF(B2_MF_BODY_POS);
  const real* bquat = MF(B2_MF_BODY_QUAT);
  for (int b = B2_LANE; b < m.nbody; b += B2_NLANE) {
    int k = mocap[b];
    if (k < 0) continue;
    for (int c = 0; c < 3; ++c) d.df[B2_DF_MOCAP_POS][((size_t)w * m.nmocap + k) * 3 + c] = bpos[3 * b + c];
    for (int c = 0; c < 4; ++c) d.df[B2_DF_MOCAP_QUAT][((size_t)w * m.nmocap + k) * 4 + c] = bquat[4 * b + c];
  }
  if (B2_LANE == 0) {
    d.df[B2_DF_TIME][w] = 0;
    d.di[B2_DI_NCON][w] = 0; d.di[B2_DI_NEFC][w] = 0; d.di[B2_DI_NL][w] = 0; d.di[B2_DI_NF][w] = 0; d.di[B2_DI_SOLVER_NITER][w] = 0;
  }
  B2_WSYNC();
}

B2_DEVN void env_reset_world(const KModel& m, const KData& d, const KEnv& e, int w, long long step);

// ---------------------------------------------------------------------------
// b2_env_post: counters, terminations, rewards, reset
// ---------------------------------------------------------------------------
B2_DEV void env_post_world(const KModel& m, const KData& d, const KEnv& e, int w) {
  const int nj = EI(B2E_IP_NJ), nf = EI(B2E_IP_NFEET), nworld = d.nworld;
  const long long step = EBL(B2E_B_STEP_COUNTER)[0];
  const real dt = EF(B2E_FP_STEP_DT);
  const real* sd = d.df[B2_DF_SENSORDATA] + (size_t)w * m.nsensordata;
  const real* qpos = d.df[B2_DF_QPOS] + (size_t)w * m.nq;
  real* logv = EBR(B2E_B_LOG_ENV) + (size_t)w * B2E_NLOG;
  for (int i = B2_LANE; i < B2E_NLOG; i += B2_NLANE) logv[i] = 0;
  B2_WSYNC();

  EnvRoot R;
  env_root(m, d, e, w, R);
  const real* cmd = EBR(B2E_B_COMMAND) + (size_t)w * 3;
  const real c0 = cmd[0], c1 = cmd[1], c2 = cmd[2];
  const real total_cmd = sqrt(c0 * c0 + c1 * c1) + fabs(c2);

  // ---- episode counter + terminations --------------------------------------------------
  long long ep = EBL(B2E_B_EPISODE_LENGTH)[w] + 1;
  bool time_out = ep >= (long long)EI(B2E_IP_MAX_EPISODE_LENGTH);
  bool fell = fabs(acos(-R.pg[2])) > EF(B2E_FP_LIMIT_ANGLE);
  int ill = 0;
  for (int k = B2_LANE; k < EI(B2E_IP_ILLEGAL_COUNT); k += B2_NLANE) ill |= sd[EI(B2E_IP_ILLEGAL_ADR) + k] != 0;
  bool illegal = b2_wor_i(ill) != 0;
  const bool has_illegal = EI(B2E_IP_ILLEGAL_COUNT) > 0;
  // generic terminations (envs/mdp/terminations.py:35-47), off unless the configuration lists them
  const int tflags = EI(B2E_IP_TERM_FLAGS);
  const bool root_low = (tflags & 1) && R.pos[2] < EF(B2E_FP_MIN_ROOT_HEIGHT);
  bool has_nan = false;
  if (tflags & 2) {   // NanGuard.detect_nans: qpos, qvel, qacc, qacc_warmstart
    int bad = 0;
    const real* qv = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
    const real* qa = d.df[B2_DF_QACC] + (size_t)w * m.nv;
    const real* qw = d.df[B2_DF_QACC_WARMSTART] + (size_t)w * m.nv;
    for (int i = B2_LANE; i < m.nq; i += B2_NLANE) bad |= !env_finite(qpos[i]);
    for (int i = B2_LANE; i < m.nv; i += B2_NLANE) bad |= !(env_finite(qv[i]) && env_finite(qa[i]) && env_finite(qw[i]));
    has_nan = b2_wor_i(bad) != 0;
  }
  bool terminated = fell || (has_illegal && illegal) || root_low || has_nan;
  bool reset = terminated || time_out;

  // ---- per-joint reward sums ------------------------------------------------------------
  const int* jq = EBI(B2E_B_JOINT_Q_ADR);
  real pose_acc = 0, lim_acc = 0, rate_acc = 0, tau_acc = 0, vel_acc = 0, acc_acc = 0, aacc_acc = 0, pow_acc = 0;
  {
    const int* jv = EBI(B2E_B_JOINT_V_ADR);
    const int* cid = EBI(B2E_B_CTRL_IDS);
    const real* g_qvel = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
    const real* g_qacc = d.df[B2_DF_QACC] + (size_t)w * m.nv;
    const real* g_af = d.df[B2_DF_ACTUATOR_FORCE] + (size_t)w * m.nu;
    const real wt = EF(B2E_FP_WALKING_THRESHOLD), rt = EF(B2E_FP_RUNNING_THRESHOLD);
    const real ms = total_cmd < wt ? 1 : 0, mw = (total_cmd >= wt && total_cmd < rt) ? 1 : 0, mr = total_cmd >= rt ? 1 : 0;
    for (int j = B2_LANE; j < nj; j += B2_NLANE) {
      size_t o = (size_t)w * nj + j;
      real q = qpos[jq[j]];
      real std = EBR(B2E_B_STD_STANDING)[j] * ms + EBR(B2E_B_STD_WALKING)[j] * mw + EBR(B2E_B_STD_RUNNING)[j] * mr;
      real err = q - EBR(B2E_B_DEFAULT_JOINT_POS)[j];
      pose_acc += err * err / (std * std);
      real lo = q - EBR(B2E_B_SOFT_LO)[j], hi = q - EBR(B2E_B_SOFT_HI)[j];
      lim_acc += -(lo < 0 ? lo : (real)0) + (hi > 0 ? hi : (real)0);
      real da = EBR(B2E_B_ACTION)[o] - EBR(B2E_B_PREV_ACTION)[o];
      rate_acc += da * da;
      // generic library: joint_torques_l2, joint_vel_l2, joint_acc_l2, action_acc_l2 (envs/mdp/rewards.py:32-70)
      real tau = g_af[cid[j]], qd = g_qvel[jv[j]], qdd = g_qacc[jv[j]];
      tau_acc += tau * tau; vel_acc += qd * qd; acc_acc += qdd * qdd;
      { const real mech = tau * qd; pow_acc += mech > 0 ? mech : (real)0; }   // electrical_power_cost: regeneration is not penalised
      real dda = EBR(B2E_B_ACTION)[o] - 2 * EBR(B2E_B_PREV_ACTION)[o] + EBR(B2E_B_PREV_PREV_ACTION)[o];
      aacc_acc += dda * dda;
    }
    pose_acc = b2_wsum(pose_acc); lim_acc = b2_wsum(lim_acc); rate_acc = b2_wsum(rate_acc);
    tau_acc = b2_wsum(tau_acc); vel_acc = b2_wsum(vel_acc); acc_acc = b2_wsum(acc_acc); aacc_acc = b2_wsum(aacc_acc);
    pow_acc = b2_wsum(pow_acc);
  }

  // ---- per-foot quantities (lane f < nfeet) ------------------------------------------------
  real air_rew = 0, clear_cost = 0, swing_cost = 0, slip_cost = 0, land_cost = 0;
  real m_air_sum = 0, m_air_cnt = 0, m_peak = 0, m_land = 0, m_slip = 0, m_ncon = 0, m_force = 0;
  {
    for (int f = B2_LANE; f < nf; f += B2_NLANE) {
      const int rb = EI(B2E_IP_ROOT_BODY);
      const real* st = EBR(B2E_B_AIR_STATE) + (size_t)w * 4 * nf;
      const int site = EI(B2E_IP_FOOT_SITE + f), sbody = EI(B2E_IP_FOOT_SITE_BODY + f);
      const real* sp = d.df[B2_DF_SITE_XPOS] + ((size_t)w * m.nsite + site) * 3;
      const real* com = d.df[B2_DF_SUBTREE_COM] + ((size_t)w * m.nbody + rb) * 3;
      const real* cv = d.df[B2_DF_CVEL] + ((size_t)w * m.nbody + sbody) * 6;
      real lin[3], ang[3];
      env_vel_from_cvel(lin, ang, sp, com, cv);
      const real vel_norm = sqrt(lin[0] * lin[0] + lin[1] * lin[1]);
      const real found = sd[EI(B2E_IP_FOOT_FOUND_ADR + f)];
      const real* fr = sd + EI(B2E_IP_FOOT_FORCE_ADR + f);
      const real cur_air = st[f], cur_con = st[2 * nf + f];
      const bool first_contact = cur_con > 0 && cur_con < dt + (real)1e-8;
      // feet_air_time (rewards.py:125-153)
      air_rew += (cur_air > EF(B2E_FP_AIR_THR_MIN) && cur_air < EF(B2E_FP_AIR_THR_MAX)) ? 1 : 0;
      if (cur_air > 0) { m_air_sum += cur_air; m_air_cnt += 1; }
      // feet_clearance (rewards.py:156-177)
      clear_cost += fabs(sp[2] - EF(B2E_FP_TARGET_HEIGHT)) * vel_norm;
      // feet_swing_height (rewards.py:180-229): stateful peak tracking, evaluated at landing
      if (EF(B2E_FP_REWARD_WEIGHT + B2E_REW_FOOT_SWING) != 0) {
        real* peak = EBR(B2E_B_PEAK_HEIGHTS) + (size_t)w * nf + f;
        real pk = *peak;
        if (found == 0) pk = pk > sp[2] ? pk : sp[2];
        real err = pk / EF(B2E_FP_TARGET_HEIGHT) - 1;
        if (first_contact) { swing_cost += err * err; m_peak += pk; pk = 0; }
        *peak = pk;
      }
      // feet_slip (rewards.py:232-259)
      if (found > 0) { slip_cost += vel_norm * vel_norm; m_slip += vel_norm; m_ncon += 1; }
      // soft_landing (rewards.py:262-288)
      if (first_contact) { real fm = sqrt(fr[0] * fr[0] + fr[1] * fr[1] + fr[2] * fr[2]); land_cost += fm; m_force += fm; m_land += 1; }
    }
    air_rew = b2_wsum(air_rew); clear_cost = b2_wsum(clear_cost); swing_cost = b2_wsum(swing_cost);
    slip_cost = b2_wsum(slip_cost); land_cost = b2_wsum(land_cost);
    m_air_sum = b2_wsum(m_air_sum); m_air_cnt = b2_wsum(m_air_cnt); m_peak = b2_wsum(m_peak); m_land = b2_wsum(m_land);
    m_slip = b2_wsum(m_slip); m_ncon = b2_wsum(m_ncon); m_force = b2_wsum(m_force);
  }

  // ---- scalar terms + accumulation (lane 0) --------------------------------------------------
  if (B2_LANE == 0) {
    real term[B2E_NREW];
    const real active_air = total_cmd > EF(B2E_FP_AIR_CMD_THRESHOLD) ? 1 : 0;
    const real active = total_cmd > EF(B2E_FP_CMD_THRESHOLD) ? 1 : 0;
    {
      real ex = c0 - R.lin_b[0], ey = c1 - R.lin_b[1];
      real err = (ex * ex + ey * ey) + R.lin_b[2] * R.lin_b[2];
      real s = EF(B2E_FP_STD_LIN);
      term[B2E_REW_TRACK_LIN] = exp(-err / (s * s));
    }
    {
      real ez = c2 - R.ang_b[2];
      real err = ez * ez + (R.ang_b[0] * R.ang_b[0] + R.ang_b[1] * R.ang_b[1]);
      real s = EF(B2E_FP_STD_ANG);
      term[B2E_REW_TRACK_ANG] = exp(-err / (s * s));
    }
    {
      const int ub = EI(B2E_IP_UPRIGHT_BODY);
      const real* uq = d.df[B2_DF_XQUAT] + ((size_t)w * m.nbody + ub) * 4;
      const real g[3] = {0, 0, -1};
      real pgu[3];
      env_quat_apply_inv(pgu, uq, g);
      real s = EF(B2E_FP_STD_UPRIGHT);
      term[B2E_REW_UPRIGHT] = exp(-(pgu[0] * pgu[0] + pgu[1] * pgu[1]) / (s * s));
      const real* ucv = d.df[B2_DF_CVEL] + ((size_t)w * m.nbody + ub) * 6;
      term[B2E_REW_BODY_ANG_VEL] = ucv[0] * ucv[0] + ucv[1] * ucv[1];
    }
    term[B2E_REW_POSE] = exp(-(pose_acc / (real)nj));
    {
      const real* am = sd + EI(B2E_IP_ANGMOM_ADR);
      real sq = am[0] * am[0] + am[1] * am[1] + am[2] * am[2];
      term[B2E_REW_ANGMOM] = sq;
      if (EF(B2E_FP_REWARD_WEIGHT + B2E_REW_ANGMOM) != 0) logv[B2E_LOG_ANGMOM] = sqrt(sq);
    }
    term[B2E_REW_DOF_POS_LIMITS] = lim_acc;
    term[B2E_REW_ACTION_RATE] = rate_acc;
    term[B2E_REW_AIR_TIME] = air_rew * active_air;
    term[B2E_REW_FOOT_CLEARANCE] = clear_cost * active;
    term[B2E_REW_FOOT_SWING] = swing_cost * active;
    term[B2E_REW_FOOT_SLIP] = slip_cost * active;
    term[B2E_REW_SOFT_LANDING] = land_cost * active;
    term[B2E_REW_SELF_COLLISIONS] = EI(B2E_IP_SELFCOL_ADR) >= 0 ? sd[EI(B2E_IP_SELFCOL_ADR)] : (real)0;
    term[B2E_REW_IS_ALIVE] = terminated ? 0 : 1;
    term[B2E_REW_IS_TERMINATED] = terminated ? 1 : 0;
    term[B2E_REW_JOINT_TORQUES_L2] = tau_acc;
    term[B2E_REW_JOINT_VEL_L2] = vel_acc;
    term[B2E_REW_JOINT_ACC_L2] = acc_acc;
    term[B2E_REW_ACTION_ACC_L2] = aacc_acc;
    term[B2E_REW_FLAT_ORIENTATION_L2] = R.pg[0] * R.pg[0] + R.pg[1] * R.pg[1];
    term[B2E_REW_ELECTRICAL_POWER] = pow_acc;
    if (EF(B2E_FP_REWARD_WEIGHT + B2E_REW_AIR_TIME) != 0) { logv[B2E_LOG_AIR_TIME_SUM] = m_air_sum; logv[B2E_LOG_AIR_COUNT] = m_air_cnt; }
    if (EF(B2E_FP_REWARD_WEIGHT + B2E_REW_FOOT_SWING) != 0) logv[B2E_LOG_PEAK_SUM] = m_peak;
    logv[B2E_LOG_LANDINGS] = m_land;
    if (EF(B2E_FP_REWARD_WEIGHT + B2E_REW_FOOT_SLIP) != 0) { logv[B2E_LOG_SLIP_SUM] = m_slip; logv[B2E_LOG_NCONTACT] = m_ncon; }
    if (EF(B2E_FP_REWARD_WEIGHT + B2E_REW_SOFT_LANDING) != 0) logv[B2E_LOG_LANDING_FORCE_SUM] = m_force;

    // RewardManager.compute (reward_manager.py:77-91)
    real total = 0;
    real* step_rew = EBR(B2E_B_STEP_REWARD) + (size_t)w * B2E_NREW;
    real* sums = EBR(B2E_B_EPISODE_SUMS) + (size_t)w * B2E_NREW;
    for (int k = 0; k < B2E_NREW; ++k) {
      real wk = EF(B2E_FP_REWARD_WEIGHT + k);
      if (wk == 0) { step_rew[k] = 0; continue; }
      real v = env_nan_to_num(term[k] * wk * dt);
      total += v;
      sums[k] += v;
      step_rew[k] = v / dt;
    }
    EBR(B2E_B_REWARD)[w] = total;
    unsigned char* td = EBU(B2E_B_TERM_DONES) + (size_t)w * B2E_NTERM;
    td[B2E_TERM_TIME_OUT] = time_out; td[B2E_TERM_FELL_OVER] = fell; td[B2E_TERM_ILLEGAL_CONTACT] = has_illegal && illegal;
    td[B2E_TERM_ROOT_HEIGHT] = root_low; td[B2E_TERM_NAN] = has_nan;
    EBU(B2E_B_TERMINATED)[w] = terminated; EBU(B2E_B_TIME_OUTS)[w] = time_out; EBU(B2E_B_RESET_BUF)[w] = reset;
    EBL(B2E_B_EPISODE_LENGTH)[w] = ep;
  }
  B2_WSYNC();
  if (reset) env_reset_world(m, d, e, w, step);
  if (EI(B2E_IP_TERRAIN_MODE) && B2_LANE == 0) logv[B2E_LOG_TERRAIN_LEVEL] = (real)EBI(B2E_B_TERRAIN_LEVEL)[w];
}

// ---------------------------------------------------------------------------
// _reset_idx for one env (manager_based_rl_env.py:381-416); also the body of b2_env_reset
// ---------------------------------------------------------------------------
B2_DEVN void env_reset_world(const KModel& m, const KData& d, const KEnv& e, int w, long long step) {
  const int nj = EI(B2E_IP_NJ), nf = EI(B2E_IP_NFEET), nworld = d.nworld;
  const int* jq = EBI(B2E_B_JOINT_Q_ADR);
  if (B2_LANE == 0) {
    // log contributions of the finished episode (reward_manager.py:59-75, termination_manager.py:74-85, command_manager.py:40-51)
    real* logv = EBR(B2E_B_LOG_ENV) + (size_t)w * B2E_NLOG;
    real* sums = EBR(B2E_B_EPISODE_SUMS) + (size_t)w * B2E_NREW;
    const unsigned char* td = EBU(B2E_B_TERM_DONES) + (size_t)w * B2E_NTERM;
    logv[B2E_LOG_NRESET] = 1;
    for (int k = 0; k < B2E_NREW; ++k) { logv[B2E_LOG_EPISODE_SUM + k] = sums[k]; sums[k] = 0; }
    for (int k = 0; k < B2E_NTERM; ++k) logv[B2E_LOG_TERM_COUNT + k] = td[k];
    real* met = EBR(B2E_B_CMD_METRICS) + (size_t)w * 2;
    logv[B2E_LOG_CMD_ERR_XY] = met[0]; logv[B2E_LOG_CMD_ERR_YAW] = met[1];
    met[0] = 0; met[1] = 0;
    EBL(B2E_B_EPISODE_LENGTH)[w] = 0;
  }
  const int tmode = EI(B2E_IP_TERRAIN_MODE);
  if ((tmode & 1) && B2_LANE == 0) {
    // curriculum_manager.compute first (manager_based_rl_env.py:382): terrain_levels_vel (tasks/velocity/mdp/curriculums.py:30-64)
    // on the pre-reset state + TerrainImporter.update_env_origins (terrains/terrain_importer.py:196-209)
    const real* rp = d.df[B2_DF_XPOS] + ((size_t)w * m.nbody + EI(B2E_IP_ROOT_BODY)) * 3;   // root_link_pos_w
    real* org = EBR(B2E_B_ENV_ORIGINS) + (size_t)w * 3;
    const real* cmd = EBR(B2E_B_COMMAND) + (size_t)w * 3;
    const real dx = rp[0] - org[0], dy = rp[1] - org[1];
    const real dist = sqrt(dx * dx + dy * dy);
    const bool up = dist > EF(B2E_FP_TERRAIN_HALF_SIZE);
    const bool down = (dist < sqrt(cmd[0] * cmd[0] + cmd[1] * cmd[1]) * EF(B2E_FP_MAX_EPISODE_LENGTH_S) * (real)0.5) && !up;
    const int rows = EI(B2E_IP_TERRAIN_ROWS), cols = EI(B2E_IP_TERRAIN_COLS);
    int lv = EBI(B2E_B_TERRAIN_LEVEL)[w] + (up ? 1 : 0) - (down ? 1 : 0);
    if (lv >= rows) {   // past the hardest row: restart at a uniformly drawn level
      lv = (int)(env_draw(e, nworld, w, step, B2E_DRAW_TERRAIN) * (real)rows);
      if (lv > rows - 1) lv = rows - 1;
    } else if (lv < 0) lv = 0;
    EBI(B2E_B_TERRAIN_LEVEL)[w] = lv;
    const real* to = EBR(B2E_B_TERRAIN_ORIGINS) + ((size_t)lv * cols + EBI(B2E_B_TERRAIN_TYPE)[w]) * 3;
    org[0] = to[0]; org[1] = to[1]; org[2] = to[2];
  }
  B2_WSYNC();
  reset_world(m, d, w);                                   // sim.reset
  {                                                       // scene.reset: entity.clear_state + contact sensor reset
    real* st = EBR(B2E_B_AIR_STATE) + (size_t)w * 4 * nf;
    for (int i = B2_LANE; i < 4 * nf; i += B2_NLANE) st[i] = 0;
    if (B2_LANE == 0) EBR(B2E_B_AIR_LAST_TIME)[w] = d.df[B2_DF_TIME][w];
  }
  real* qp = d.df[B2_DF_QPOS] + (size_t)w * m.nq;
  real* qv = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
  if (B2_LANE == 0) {
    // reset_root_state_uniform (events.py:81-168): floating base, free joint at qpos[0:7] / qvel[0:6]
    const real* drs = EBR(B2E_B_DEFAULT_ROOT_STATE);
    const real* org = EBR(B2E_B_ENV_ORIGINS) + (size_t)w * 3;
    real ps[6], vs[6];
    for (int k = 0; k < 6; ++k) {
      ps[k] = env_uniform(e, nworld, w, step, B2E_DRAW_RESET_POSE + k, EF(B2E_FP_RESET_POSE + 2 * k), EF(B2E_FP_RESET_POSE + 2 * k + 1));
      vs[k] = env_uniform(e, nworld, w, step, B2E_DRAW_RESET_VEL + k, EF(B2E_FP_RESET_VEL + 2 * k), EF(B2E_FP_RESET_VEL + 2 * k + 1));
    }
    for (int c = 0; c < 3; ++c) qp[c] = drs[c] + ps[c] + org[c];
    real cr = cos(ps[3] * (real)0.5), sr = sin(ps[3] * (real)0.5), cp = cos(ps[4] * (real)0.5), sp_ = sin(ps[4] * (real)0.5);
    real cy = cos(ps[5] * (real)0.5), sy = sin(ps[5] * (real)0.5);
    real dq[4] = {cy * cr * cp + sy * sr * sp_, cy * sr * cp - sy * cr * sp_, cy * cr * sp_ + sy * sr * cp, sy * cr * cp - cy * sr * sp_};
    real q[4];
    quat_mul(q, drs + 3, dq);
    for (int c = 0; c < 4; ++c) qp[3 + c] = q[c];
    // write_root_velocity (entity/data.py:99-110): world-frame angular velocity -> body frame with the NEW quaternion
    real vel[6], angb[3];
    for (int k = 0; k < 6; ++k) vel[k] = drs[7 + k] + vs[k];
    env_quat_apply_inv(angb, q, vel + 3);
    for (int c = 0; c < 3; ++c) { qv[c] = vel[c]; qv[3 + c] = angb[c]; }
    if (tmode & 2) {
      // play mode: randomize_terrain (envs/mdp/events.py:26-37) is the LAST reset event, so the root pose above used the
      // previous origin and the new sub-terrain takes effect at the next reset - as in the reference
      const int rows = EI(B2E_IP_TERRAIN_ROWS), cols = EI(B2E_IP_TERRAIN_COLS);
      int lv = (int)(env_draw(e, nworld, w, step, B2E_DRAW_TERRAIN + 1) * (real)rows);
      int ty = (int)(env_draw(e, nworld, w, step, B2E_DRAW_TERRAIN + 2) * (real)cols);
      if (lv > rows - 1) lv = rows - 1;
      if (ty > cols - 1) ty = cols - 1;
      EBI(B2E_B_TERRAIN_LEVEL)[w] = lv; EBI(B2E_B_TERRAIN_TYPE)[w] = ty;
      real* o2 = EBR(B2E_B_ENV_ORIGINS) + (size_t)w * 3;
      const real* to = EBR(B2E_B_TERRAIN_ORIGINS) + ((size_t)lv * cols + ty) * 3;
      o2[0] = to[0]; o2[1] = to[1]; o2[2] = to[2];
    }
  }
  {
    // reset_joints_by_offset (events.py:171-206)
    const int* jv = EBI(B2E_B_JOINT_V_ADR);
    for (int j = B2_LANE; j < nj; j += B2_NLANE) {
      real p = EBR(B2E_B_DEFAULT_JOINT_POS)[j] + env_uniform(e, nworld, w, step, B2E_DRAW_RESET_JOINT_POS + j, EF(B2E_FP_RESET_JOINT_POS), EF(B2E_FP_RESET_JOINT_POS + 1));
      p = clampr(p, EBR(B2E_B_SOFT_LO)[j], EBR(B2E_B_SOFT_HI)[j]);
      real v = EBR(B2E_B_DEFAULT_JOINT_VEL)[j] + env_uniform(e, nworld, w, step, B2E_DRAW_RESET_JOINT_VEL + j, EF(B2E_FP_RESET_JOINT_VEL), EF(B2E_FP_RESET_JOINT_VEL + 1));
      qp[jq[j]] = p;
      qv[jv[j]] = v;
      // action manager reset + entity.clear_state
      size_t o = (size_t)w * nj + j;
      EBR(B2E_B_ACTION)[o] = 0; EBR(B2E_B_PREV_ACTION)[o] = 0; EBR(B2E_B_PREV_PREV_ACTION)[o] = 0; EBR(B2E_B_RAW_ACTION)[o] = 0;
      EBR(B2E_B_JOINT_POS_TARGET)[o] = 0;
    }
  }
  if (B2_LANE == 0) {
    // command manager reset (command_manager.py:40-51): counter zeroed, then one resample
    EBI(B2E_B_CMD_COUNTER)[w] = 0;
    env_resample_command(e, nworld, w, step, B2E_DRAW_RESET_CMD);
  }
  B2_WSYNC();
}

// ---------------------------------------------------------------------------
// b2_env_obs: command update, interval events, observations
// ---------------------------------------------------------------------------
// with_update == 0: observations only (ManagerBasedRlEnv.reset, manager_based_rl_env.py:254-270)
B2_DEV void env_obs_world(const KModel& m, const KData& d, const KEnv& e, int w, int with_update) {
  const int nj = EI(B2E_IP_NJ), nf = EI(B2E_IP_NFEET), nworld = d.nworld;
  const long long step = EBL(B2E_B_STEP_COUNTER)[0];
  const real dt = EF(B2E_FP_STEP_DT);
  const real* sd = d.df[B2_DF_SENSORDATA] + (size_t)w * m.nsensordata;
  EnvRoot R;
  env_root(m, d, e, w, R);
  real* cmd = EBR(B2E_B_COMMAND) + (size_t)w * 3;
  if (B2_LANE == 0 && with_update) {
    // CommandManager.compute (command_manager.py:53-59)
    real* met = EBR(B2E_B_CMD_METRICS) + (size_t)w * 2;
    const real max_step = EF(B2E_FP_CMD_RESAMPLE + 1) / dt;
    real ex = cmd[0] - R.lin_b[0], ey = cmd[1] - R.lin_b[1];
    met[0] += sqrt(ex * ex + ey * ey) / max_step;
    met[1] += fabs(cmd[2] - R.ang_b[2]) / max_step;
    real tl = EBR(B2E_B_CMD_TIME_LEFT)[w] - dt;
    EBR(B2E_B_CMD_TIME_LEFT)[w] = tl;
    if (tl <= 0) env_resample_command(e, nworld, w, step, B2E_DRAW_CMD);
    // _update_command (velocity_command.py:91-101)
    if (EI(B2E_IP_HEADING_COMMAND)) {
      const real fwd[3] = {1, 0, 0};
      real fw[3];
      env_quat_apply(fw, R.quat, fwd);
      real herr = env_wrap_to_pi(EBR(B2E_B_HEADING_TARGET)[w] - atan2(fw[1], fw[0]));
      EBR(B2E_B_HEADING_ERROR)[w] = herr;
      if (EBU(B2E_B_IS_HEADING)[w]) cmd[2] = clampr(EF(B2E_FP_HEADING_STIFFNESS) * herr, EF(B2E_FP_CMD_ANG_Z), EF(B2E_FP_CMD_ANG_Z + 1));
    }
    if (EBU(B2E_B_IS_STANDING)[w]) { cmd[0] = 0; cmd[1] = 0; cmd[2] = 0; }
    // interval event push_by_setting_velocity (event_manager.py:124-146, events.py:209-223)
    if (EI(B2E_IP_PUSH_ENABLED)) {
      real pt = EBR(B2E_B_PUSH_TIME_LEFT)[w] - dt;
      if (pt < (real)1e-6) {
        pt = env_uniform(e, nworld, w, step, B2E_DRAW_PUSH, EF(B2E_FP_PUSH_INTERVAL), EF(B2E_FP_PUSH_INTERVAL + 1));
        real vel[6], angb[3];
        for (int k = 0; k < 3; ++k) {
          vel[k] = R.lin_w[k] + env_uniform(e, nworld, w, step, B2E_DRAW_PUSH + 1 + k, EF(B2E_FP_PUSH_VEL + 2 * k), EF(B2E_FP_PUSH_VEL + 2 * k + 1));
          vel[3 + k] = R.ang_w[k] + env_uniform(e, nworld, w, step, B2E_DRAW_PUSH + 4 + k, EF(B2E_FP_PUSH_VEL + 6 + 2 * k), EF(B2E_FP_PUSH_VEL + 6 + 2 * k + 1));
        }
        real* qv = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
        const real* qq = d.df[B2_DF_QPOS] + (size_t)w * m.nq + 3;
        env_quat_apply_inv(angb, qq, vel + 3);
        for (int c = 0; c < 3; ++c) { qv[c] = vel[c]; qv[3 + c] = angb[c]; }
      }
      EBR(B2E_B_PUSH_TIME_LEFT)[w] = pt;
    }
  }
  B2_WSYNC();

  // ---- observations (velocity_env_cfg.py:40-101) ---------------------------------------------
  const int np = EI(B2E_IP_N_OBS_POLICY), nc = EI(B2E_IP_N_OBS_CRITIC);
  const real* qpos = d.df[B2_DF_QPOS] + (size_t)w * m.nq;
  const real* qvel = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
  const int* jq = EBI(B2E_B_JOINT_Q_ADR);
  const int* jv = EBI(B2E_B_JOINT_V_ADR);
  real* op = EBR(B2E_B_OBS_POLICY) + (size_t)w * np;
  real* oc = EBR(B2E_B_OBS_CRITIC) + (size_t)w * nc;
  const bool corrupt = EI(B2E_IP_ENABLE_CORRUPTION) != 0;
  const real* st = EBR(B2E_B_AIR_STATE) + (size_t)w * 4 * nf;
  for (int i = B2_LANE; i < nc; i += B2_NLANE) {
    real v;
    if (i < 3) v = sd[EI(B2E_IP_LINVEL_ADR) + i];
    else if (i < 6) v = sd[EI(B2E_IP_ANGVEL_ADR) + i - 3];
    else if (i < 9) v = R.pg[i - 6];
    else if (i < 9 + nj) { int j = i - 9; v = qpos[jq[j]] - EBR(B2E_B_DEFAULT_JOINT_POS)[j]; }
    else if (i < 9 + 2 * nj) { int j = i - 9 - nj; v = qvel[jv[j]] - EBR(B2E_B_DEFAULT_JOINT_VEL)[j]; }
    else if (i < 9 + 3 * nj) v = EBR(B2E_B_ACTION)[(size_t)w * nj + i - 9 - 2 * nj];
    else if (i < np) v = cmd[i - 9 - 3 * nj];
    else {
      int k = i - np;
      if (k < nf) v = d.df[B2_DF_SITE_XPOS][((size_t)w * m.nsite + EI(B2E_IP_FOOT_SITE + k)) * 3 + 2];
      else if (k < 2 * nf) v = st[k - nf];
      else if (k < 3 * nf) v = sd[EI(B2E_IP_FOOT_FOUND_ADR + k - 2 * nf)] > 0 ? (real)1 : (real)0;
      else {
        int q = k - 3 * nf;
        real f = sd[EI(B2E_IP_FOOT_FORCE_ADR + q / 3) + q % 3];
        real s = f > 0 ? (real)1 : (f < 0 ? (real)-1 : (real)0);
        v = s * log1p(fabs(f));
      }
    }
    oc[i] = v;
    if (i < np) {
      if (corrupt) {
        real lo = EBR(B2E_B_NOISE_LO)[i], hi = EBR(B2E_B_NOISE_HI)[i];
        if (lo != hi) v = v + (env_draw(e, nworld, w, step, B2E_DRAW_NOISE + i) * (hi - lo) + lo);
      }
      op[i] = v;
    }
  }
  for (int i = B2_LANE; i < 6; i += B2_NLANE) EBR(B2E_B_OBS_EXTRA)[(size_t)w * 6 + i] = i < 3 ? R.lin_b[i] : R.ang_b[i - 3];   // base_lin_vel, base_ang_vel
  B2_WSYNC();
}
