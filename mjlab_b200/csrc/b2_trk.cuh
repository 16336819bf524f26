// b2_trk.cuh - fused env step of the motion-tracking task: one warp per env (lanes over tracked bodies / joints /
// observation entries).  Per-env bodies of the kernels of include/b200trk.h; compiles single-lane for the host as well.
//
// Reference semantics restated here (file:line relative to /root/reference/src/mjlab):
//   tasks/tracking/mdp/commands.py:70-412       MotionCommand (frame lookup, _resample_command, _update_command, metrics)
//   tasks/tracking/mdp/rewards.py:26-120        six exp tracking rewards; envs/mdp/rewards.py:56-88
//   tasks/tracking/mdp/terminations.py:28-86    anchor z / anchor orientation / end-effector z
//   tasks/tracking/mdp/observations.py:17-69    anchor and body poses in the robot anchor frame
//   tasks/tracking/tracking_env_cfg.py:44-317   term order, weights, noise, ranges
//   utils/lab_api/math.py:166-196,478-507,566-605,832-864  matrix_from_quat, axis_angle_from_quat, yaw_quat, box_minus, subtract_frame_transforms
#pragma once

#include "b2_env.cuh"
#include "../../include/b200trk.h"

struct KTrk {
  int ip[B2T_IP_COUNT];
  float fp[B2T_FP_COUNT];
  unsigned long long seed;
  void* buf[B2T_B_COUNT];
};

#define TI(id) (e.ip[id])
#define TF(id) ((real)e.fp[id])
#define TBR(id) (reinterpret_cast<real*>(e.buf[id]))
#define TBI(id) (reinterpret_cast<int*>(e.buf[id]))
#define TBU(id) (reinterpret_cast<unsigned char*>(e.buf[id]))
#define TBL(id) (reinterpret_cast<long long*>(e.buf[id]))

B2_DEV real trk_draw(const KTrk& e, int w, long long step, int slot) {
  const real* inj = TBR(B2T_B_INJECT_U);
  if (inj) return inj[(size_t)w * B2T_NDRAW + slot];
  uint32_t o[4];
  b2_philox((uint32_t)e.seed, (uint32_t)(e.seed >> 32), (uint32_t)w, (uint32_t)step, (uint32_t)(slot >> 2), (uint32_t)((unsigned long long)step >> 32), o);
  return (real)(o[slot & 3] >> 8) * (real)(1.0 / 16777216.0);
}
B2_DEV real trk_uniform(const KTrk& e, int w, long long step, int slot, real lo, real hi) { return trk_draw(e, w, step, slot) * (hi - lo) + lo; }

B2_DEV void trk_quat_conj(real* r, const real* q) { r[0] = q[0]; r[1] = -q[1]; r[2] = -q[2]; r[3] = -q[3]; }
B2_DEV void trk_quat_inv(real* r, const real* q) {
  real n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  n = n > (real)1e-9 ? n : (real)1e-9;
  r[0] = q[0] / n; r[1] = -q[1] / n; r[2] = -q[2] / n; r[3] = -q[3] / n;
}
B2_DEV void trk_yaw_quat(real* r, const real* q) {
  real yaw = atan2(2 * (q[0] * q[3] + q[1] * q[2]), 1 - 2 * (q[2] * q[2] + q[3] * q[3]));
  real c = cos(yaw / 2), s = sin(yaw / 2);
  real n = sqrt(c * c + s * s);
  n = n > (real)1e-9 ? n : (real)1e-9;
  r[0] = c / n; r[1] = 0; r[2] = 0; r[3] = s / n;
}
// |log(q1 * conj(q2))| (quat_error_magnitude)
B2_DEV real trk_quat_err(const real* q1, const real* q2) {
  real c[4], d[4];
  trk_quat_conj(c, q2);
  quat_mul(d, q1, c);
  real sgn = d[0] < 0 ? (real)-1 : (real)1;
  for (int k = 0; k < 4; ++k) d[k] *= sgn;
  real mag = sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
  real half = atan2(mag, d[0]), ang = 2 * half;
  real s = fabs(ang) > (real)1e-6 ? sin(half) / ang : (real)0.5 - ang * ang / 48;
  return mag / s;
}
// first two columns of the rotation matrix of q, row-major [r00 r01 r10 r11 r20 r21]
B2_DEV void trk_mat6(real* o, const real* q) {
  real r = q[0], i = q[1], j = q[2], k = q[3];
  real two_s = 2 / (r * r + i * i + j * j + k * k);
  o[0] = 1 - two_s * (j * j + k * k); o[1] = two_s * (i * j - k * r);
  o[2] = two_s * (i * j + k * r);     o[3] = 1 - two_s * (i * i + k * k);
  o[4] = two_s * (i * k - j * r);     o[5] = two_s * (j * k + i * r);
}

struct TrkBody { real pos[3], quat[4], lin[3], ang[3]; };
// robot_body_* of tracked body `sel` (commands.py:187-217)
B2_DEV void trk_robot_body(const KModel& m, const KData& d, const KTrk& e, int w, int sel, TrkBody& b) {
  const int g = TBI(B2T_B_BODY_GLOBAL)[sel], rb = TI(B2T_IP_ROOT_BODY);
  const real* xp = d.df[B2_DF_XPOS] + ((size_t)w * m.nbody + g) * 3;
  const real* xq = d.df[B2_DF_XQUAT] + ((size_t)w * m.nbody + g) * 4;
  const real* com = d.df[B2_DF_SUBTREE_COM] + ((size_t)w * m.nbody + rb) * 3;
  const real* cv = d.df[B2_DF_CVEL] + ((size_t)w * m.nbody + g) * 6;
  for (int c = 0; c < 3; ++c) b.pos[c] = xp[c];
  for (int c = 0; c < 4; ++c) b.quat[c] = xq[c];
  env_vel_from_cvel(b.lin, b.ang, xp, com, cv);
}
// motion frame of tracked body `sel` at time step t (body_pos_w adds the env origin)
B2_DEV void trk_motion_body(const KTrk& e, int w, long long t, int sel, TrkBody& b) {
  const int nbm = TI(B2T_IP_NBODY_MOTION), mb = TBI(B2T_B_BODY_MOTION)[sel];
  const size_t o = (size_t)t * nbm + mb;
  const real* org = TBR(B2T_B_ENV_ORIGINS) + (size_t)w * 3;
  for (int c = 0; c < 3; ++c) {
    b.pos[c] = TBR(B2T_B_M_BODY_POS)[o * 3 + c] + org[c];
    b.lin[c] = TBR(B2T_B_M_BODY_LIN_VEL)[o * 3 + c];
    b.ang[c] = TBR(B2T_B_M_BODY_ANG_VEL)[o * 3 + c];
  }
  for (int c = 0; c < 4; ++c) b.quat[c] = TBR(B2T_B_M_BODY_QUAT)[o * 4 + c];
}

// ---------------------------------------------------------------------------
B2_DEV void trk_action_world(const KModel& m, const KData& d, const KTrk& e, int w) {
  const int nj = TI(B2T_IP_NJ);
  const int* ctrl_ids = TBI(B2T_B_CTRL_IDS);
  B2_UNROLL1
  for (int j = B2_LANE; j < nj; j += B2_NLANE) {
    size_t o = (size_t)w * nj + j;
    real a_in = TBR(B2T_B_ACTION_IN)[o];
    TBR(B2T_B_PREV_PREV_ACTION)[o] = TBR(B2T_B_PREV_ACTION)[o];
    TBR(B2T_B_PREV_ACTION)[o] = TBR(B2T_B_ACTION)[o];
    TBR(B2T_B_ACTION)[o] = a_in;
    TBR(B2T_B_RAW_ACTION)[o] = a_in;
    real p = a_in * TBR(B2T_B_ACTION_SCALE)[j] + TBR(B2T_B_ACTION_OFFSET)[j];
    TBR(B2T_B_PROCESSED_ACTION)[o] = p;
    real target = p - TBR(B2T_B_ENCODER_BIAS)[o];
    TBR(B2T_B_JOINT_POS_TARGET)[o] = target;
    d.df[B2_DF_CTRL][(size_t)w * m.nu + ctrl_ids[j]] = target;
  }
}

// MotionCommand._resample_command for env w (all lanes call; draws at slot0..slot0+45)
B2_DEVN void trk_resample(const KModel& m, const KData& d, const KTrk& e, int w, long long step, int slot0) {
  const int nj = TI(B2T_IP_NJ), T = TI(B2T_IP_TIME_STEP_TOTAL), nbin = TI(B2T_IP_BIN_COUNT), mode = TI(B2T_IP_SAMPLING_MODE);
  long long t_new = 0;
  if (B2_LANE == 0) {
    const long long t_old = TBL(B2T_B_TIME_STEPS)[w];
    if (mode == 1) {
      t_new = (long long)(trk_draw(e, w, step, slot0) * (real)T);
      if (t_new > T - 1) t_new = T - 1;
    } else if (mode == 2) {
      // failed-bin count of this episode (commands.py:259-268)
      if (TBU(B2T_B_TERMINATED)[w]) {
        long long cur = (t_old * nbin) / (T > 1 ? T : 1);
        cur = cur < 0 ? 0 : (cur > nbin - 1 ? nbin - 1 : cur);
#ifdef B2_HOSTEMU
        TBR(B2T_B_CURRENT_BIN_FAILED)[cur] += 1;
#else
        atomicAdd(reinterpret_cast<float*>(e.buf[B2T_B_CURRENT_BIN_FAILED]) + cur, 1.0f);  // integer-valued: exact in any order
#endif
      }
      // smoothed sampling probabilities and their CDF (commands.py:270-288), inverse-CDF draw
      const real* bfc = TBR(B2T_B_BIN_FAILED_COUNT);
      const int ks = TI(B2T_IP_KERNEL_SIZE);
      const real floor_ = TF(B2T_FP_ADAPTIVE_UNIFORM_RATIO) / (real)nbin, lam = TF(B2T_FP_ADAPTIVE_LAMBDA);
      real ksum = 0, kp = 1;
      for (int k = 0; k < ks; ++k) { ksum += kp; kp *= lam; }
      real total = 0;
      for (int b = 0; b < nbin; ++b) {
        real acc = 0; kp = 1;
        for (int k = 0; k < ks; ++k) { int bb = b + k < nbin ? b + k : nbin - 1; acc += (bfc[bb] + floor_) * (kp / ksum); kp *= lam; }
        total += acc;
      }
      const real u0 = trk_draw(e, w, step, slot0);
      double cdf = 0;
      int bin = nbin - 1;
      for (int b = 0; b < nbin; ++b) {
        real acc = 0; kp = 1;
        for (int k = 0; k < ks; ++k) { int bb = b + k < nbin ? b + k : nbin - 1; acc += (bfc[bb] + floor_) * (kp / ksum); kp *= lam; }
        cdf += (double)(acc / total);
        if (u0 < (real)cdf) { bin = b; break; }
      }
      t_new = (long long)(((real)bin + trk_draw(e, w, step, slot0 + 1)) / (real)nbin * (real)(T - 1));
    }
    TBL(B2T_B_TIME_STEPS)[w] = t_new;
  }
#ifndef B2_HOSTEMU
  t_new = __shfl_sync(0xffffffffu, t_new, 0);
#endif
  real* qp = d.df[B2_DF_QPOS] + (size_t)w * m.nq;
  real* qv = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
  const int* jq = TBI(B2T_B_JOINT_Q_ADR);
  const int* jv = TBI(B2T_B_JOINT_V_ADR);
  B2_UNROLL1
  for (int j = B2_LANE; j < nj; j += B2_NLANE) {
    real p = TBR(B2T_B_M_JOINT_POS)[(size_t)t_new * nj + j] + trk_uniform(e, w, step, slot0 + 14 + j, TF(B2T_FP_JOINT_POS_RANGE), TF(B2T_FP_JOINT_POS_RANGE + 1));
    qp[jq[j]] = clampr(p, TBR(B2T_B_SOFT_LO)[j], TBR(B2T_B_SOFT_HI)[j]);
    qv[jv[j]] = TBR(B2T_B_M_JOINT_VEL)[(size_t)t_new * nj + j];
    TBR(B2T_B_JOINT_POS_TARGET)[(size_t)w * nj + j] = 0;   // robot.clear_state
  }
  if (B2_LANE == 0) {
    TrkBody rt;
    trk_motion_body(e, w, t_new, 0, rt);   // tracked body 0 is the root (commands.py:319-322)
    real ps[6], vs[6];
    for (int k = 0; k < 6; ++k) {
      ps[k] = trk_uniform(e, w, step, slot0 + 2 + k, TF(B2T_FP_POSE_RANGE + 2 * k), TF(B2T_FP_POSE_RANGE + 2 * k + 1));
      vs[k] = trk_uniform(e, w, step, slot0 + 8 + k, TF(B2T_FP_VELOCITY_RANGE + 2 * k), TF(B2T_FP_VELOCITY_RANGE + 2 * k + 1));
    }
    real cr = cos(ps[3] * (real)0.5), sr = sin(ps[3] * (real)0.5), cp = cos(ps[4] * (real)0.5), sp_ = sin(ps[4] * (real)0.5);
    real cy = cos(ps[5] * (real)0.5), sy = sin(ps[5] * (real)0.5);
    real dq[4] = {cy * cr * cp + sy * sr * sp_, cy * sr * cp - sy * cr * sp_, cy * cr * sp_ + sy * sr * cp, sy * cr * cp - cy * sr * sp_};
    real q[4], angw[3], angb[3];
    quat_mul(q, dq, rt.quat);
    for (int c = 0; c < 3; ++c) { qp[c] = rt.pos[c] + ps[c]; qv[c] = rt.lin[c] + vs[c]; angw[c] = rt.ang[c] + vs[3 + c]; }
    for (int c = 0; c < 4; ++c) qp[3 + c] = q[c];
    env_quat_apply_inv(angb, q, angw);
    for (int c = 0; c < 3; ++c) qv[3 + c] = angb[c];
  }
  B2_WSYNC();
}

// _reset_idx for one env
B2_DEVN void trk_reset_world(const KModel& m, const KData& d, const KTrk& e, int w, long long step) {
  const int nj = TI(B2T_IP_NJ);
  if (B2_LANE == 0) {
    real* logv = TBR(B2T_B_LOG_ENV) + (size_t)w * B2T_NLOG;
    real* sums = TBR(B2T_B_EPISODE_SUMS) + (size_t)w * B2T_NREW;
    const unsigned char* td = TBU(B2T_B_TERM_DONES) + (size_t)w * B2T_NTERM;
    real* met = TBR(B2T_B_METRICS) + (size_t)w * B2T_NMETRIC;
    logv[B2T_LOG_NRESET] = 1;
    for (int k = 0; k < B2T_NREW; ++k) { logv[B2T_LOG_EPISODE_SUM + k] = sums[k]; sums[k] = 0; }
    for (int k = 0; k < B2T_NTERM; ++k) logv[B2T_LOG_TERM_COUNT + k] = td[k];
    for (int k = 0; k < B2T_NMETRIC; ++k) { logv[B2T_LOG_METRIC + k] = met[k]; met[k] = 0; }
    TBL(B2T_B_EPISODE_LENGTH)[w] = 0;
  }
  reset_world(m, d, w);
  B2_UNROLL1
  for (int j = B2_LANE; j < nj; j += B2_NLANE) {
    size_t o = (size_t)w * nj + j;
    TBR(B2T_B_ACTION)[o] = 0; TBR(B2T_B_PREV_ACTION)[o] = 0; TBR(B2T_B_PREV_PREV_ACTION)[o] = 0; TBR(B2T_B_RAW_ACTION)[o] = 0;
  }
  trk_resample(m, d, e, w, step, B2T_DRAW_RESET);
}

// ---------------------------------------------------------------------------
// b2_trk_post: counters, terminations, rewards, reset
// ---------------------------------------------------------------------------
B2_DEV void trk_post_world(const KModel& m, const KData& d, const KTrk& e, int w) {
  const int nj = TI(B2T_IP_NJ), nb = TI(B2T_IP_NBODY), asel = TI(B2T_IP_ANCHOR_SEL);
  const long long step = TBL(B2T_B_STEP_COUNTER)[0];
  const long long t = TBL(B2T_B_TIME_STEPS)[w];
  const real dt = TF(B2T_FP_STEP_DT);
  real* logv = TBR(B2T_B_LOG_ENV) + (size_t)w * B2T_NLOG;
  B2_UNROLL1
  for (int i = B2_LANE; i < B2T_NLOG; i += B2_NLANE) logv[i] = 0;
  B2_WSYNC();

  // per tracked body (lanes over bodies): relative-target errors, velocity errors, end-effector z test
  real e_pos = 0, e_ori = 0, e_lin = 0, e_ang = 0;
  int ee_bad = 0;
  B2_UNROLL1
  for (int b = B2_LANE; b < nb; b += B2_NLANE) {
    TrkBody rb, mb;
    trk_robot_body(m, d, e, w, b, rb);
    trk_motion_body(e, w, t, b, mb);
    const real* rp = TBR(B2T_B_REL_POS) + ((size_t)w * nb + b) * 3;
    const real* rq = TBR(B2T_B_REL_QUAT) + ((size_t)w * nb + b) * 4;
    real dp[3] = {rp[0] - rb.pos[0], rp[1] - rb.pos[1], rp[2] - rb.pos[2]};
    e_pos += dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
    real qe = trk_quat_err(rq, rb.quat);
    e_ori += qe * qe;
    real dl[3] = {mb.lin[0] - rb.lin[0], mb.lin[1] - rb.lin[1], mb.lin[2] - rb.lin[2]};
    real da[3] = {mb.ang[0] - rb.ang[0], mb.ang[1] - rb.ang[1], mb.ang[2] - rb.ang[2]};
    e_lin += dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2];
    e_ang += da[0] * da[0] + da[1] * da[1] + da[2] * da[2];
    // bad_motion_body_pos(_z_only): distance (or height difference) of an end-effector body to its motion target
    const real ee_err = (TI(B2T_IP_TERM_FULL3D) & 2) ? sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) : fabs(dp[2]);
    for (int k = 0; k < TI(B2T_IP_NEE); ++k)
      if (TI(B2T_IP_EE_SEL + k) == b && ee_err > TF(B2T_FP_EE_POS_THRESHOLD)) ee_bad = 1;
  }
  e_pos = b2_wsum(e_pos) / (real)nb; e_ori = b2_wsum(e_ori) / (real)nb; e_lin = b2_wsum(e_lin) / (real)nb; e_ang = b2_wsum(e_ang) / (real)nb;
  const bool ee = b2_wor_i(ee_bad) != 0;

  // per joint: action rate, soft-limit violation
  real rate_acc = 0, lim_acc = 0;
  const real* qpos = d.df[B2_DF_QPOS] + (size_t)w * m.nq;
  const int* jq = TBI(B2T_B_JOINT_Q_ADR);
  B2_UNROLL1
  for (int j = B2_LANE; j < nj; j += B2_NLANE) {
    size_t o = (size_t)w * nj + j;
    real da = TBR(B2T_B_ACTION)[o] - TBR(B2T_B_PREV_ACTION)[o];
    rate_acc += da * da;
    real q = qpos[jq[j]];
    real lo = q - TBR(B2T_B_SOFT_LO)[j], hi = q - TBR(B2T_B_SOFT_HI)[j];
    lim_acc += -(lo < 0 ? lo : (real)0) + (hi > 0 ? hi : (real)0);
  }
  rate_acc = b2_wsum(rate_acc); lim_acc = b2_wsum(lim_acc);

  // anchor terms (every lane computes them: uniform control flow for the reset below)
  TrkBody ra, ma;
  trk_robot_body(m, d, e, w, asel, ra);
  trk_motion_body(e, w, t, asel, ma);
  const real g[3] = {0, 0, -1};
  real pgm[3], pgr[3];
  env_quat_apply_inv(pgm, ma.quat, g);
  env_quat_apply_inv(pgr, ra.quat, g);
  const long long ep = TBL(B2T_B_EPISODE_LENGTH)[w] + 1;
  const bool time_out = ep >= (long long)TI(B2T_IP_MAX_EPISODE_LENGTH);
  const real adx = ma.pos[0] - ra.pos[0], ady = ma.pos[1] - ra.pos[1], adz = ma.pos[2] - ra.pos[2];
  const bool apos = ((TI(B2T_IP_TERM_FULL3D) & 1) ? sqrt(adx * adx + ady * ady + adz * adz) : fabs(adz)) > TF(B2T_FP_ANCHOR_POS_THRESHOLD);
  const bool aori = fabs(pgm[2] - pgr[2]) > TF(B2T_FP_ANCHOR_ORI_THRESHOLD);
  const bool terminated = apos || aori || ee;
  const bool reset = terminated || time_out;

  if (B2_LANE == 0) {
    real term[B2T_NREW];
    real dp[3] = {ma.pos[0] - ra.pos[0], ma.pos[1] - ra.pos[1], ma.pos[2] - ra.pos[2]};
    real s0 = TF(B2T_FP_STD), s1 = TF(B2T_FP_STD + 1), s2 = TF(B2T_FP_STD + 2), s3 = TF(B2T_FP_STD + 3), s4 = TF(B2T_FP_STD + 4), s5 = TF(B2T_FP_STD + 5);
    term[B2T_REW_ANCHOR_POS] = exp(-(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) / (s0 * s0));
    real qe = trk_quat_err(ma.quat, ra.quat);
    term[B2T_REW_ANCHOR_ORI] = exp(-(qe * qe) / (s1 * s1));
    term[B2T_REW_BODY_POS] = exp(-e_pos / (s2 * s2));
    term[B2T_REW_BODY_ORI] = exp(-e_ori / (s3 * s3));
    term[B2T_REW_BODY_LIN_VEL] = exp(-e_lin / (s4 * s4));
    term[B2T_REW_BODY_ANG_VEL] = exp(-e_ang / (s5 * s5));
    term[B2T_REW_ACTION_RATE] = rate_acc;
    term[B2T_REW_JOINT_LIMIT] = lim_acc;
    term[B2T_REW_SELF_COLLISIONS] = TI(B2T_IP_SELFCOL_ADR) >= 0 ? d.df[B2_DF_SENSORDATA][(size_t)w * m.nsensordata + TI(B2T_IP_SELFCOL_ADR)] : (real)0;
    real total = 0;
    real* step_rew = TBR(B2T_B_STEP_REWARD) + (size_t)w * B2T_NREW;
    real* sums = TBR(B2T_B_EPISODE_SUMS) + (size_t)w * B2T_NREW;
    for (int k = 0; k < B2T_NREW; ++k) {
      real wk = TF(B2T_FP_REWARD_WEIGHT + k);
      if (wk == 0) { step_rew[k] = 0; continue; }
      real v = env_nan_to_num(term[k] * wk * dt);
      total += v; sums[k] += v; step_rew[k] = v / dt;
    }
    TBR(B2T_B_REWARD)[w] = total;
    unsigned char* td = TBU(B2T_B_TERM_DONES) + (size_t)w * B2T_NTERM;
    td[B2T_TERM_TIME_OUT] = time_out; td[B2T_TERM_ANCHOR_POS] = apos; td[B2T_TERM_ANCHOR_ORI] = aori; td[B2T_TERM_EE_BODY_POS] = ee;
    TBU(B2T_B_TERMINATED)[w] = terminated; TBU(B2T_B_TIME_OUTS)[w] = time_out; TBU(B2T_B_RESET_BUF)[w] = reset;
    TBL(B2T_B_EPISODE_LENGTH)[w] = ep;
  }
  B2_WSYNC();
  if (reset) trk_reset_world(m, d, e, w, step);
}

// ---------------------------------------------------------------------------
// b2_trk_obs: metrics, frame advance + relative targets, push event, observations
// ---------------------------------------------------------------------------
B2_DEV void trk_obs_world(const KModel& m, const KData& d, const KTrk& e, int w, int with_update) {
  const int nj = TI(B2T_IP_NJ), nb = TI(B2T_IP_NBODY), asel = TI(B2T_IP_ANCHOR_SEL), T = TI(B2T_IP_TIME_STEP_TOTAL);
  const long long step = TBL(B2T_B_STEP_COUNTER)[0];
  const real dt = TF(B2T_FP_STEP_DT);
  const int* jq = TBI(B2T_B_JOINT_Q_ADR);
  const int* jv = TBI(B2T_B_JOINT_V_ADR);
  real* qpos = d.df[B2_DF_QPOS] + (size_t)w * m.nq;
  real* qvel = d.df[B2_DF_QVEL] + (size_t)w * m.nv;
  TrkBody ra;
  trk_robot_body(m, d, e, w, asel, ra);
  long long t = TBL(B2T_B_TIME_STEPS)[w];

  if (with_update) {
    // ---- _update_metrics (commands.py:219-255), with the frame and relative targets of the previous update ----
    real mb_pos = 0, mb_rot = 0, mj_pos = 0, mj_vel = 0;
    B2_UNROLL1
    for (int b = B2_LANE; b < nb; b += B2_NLANE) {
      TrkBody rb;
      trk_robot_body(m, d, e, w, b, rb);
      const real* rp = TBR(B2T_B_REL_POS) + ((size_t)w * nb + b) * 3;
      const real* rq = TBR(B2T_B_REL_QUAT) + ((size_t)w * nb + b) * 4;
      real dp[3] = {rp[0] - rb.pos[0], rp[1] - rb.pos[1], rp[2] - rb.pos[2]};
      mb_pos += sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
      mb_rot += trk_quat_err(rq, rb.quat);
    }
    B2_UNROLL1
    for (int j = B2_LANE; j < nj; j += B2_NLANE) {
      real dp = TBR(B2T_B_M_JOINT_POS)[(size_t)t * nj + j] - qpos[jq[j]];
      real dv = TBR(B2T_B_M_JOINT_VEL)[(size_t)t * nj + j] - qvel[jv[j]];
      mj_pos += dp * dp; mj_vel += dv * dv;
    }
    mb_pos = b2_wsum(mb_pos) / (real)nb; mb_rot = b2_wsum(mb_rot) / (real)nb; mj_pos = sqrt(b2_wsum(mj_pos)); mj_vel = sqrt(b2_wsum(mj_vel));
    if (B2_LANE == 0) {
      TrkBody ma;
      trk_motion_body(e, w, t, asel, ma);
      real* met = TBR(B2T_B_METRICS) + (size_t)w * B2T_NMETRIC;
      real dp[3] = {ma.pos[0] - ra.pos[0], ma.pos[1] - ra.pos[1], ma.pos[2] - ra.pos[2]};
      real dl[3] = {ma.lin[0] - ra.lin[0], ma.lin[1] - ra.lin[1], ma.lin[2] - ra.lin[2]};
      real da[3] = {ma.ang[0] - ra.ang[0], ma.ang[1] - ra.ang[1], ma.ang[2] - ra.ang[2]};
      met[0] = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
      met[1] = trk_quat_err(ma.quat, ra.quat);
      met[2] = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
      met[3] = sqrt(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]);
      met[4] = mb_pos; met[5] = mb_rot; met[6] = mj_pos; met[7] = mj_vel;
    }
    B2_WSYNC();
    // ---- _update_command (commands.py:379-412): advance the frame, resample at the end of the clip ----
    t += 1;
    if (B2_LANE == 0) TBL(B2T_B_TIME_STEPS)[w] = t;
    B2_WSYNC();
    if (t >= T) {
      trk_resample(m, d, e, w, step, B2T_DRAW_CMD);
      t = TBL(B2T_B_TIME_STEPS)[w];
    }
    // anchor-relative targets: yaw-only rotation from the motion anchor to the robot anchor, xy from the robot, z from the motion
    TrkBody ma;
    trk_motion_body(e, w, t, asel, ma);
    real inv[4], dq[4], dyaw[4];
    trk_quat_inv(inv, ma.quat);
    quat_mul(dq, ra.quat, inv);
    trk_yaw_quat(dyaw, dq);
    B2_UNROLL1
    for (int b = B2_LANE; b < nb; b += B2_NLANE) {
      TrkBody mb;
      trk_motion_body(e, w, t, b, mb);
      real* rp = TBR(B2T_B_REL_POS) + ((size_t)w * nb + b) * 3;
      real* rq = TBR(B2T_B_REL_QUAT) + ((size_t)w * nb + b) * 4;
      real q[4], off[3] = {mb.pos[0] - ma.pos[0], mb.pos[1] - ma.pos[1], mb.pos[2] - ma.pos[2]}, r[3];
      quat_mul(q, dyaw, mb.quat);
      env_quat_apply(r, dyaw, off);
      rq[0] = q[0]; rq[1] = q[1]; rq[2] = q[2]; rq[3] = q[3];
      rp[0] = ra.pos[0] + r[0]; rp[1] = ra.pos[1] + r[1]; rp[2] = ma.pos[2] + r[2];
    }
    // ---- interval event: push (tracking_env_cfg.py:173-178) ----
    if (B2_LANE == 0 && TI(B2T_IP_PUSH_ENABLED)) {
      real pt = TBR(B2T_B_PUSH_TIME_LEFT)[w] - dt;
      if (pt < (real)1e-6) {
        pt = trk_uniform(e, w, step, B2T_DRAW_PUSH, TF(B2T_FP_PUSH_INTERVAL), TF(B2T_FP_PUSH_INTERVAL + 1));
        const int rb = TI(B2T_IP_ROOT_BODY);
        real lin[3], ang[3], vel[6], angb[3];
        env_vel_from_cvel(lin, ang, d.df[B2_DF_XPOS] + ((size_t)w * m.nbody + rb) * 3, d.df[B2_DF_SUBTREE_COM] + ((size_t)w * m.nbody + rb) * 3,
                          d.df[B2_DF_CVEL] + ((size_t)w * m.nbody + rb) * 6);
        for (int k = 0; k < 3; ++k) {
          vel[k] = lin[k] + trk_uniform(e, w, step, B2T_DRAW_PUSH + 1 + k, TF(B2T_FP_PUSH_VEL + 2 * k), TF(B2T_FP_PUSH_VEL + 2 * k + 1));
          vel[3 + k] = ang[k] + trk_uniform(e, w, step, B2T_DRAW_PUSH + 4 + k, TF(B2T_FP_PUSH_VEL + 6 + 2 * k), TF(B2T_FP_PUSH_VEL + 6 + 2 * k + 1));
        }
        env_quat_apply_inv(angb, qpos + 3, vel + 3);
        for (int c = 0; c < 3; ++c) { qvel[c] = vel[c]; qvel[3 + c] = angb[c]; }
      }
      TBR(B2T_B_PUSH_TIME_LEFT)[w] = pt;
    }
    B2_WSYNC();
  }

  // ---- observations (tracking_env_cfg.py:51-122) ----------------------------------------------------------
  const int np = TI(B2T_IP_N_OBS_POLICY), nc = TI(B2T_IP_N_OBS_CRITIC);
  real* op = TBR(B2T_B_OBS_POLICY) + (size_t)w * np;
  real* oc = TBR(B2T_B_OBS_CRITIC) + (size_t)w * nc;
  const real* sd = d.df[B2_DF_SENSORDATA] + (size_t)w * m.nsensordata;
  TrkBody ma;
  trk_motion_body(e, w, t, asel, ma);
  real q10[4], apos_b[3], aori[4], a6[6], dpos[3] = {ma.pos[0] - ra.pos[0], ma.pos[1] - ra.pos[1], ma.pos[2] - ra.pos[2]};
  trk_quat_inv(q10, ra.quat);
  env_quat_apply(apos_b, q10, dpos);
  quat_mul(aori, q10, ma.quat);
  trk_mat6(a6, aori);
  const bool corrupt = TI(B2T_IP_ENABLE_CORRUPTION) != 0;
  // layout: policy = [cmd 2nj | anchor_pos 3 | anchor_ori 6 | lin 3 | ang 3 | jpos nj | jvel nj | act nj]
  //         critic = [cmd 2nj | anchor_pos 3 | anchor_ori 6 | body_pos 3nb | body_ori 6nb | lin 3 | ang 3 | jpos nj | jvel nj | act nj]
  const int c0 = 2 * nj + 9, cb = c0 + 9 * nb;
  B2_UNROLL1
  for (int i = B2_LANE; i < nc; i += B2_NLANE) {
    real v;
    int pi = -1;   // index of the same entry in the policy group (-1: critic only)
    bool biased = false;
    if (i < 2 * nj) { v = i < nj ? TBR(B2T_B_M_JOINT_POS)[(size_t)t * nj + i] : TBR(B2T_B_M_JOINT_VEL)[(size_t)t * nj + i - nj]; pi = i; }
    else if (i < 2 * nj + 3) { v = apos_b[i - 2 * nj]; pi = i; }
    else if (i < c0) { v = a6[i - 2 * nj - 3]; pi = i; }
    else if (i < c0 + 3 * nb) {
      int b = (i - c0) / 3, c = (i - c0) % 3;
      TrkBody rb;
      trk_robot_body(m, d, e, w, b, rb);
      real off[3] = {rb.pos[0] - ra.pos[0], rb.pos[1] - ra.pos[1], rb.pos[2] - ra.pos[2]}, r[3];
      env_quat_apply(r, q10, off);
      v = r[c];
    } else if (i < cb) {
      int b = (i - c0 - 3 * nb) / 6, c = (i - c0 - 3 * nb) % 6;
      TrkBody rb;
      trk_robot_body(m, d, e, w, b, rb);
      real q[4], o6[6];
      quat_mul(q, q10, rb.quat);
      trk_mat6(o6, q);
      v = o6[c];
    } else {
      int k = i - cb;   // [lin 3 | ang 3 | jpos nj | jvel nj | act nj]
      pi = c0 + k;
      if (k < 3) v = sd[TI(B2T_IP_LINVEL_ADR) + k];
      else if (k < 6) v = sd[TI(B2T_IP_ANGVEL_ADR) + k - 3];
      else if (k < 6 + nj) { int j = k - 6; v = qpos[jq[j]] - TBR(B2T_B_DEFAULT_JOINT_POS)[j]; biased = true; }
      else if (k < 6 + 2 * nj) { int j = k - 6 - nj; v = qvel[jv[j]]; }
      else v = TBR(B2T_B_ACTION)[(size_t)w * nj + k - 6 - 2 * nj];
    }
    oc[i] = v;
    if (pi >= 0) {
      if (biased) v += TBR(B2T_B_ENCODER_BIAS)[(size_t)w * nj + (i - cb - 6)];   // joint_pos_rel(biased=True) in the policy group
      if (corrupt) {
        real lo = TBR(B2T_B_NOISE_LO)[pi], hi = TBR(B2T_B_NOISE_HI)[pi];
        if (lo != hi) v = v + (trk_draw(e, w, step, B2T_DRAW_NOISE + pi) * (hi - lo) + lo);
      }
      op[pi] = v;
    }
  }
  B2_WSYNC();
}
