/* b200trk.h - C ABI of the fused manager-based env step of the motion-tracking task (BASELINE config 5).
 *
 * Same role as b200env.h for `Mjlab-Tracking-Flat-Unitree-G1` (reference: src/mjlab/tasks/tracking/tracking_env_cfg.py:44-317,
 * config/g1/env_cfgs.py:15-100):
 *
 *   ActionManager / JointPositionAction (+ encoder bias)     envs/mdp/actions/joint_actions.py:84-104       -> b2_trk_action
 *   TerminationManager (time_out, anchor_pos z, anchor_ori, ee_body_pos z)  tasks/tracking/mdp/terminations.py:28-86,
 *   RewardManager (6 motion terms, action_rate, joint_limit, self_collisions) tasks/tracking/mdp/rewards.py:26-120,
 *   _reset_idx + MotionCommand._resample_command (start / uniform / adaptive bin sampling, perturbed motion state
 *     written to qpos / qvel)                                 tasks/tracking/mdp/commands.py:257-377        -> b2_trk_post / b2_trk_reset
 *   MotionCommand._update_metrics / _update_command (frame advance, anchor-relative body targets),
 *   push event, ObservationManager (policy 160, critic 286 for G1)  commands.py:219-255,379-412,
 *                                                             tasks/tracking/mdp/observations.py:17-69      -> b2_trk_obs
 *   log sums + the adaptive sampler's EMA of failed bins      commands.py:405-410                           -> b2_trk_log
 *
 * Conventions as in b200env.h (caller-owned device buffers, stream-ordered, graph-capturable, Philox draws).
 * The motion clip lives on the device in the reference's npz layout (all robot bodies; the tracked subset is an index table).
 */
#ifndef B200TRK_H_
#define B200TRK_H_

#include "b200sim.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B2T_NREW 9
#define B2T_NTERM 4
#define B2T_NMETRIC 8
#define B2T_MAXEE 8
#define B2T_MAXBODY 32
#define B2T_MAXKERNEL 8

enum b2t_rew { B2T_REW_ANCHOR_POS = 0, B2T_REW_ANCHOR_ORI, B2T_REW_BODY_POS, B2T_REW_BODY_ORI, B2T_REW_BODY_LIN_VEL,
               B2T_REW_BODY_ANG_VEL, B2T_REW_ACTION_RATE, B2T_REW_JOINT_LIMIT, B2T_REW_SELF_COLLISIONS };
enum b2t_term { B2T_TERM_TIME_OUT = 0, B2T_TERM_ANCHOR_POS, B2T_TERM_ANCHOR_ORI, B2T_TERM_EE_BODY_POS };

enum b2t_ip {
  B2T_IP_NJ = 0,
  B2T_IP_NBODY,             /* tracked bodies */
  B2T_IP_NBODY_MOTION,      /* bodies per frame in the motion arrays */
  B2T_IP_ROOT_BODY,         /* global id of the robot's root body */
  B2T_IP_ANCHOR_SEL,        /* index of the anchor body inside the tracked list */
  B2T_IP_N_OBS_POLICY, B2T_IP_N_OBS_CRITIC,
  B2T_IP_LINVEL_ADR, B2T_IP_ANGVEL_ADR, B2T_IP_SELFCOL_ADR,
  B2T_IP_MAX_EPISODE_LENGTH,
  B2T_IP_ENABLE_CORRUPTION, B2T_IP_PUSH_ENABLED,
  B2T_IP_SAMPLING_MODE,     /* 0 start, 1 uniform, 2 adaptive */
  B2T_IP_TIME_STEP_TOTAL, B2T_IP_BIN_COUNT, B2T_IP_KERNEL_SIZE,
  B2T_IP_TERM_FULL3D,       /* bit 0: bad_anchor_pos (3-D distance) instead of bad_anchor_pos_z_only; bit 1: bad_motion_body_pos
                             * instead of bad_motion_body_pos_z_only (tasks/tracking/mdp/terminations.py:18-86) */
  B2T_IP_NEE,
  B2T_IP_EE_SEL,            /* [B2T_MAXEE] indices inside the tracked list */
  B2T_IP_COUNT = B2T_IP_EE_SEL + B2T_MAXEE
};

enum b2t_fp {
  B2T_FP_STEP_DT = 0, B2T_FP_MAX_EPISODE_LENGTH_S,
  B2T_FP_ANCHOR_POS_THRESHOLD, B2T_FP_ANCHOR_ORI_THRESHOLD, B2T_FP_EE_POS_THRESHOLD,
  B2T_FP_STD,               /* [6]: anchor pos, anchor ori, body pos, body ori, body lin vel, body ang vel */
  B2T_FP_ADAPTIVE_LAMBDA = B2T_FP_STD + 6, B2T_FP_ADAPTIVE_UNIFORM_RATIO, B2T_FP_ADAPTIVE_ALPHA,
  B2T_FP_POSE_RANGE,        /* [6][2] */
  B2T_FP_VELOCITY_RANGE = B2T_FP_POSE_RANGE + 12,   /* [6][2] */
  B2T_FP_JOINT_POS_RANGE = B2T_FP_VELOCITY_RANGE + 12,  /* [2] */
  B2T_FP_PUSH_INTERVAL = B2T_FP_JOINT_POS_RANGE + 2,     /* [2] */
  B2T_FP_PUSH_VEL = B2T_FP_PUSH_INTERVAL + 2,            /* [6][2] */
  B2T_FP_REWARD_WEIGHT = B2T_FP_PUSH_VEL + 12,           /* [B2T_NREW] */
  B2T_FP_COUNT = B2T_FP_REWARD_WEIGHT + B2T_NREW
};

enum b2t_buf {
  /* tables */
  B2T_B_JOINT_Q_ADR = 0, B2T_B_JOINT_V_ADR, B2T_B_CTRL_IDS,      /* int [nj] */
  B2T_B_DEFAULT_JOINT_POS, B2T_B_SOFT_LO, B2T_B_SOFT_HI, B2T_B_ACTION_SCALE, B2T_B_ACTION_OFFSET,   /* [nj] */
  B2T_B_NOISE_LO, B2T_B_NOISE_HI,                                  /* [n_obs_policy] */
  B2T_B_BODY_GLOBAL,        /* int [nbody]: global body id of each tracked body */
  B2T_B_BODY_MOTION,        /* int [nbody]: index of each tracked body in the motion arrays */
  /* motion clip, [T, ...] */
  B2T_B_M_JOINT_POS, B2T_B_M_JOINT_VEL,                            /* [nj] */
  B2T_B_M_BODY_POS, B2T_B_M_BODY_QUAT, B2T_B_M_BODY_LIN_VEL, B2T_B_M_BODY_ANG_VEL,   /* [nbody_motion, 3|4] */
  /* per-env state */
  B2T_B_ENV_ORIGINS,        /* [3] */
  B2T_B_ACTION_IN, B2T_B_ACTION, B2T_B_PREV_ACTION, B2T_B_PREV_PREV_ACTION, B2T_B_RAW_ACTION, B2T_B_PROCESSED_ACTION,
  B2T_B_ENCODER_BIAS, B2T_B_JOINT_POS_TARGET,                      /* [nj] */
  B2T_B_TIME_STEPS,         /* int64 [1] */
  B2T_B_REL_POS, B2T_B_REL_QUAT,                                   /* [nbody,3], [nbody,4] body_*_relative_w */
  B2T_B_METRICS,            /* [B2T_NMETRIC] */
  B2T_B_PUSH_TIME_LEFT,
  B2T_B_EPISODE_LENGTH,     /* int64 [1] */
  B2T_B_REWARD, B2T_B_STEP_REWARD, B2T_B_EPISODE_SUMS,             /* [1], [B2T_NREW], [B2T_NREW] */
  B2T_B_TERM_DONES,         /* uint8 [B2T_NTERM] */
  B2T_B_TERMINATED, B2T_B_TIME_OUTS, B2T_B_RESET_BUF,              /* uint8 [1] */
  B2T_B_OBS_POLICY, B2T_B_OBS_CRITIC,
  B2T_B_LOG_ENV,            /* [B2T_NLOG] */
  /* scalars */
  B2T_B_BIN_FAILED_COUNT, B2T_B_CURRENT_BIN_FAILED,                /* [bin_count] */
  B2T_B_LOG,                /* [B2T_NLOG] */
  B2T_B_STEP_COUNTER,       /* int64 [1] */
  B2T_B_ANY_RESET,          /* int [1] */
  B2T_B_INJECT_U,           /* optional [nworld, B2T_NDRAW] */
  B2T_B_COUNT
};

enum b2t_log {
  B2T_LOG_NRESET = 0,
  B2T_LOG_EPISODE_SUM,                                   /* [B2T_NREW] */
  B2T_LOG_TERM_COUNT = B2T_LOG_EPISODE_SUM + B2T_NREW,   /* [B2T_NTERM] */
  B2T_LOG_METRIC = B2T_LOG_TERM_COUNT + B2T_NTERM,       /* [B2T_NMETRIC] sums over reset envs */
  B2T_NLOG = B2T_LOG_METRIC + B2T_NMETRIC
};

enum b2t_draw {
  B2T_DRAW_RESET = 0,      /* [46]: bin, fraction, pose[6], velocity[6], joint[32] */
  B2T_DRAW_CMD = 46,       /* [46] same layout, for the resample at the end of the clip */
  B2T_DRAW_PUSH = 92,      /* [7]: interval, vel[6] */
  B2T_DRAW_NOISE = 100,    /* [n_obs_policy <= 192] */
  B2T_NDRAW = 292
};

typedef struct b2_trk {
  int ip[B2T_IP_COUNT];
  float fp[B2T_FP_COUNT];
  unsigned long long seed;
  void* buf[B2T_B_COUNT];
} b2_trk;

int b2_sizeof_trk(void);
int b2_trk_check(const b2_model* m, const b2_data* d, const b2_trk* e);
int b2_trk_action(const b2_model* m, b2_data* d, const b2_trk* e, void* stream);
int b2_trk_post(const b2_model* m, b2_data* d, const b2_trk* e, void* stream);
int b2_trk_reset(const b2_model* m, b2_data* d, const b2_trk* e, const unsigned char* mask, void* stream);
int b2_trk_obs(const b2_model* m, b2_data* d, const b2_trk* e, int with_update, void* stream);
int b2_trk_log(const b2_data* d, const b2_trk* e, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TRK_H_ */
