/* b200env.h - C ABI of the fused manager-based env step (velocity task) and the actuator ops.
 *
 * These entry points replace, for the velocity-tracking tasks (G1 / Go1), the torch code that
 * mjlab's managers run around the physics step (reference: src/mjlab/envs/manager_based_rl_env.py:272-313):
 *
 *   ActionManager.process_action / apply_action     managers/action_manager.py:113-130,
 *     JointPositionAction                            envs/mdp/actions/joint_actions.py:84-104,
 *     BuiltinActuatorGroup.apply_controls            actuator/builtin_group.py:68-81      -> b2_env_action
 *   ContactSensor._update_air_time_tracking          sensor/contact_sensor.py:327-367      -> b2_env_airtime
 *   TerminationManager.compute                       managers/termination_manager.py:87-97,
 *   RewardManager.compute                            managers/reward_manager.py:77-91,
 *   ManagerBasedRlEnv._reset_idx                     envs/manager_based_rl_env.py:381-416
 *     (sim.reset, scene.reset, reset events, manager resets)                               -> b2_env_post
 *   CommandManager.compute                           managers/command_manager.py:53-67,
 *   EventManager.apply(mode="interval")              managers/event_manager.py:124-146,
 *   ObservationManager.compute                       managers/observation_manager.py:154-208 -> b2_env_obs
 *   per-iteration log means (extras["log"])          reward_manager.py:59-75 etc.           -> b2_env_log
 *   IdealPdActuator / DcMotorActuator.compute        actuator/pd_actuator.py:94-109,
 *                                                    actuator/dc_actuator.py:136-162        -> b2_actuator_pd
 *   DelayBuffer (append + lagged read)               utils/buffers/delay_buffer.py:229-303  -> b2_delay_step
 *   LearnedMlpActuator.compute                       actuator/learned_actuator.py:144-207   -> b2_actuator_mlp
 *
 * Same conventions as b200sim.h: plain pointers and sizes, caller-owned device buffers with stable
 * addresses, every call only enqueues kernels on the caller's stream (capturable in a CUDA graph).
 * Random draws come from a counter-based Philox4x32-10 generator keyed by (seed, env id) and indexed by
 * (env-step counter, draw slot), so the whole step is a pure function of its inputs.
 */
#ifndef B200ENV_H_
#define B200ENV_H_

#include "b200sim.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B2E_MAXFEET 4
#define B2E_NREW 22
#define B2E_NTERM 5

/* reward term order (tasks/velocity/velocity_env_cfg.py:183-278 + config/g1/env_cfgs.py:122-126) */
enum b2e_rew {
  B2E_REW_TRACK_LIN = 0, B2E_REW_TRACK_ANG, B2E_REW_UPRIGHT, B2E_REW_POSE, B2E_REW_BODY_ANG_VEL, B2E_REW_ANGMOM,
  B2E_REW_DOF_POS_LIMITS, B2E_REW_ACTION_RATE, B2E_REW_AIR_TIME, B2E_REW_FOOT_CLEARANCE, B2E_REW_FOOT_SWING,
  B2E_REW_FOOT_SLIP, B2E_REW_SOFT_LANDING, B2E_REW_SELF_COLLISIONS,
  /* the generic term library (envs/mdp/rewards.py:22-70,125-155), weight 0 in the shipped velocity tasks */
  B2E_REW_IS_ALIVE, B2E_REW_IS_TERMINATED, B2E_REW_JOINT_TORQUES_L2, B2E_REW_JOINT_VEL_L2, B2E_REW_JOINT_ACC_L2,
  B2E_REW_ACTION_ACC_L2, B2E_REW_FLAT_ORIENTATION_L2,
  B2E_REW_ELECTRICAL_POWER   /* electrical_power_cost (envs/mdp/rewards.py:125-146): sum_j max(0, actuator_force_j * joint_vel_j) */
};
/* termination term order; the last two (envs/mdp/terminations.py:35-47) are off unless B2E_IP_TERM_FLAGS enables them */
enum b2e_term { B2E_TERM_TIME_OUT = 0, B2E_TERM_FELL_OVER, B2E_TERM_ILLEGAL_CONTACT, B2E_TERM_ROOT_HEIGHT, B2E_TERM_NAN };

/* ---- int parameters ----------------------------------------------------- */
enum b2e_ip {
  B2E_IP_NJ = 0,            /* actuated joints */
  B2E_IP_NFEET,
  B2E_IP_ROOT_BODY,
  B2E_IP_UPRIGHT_BODY,      /* body used by the upright / body_ang_vel terms */
  B2E_IP_N_OBS_POLICY,
  B2E_IP_N_OBS_CRITIC,
  B2E_IP_LINVEL_ADR, B2E_IP_ANGVEL_ADR, B2E_IP_ANGMOM_ADR,   /* sensordata addresses */
  B2E_IP_SELFCOL_ADR,       /* -1: no self-collision sensor */
  B2E_IP_ILLEGAL_ADR, B2E_IP_ILLEGAL_COUNT,                  /* illegal_contact termination (Go1), count 0 = off */
  B2E_IP_MAX_EPISODE_LENGTH,
  B2E_IP_ENABLE_CORRUPTION, /* policy-group observation noise */
  B2E_IP_PUSH_ENABLED,
  B2E_IP_HEADING_COMMAND,
  B2E_IP_TERRAIN_MODE,      /* bit 0: terrain_levels_vel curriculum at reset (tasks/velocity/mdp/curriculums.py:30-64);
                               bit 1: randomize_terrain reset event of play mode (envs/mdp/events.py:26-37); 0: flat plane */
  B2E_IP_TERRAIN_ROWS, B2E_IP_TERRAIN_COLS,
  B2E_IP_TERM_FLAGS,        /* bit 0: root_height_below_minimum (B2E_FP_MIN_ROOT_HEIGHT), bit 1: nan_detection */
  B2E_IP_FOOT_SITE,         /* [B2E_MAXFEET] global site ids */
  B2E_IP_FOOT_SITE_BODY = B2E_IP_FOOT_SITE + B2E_MAXFEET,    /* [B2E_MAXFEET] body of each foot site */
  B2E_IP_FOOT_FOUND_ADR = B2E_IP_FOOT_SITE_BODY + B2E_MAXFEET,
  B2E_IP_FOOT_FORCE_ADR = B2E_IP_FOOT_FOUND_ADR + B2E_MAXFEET,
  B2E_IP_COUNT = B2E_IP_FOOT_FORCE_ADR + B2E_MAXFEET
};

/* ---- float parameters ---------------------------------------------------- */
enum b2e_fp {
  B2E_FP_STEP_DT = 0,
  B2E_FP_MAX_EPISODE_LENGTH_S,
  B2E_FP_LIMIT_ANGLE,
  B2E_FP_STD_LIN, B2E_FP_STD_ANG, B2E_FP_STD_UPRIGHT,
  B2E_FP_WALKING_THRESHOLD, B2E_FP_RUNNING_THRESHOLD,
  B2E_FP_AIR_THR_MIN, B2E_FP_AIR_THR_MAX, B2E_FP_AIR_CMD_THRESHOLD,
  B2E_FP_TARGET_HEIGHT, B2E_FP_CMD_THRESHOLD,
  B2E_FP_MIN_ROOT_HEIGHT,   /* root_height_below_minimum(minimum_height) */
  B2E_FP_TERRAIN_HALF_SIZE, /* terrain_generator.size[0] / 2: an env that walked farther than this moves up a level */
  B2E_FP_CMD_RESAMPLE,      /* [2] */
  B2E_FP_CMD_LIN_X = B2E_FP_CMD_RESAMPLE + 2,
  B2E_FP_CMD_LIN_Y = B2E_FP_CMD_LIN_X + 2,
  B2E_FP_CMD_ANG_Z = B2E_FP_CMD_LIN_Y + 2,
  B2E_FP_CMD_HEADING = B2E_FP_CMD_ANG_Z + 2,
  B2E_FP_REL_STANDING = B2E_FP_CMD_HEADING + 2,
  B2E_FP_REL_HEADING,
  B2E_FP_HEADING_STIFFNESS,
  B2E_FP_RESET_POSE,        /* [6][2]: x y z roll pitch yaw */
  B2E_FP_RESET_VEL = B2E_FP_RESET_POSE + 12,       /* [6][2] */
  B2E_FP_RESET_JOINT_POS = B2E_FP_RESET_VEL + 12,  /* [2] */
  B2E_FP_RESET_JOINT_VEL = B2E_FP_RESET_JOINT_POS + 2,
  B2E_FP_PUSH_INTERVAL = B2E_FP_RESET_JOINT_VEL + 2, /* [2] */
  B2E_FP_PUSH_VEL = B2E_FP_PUSH_INTERVAL + 2,      /* [6][2] */
  B2E_FP_REWARD_WEIGHT = B2E_FP_PUSH_VEL + 12,     /* [B2E_NREW] */
  B2E_FP_COUNT = B2E_FP_REWARD_WEIGHT + B2E_NREW
};

/* ---- device buffers ------------------------------------------------------- */
enum b2e_buf {
  /* per-joint tables, [nj] */
  B2E_B_JOINT_Q_ADR = 0,    /* int */
  B2E_B_JOINT_V_ADR,        /* int */
  B2E_B_CTRL_IDS,           /* int */
  B2E_B_DEFAULT_JOINT_POS, B2E_B_DEFAULT_JOINT_VEL, B2E_B_SOFT_LO, B2E_B_SOFT_HI,
  B2E_B_ACTION_SCALE, B2E_B_ACTION_OFFSET, B2E_B_STD_STANDING, B2E_B_STD_WALKING, B2E_B_STD_RUNNING,
  B2E_B_NOISE_LO, B2E_B_NOISE_HI,          /* [n_obs_policy] */
  B2E_B_DEFAULT_ROOT_STATE, /* [13] */
  B2E_B_TERRAIN_ORIGINS,    /* [rows, cols, 3] spawn origin of every sub-terrain (terrain_importer.py:176-190); [3] dummy when flat */
  /* per-env state, [nworld, ...] float unless noted */
  B2E_B_ENV_ORIGINS,        /* [3] */
  B2E_B_TERRAIN_LEVEL, B2E_B_TERRAIN_TYPE,   /* int [1]: row (difficulty) and column (type) of the env's sub-terrain */
  B2E_B_ACTION_IN,          /* [nj] the policy's action for this step (input) */
  B2E_B_ACTION, B2E_B_PREV_ACTION, B2E_B_PREV_PREV_ACTION,  /* [nj] ActionManager history */
  B2E_B_RAW_ACTION, B2E_B_PROCESSED_ACTION,                  /* [nj] JointAction */
  B2E_B_ENCODER_BIAS,       /* [nj] */
  B2E_B_JOINT_POS_TARGET,   /* [nj] */
  B2E_B_COMMAND,            /* [3] vel_command_b */
  B2E_B_HEADING_TARGET, B2E_B_HEADING_ERROR, B2E_B_CMD_TIME_LEFT,   /* [1] */
  B2E_B_CMD_METRICS,        /* [2] error_vel_xy, error_vel_yaw */
  B2E_B_CMD_COUNTER,        /* int [1] */
  B2E_B_IS_HEADING, B2E_B_IS_STANDING,                               /* uint8 [1] */
  B2E_B_AIR_STATE,          /* [4][nfeet]: current_air, last_air, current_contact, last_contact */
  B2E_B_AIR_LAST_TIME,      /* [1] */
  B2E_B_PEAK_HEIGHTS,       /* [nfeet] feet_swing_height state */
  B2E_B_PUSH_TIME_LEFT,     /* [1] */
  B2E_B_EPISODE_LENGTH,     /* int64 [1] */
  B2E_B_REWARD,             /* [1] */
  B2E_B_STEP_REWARD,        /* [B2E_NREW] value / dt */
  B2E_B_EPISODE_SUMS,       /* [B2E_NREW] */
  B2E_B_TERM_DONES,         /* uint8 [B2E_NTERM] */
  B2E_B_TERMINATED, B2E_B_TIME_OUTS, B2E_B_RESET_BUF,               /* uint8 [1] */
  B2E_B_OBS_POLICY,         /* [n_obs_policy] */
  B2E_B_OBS_CRITIC,         /* [n_obs_critic] */
  B2E_B_OBS_EXTRA,          /* [6] terms outside the shipped groups, for cfg-built groups: base_lin_vel, base_ang_vel (envs/mdp/observations.py:24-35) */
  B2E_B_LOG_ENV,            /* [B2E_NLOG] per-env contributions to the log means */
  /* scalars on the device */
  B2E_B_LOG,                /* [B2E_NLOG] reduced over envs by b2_env_log */
  B2E_B_STEP_COUNTER,       /* int64 [1]: common_step_counter, advanced by b2_env_log (the last launch of an env step) */
  B2E_B_ANY_RESET,          /* int [1]: number of envs reset in this step (gate of the conditional forward) */
  B2E_B_INJECT_U,           /* optional [nworld, B2E_NDRAW] uniforms in [0,1) that override the generator (tests); NULL = off */
  B2E_B_COUNT
};

/* log vector layout (sums over envs; the host turns them into the reference's means) */
enum b2e_log {
  B2E_LOG_NRESET = 0,
  B2E_LOG_EPISODE_SUM,                       /* [B2E_NREW] sum over reset envs of the episode sums */
  B2E_LOG_TERM_COUNT = B2E_LOG_EPISODE_SUM + B2E_NREW,   /* [B2E_NTERM] */
  B2E_LOG_CMD_ERR_XY = B2E_LOG_TERM_COUNT + B2E_NTERM, B2E_LOG_CMD_ERR_YAW,  /* sums over reset envs */
  B2E_LOG_ANGMOM,                            /* sum over all envs of |angmom| */
  B2E_LOG_AIR_TIME_SUM, B2E_LOG_AIR_COUNT,   /* feet_air_time metric (only when its weight != 0) */
  B2E_LOG_PEAK_SUM, B2E_LOG_LANDINGS,        /* feet_swing_height metric */
  B2E_LOG_SLIP_SUM, B2E_LOG_NCONTACT,        /* feet_slip metric */
  B2E_LOG_LANDING_FORCE_SUM,                 /* soft_landing metric (denominator: LANDINGS) */
  B2E_LOG_TERRAIN_LEVEL,                     /* sum over all envs of the terrain level (Curriculum/terrain_levels = mean) */
  B2E_NLOG
};

/* draw slots of one env step (each slot is one uniform in [0,1)) */
enum b2e_draw {
  B2E_DRAW_RESET_POSE = 0,                   /* [6] */
  B2E_DRAW_RESET_VEL = 6,                    /* [6] */
  B2E_DRAW_RESET_CMD = 12,                   /* [7]: time, vx, vy, wz, heading, is_heading, is_standing */
  B2E_DRAW_CMD = 19,                         /* [7] same, for the timed resample in CommandManager.compute */
  B2E_DRAW_PUSH = 26,                        /* [7]: interval, vel[6] */
  B2E_DRAW_RESET_JOINT_POS = 33,             /* [nj] */
  B2E_DRAW_RESET_JOINT_VEL = 33 + 32,        /* [nj] */
  B2E_DRAW_NOISE = 33 + 64,                  /* [n_obs_policy] */
  B2E_DRAW_TERRAIN = 33 + 64 + 160,          /* [3]: restart level of the curriculum; level, type of randomize_terrain */
  B2E_NDRAW = 33 + 64 + 160 + 3
};

typedef struct b2_env {
  int ip[B2E_IP_COUNT];
  float fp[B2E_FP_COUNT];
  unsigned long long seed;
  void* buf[B2E_B_COUNT];
} b2_env;

int b2_sizeof_env(void);

/* Validate an env description against the model (sizes, caps). 0 = ok. */
int b2_env_check(const b2_model* m, const b2_data* d, const b2_env* e);

/* process_action + apply_action + write_data_to_sim: action history shift, raw*scale+offset, minus encoder
 * bias -> joint position target -> ctrl (builtin position actuators). */
int b2_env_action(const b2_model* m, b2_data* d, const b2_env* e, void* stream);

/* scene.update(dt) after each physics step: contact-sensor air-time state machine at the physics rate. */
int b2_env_airtime(const b2_model* m, b2_data* d, const b2_env* e, void* stream);

/* decimation x [sim.step(); scene.update(dt)] (manager_based_rl_env.py:276-283): nsub physics steps, each followed by the
 * air-time update, chunked over side streams like b2_substeps.  The first sub-step reuses the preceding forward pass's
 * position / velocity stage per world where the state is untouched (b2_substeps_after_forward, b200sim.h). */
int b2_env_substeps(const b2_model* m, b2_data* d, const b2_env* e, int nsub, void* stream);

/* episode counters, terminations, rewards, then the reset of every done env (sim reset, reset events, manager
 * resets).  Writes reset_buf (the mask the following conditional forward consumes) and any_reset. */
int b2_env_post(const b2_model* m, b2_data* d, const b2_env* e, void* stream);

/* command update (metrics, timed resample, heading control), interval events (push), observations.
 * with_update == 0 computes the observations only (ManagerBasedRlEnv.reset, manager_based_rl_env.py:254-270). */
int b2_env_obs(const b2_model* m, b2_data* d, const b2_env* e, int with_update, void* stream);

/* ManagerBasedRlEnv._reset_idx(env_ids) on its own: envs with mask[w] != 0 (NULL = all). Writes reset_buf = mask
 * and any_reset like b2_env_post. */
int b2_env_reset(const b2_model* m, b2_data* d, const b2_env* e, const unsigned char* mask, void* stream);

/* reduce log_env over envs into log (fixed-order tree: deterministic). */
int b2_env_log(const b2_data* d, const b2_env* e, void* stream);

/* mjwarp.forward restricted: worlds with mask[w] == 0 are skipped when mask != NULL; the whole launch is a no-op
 * when gate != NULL and *gate == 0 (device-side test: no host synchronisation).  mask == gate == NULL is b2_forward. */
int b2_forward_masked(const b2_model* m, b2_data* d, const unsigned char* mask, const int* gate, void* stream);

/* ---- actuator ops --------------------------------- */

/* IdealPdActuator.compute followed by DcMotorActuator._clip_effort (pass saturation == NULL for plain PD):
 *   tau = kp*(pos_target - pos) + kd*(vel_target - vel) + effort_target, clamped to +-effort_limit;
 *   DC motor: v = clamp(vel, +-vel_limit*(1 + limit/sat)); max = min(sat*(1 - v/vel_limit), limit),
 *   min = max(sat*(-1 - v/vel_limit), -limit); tau = clamp(tau, min, max).
 * All arrays [n] = nworld*nj, row-major; gains are per env (DR can make them so). */
int b2_actuator_pd(float* effort, const float* pos_target, const float* vel_target, const float* effort_target,
                   const float* pos, const float* vel, const float* kp, const float* kd, const float* effort_limit,
                   const float* saturation_effort, const float* velocity_limit, long long n, void* stream);

/* DelayBuffer.append + compute for one signal: ring[nworld, cap, width], count/head per env; out = the sample
 * lag[w] pushes ago (clamped to what has been pushed since the last reset). */
int b2_delay_step(float* out, float* ring, int* count, int* head, const int* lag, const float* in,
                  int nworld, int cap, int width, void* stream);

/* LearnedMlpActuator: one MLP evaluation per (env, joint) row of `input` [n, layer_in[0]] (the host stacks the
 * scaled position-error / velocity histories in the configured order).  weight/bias are HOST arrays of nlayer DEVICE
 * pointers, weights row-major [out, in]; hidden activation 0 elu, 1 tanh, 2 relu, 3 softsign; the scalar output is
 * multiplied by out_scale and clipped by the DC-motor limits (saturation_effort != NULL) or +-effort_limit. */
int b2_actuator_mlp(float* effort, const float* input, const float* const* weight, const float* const* bias,
                    const int* layer_in, const int* layer_out, int nlayer, int activation, float out_scale,
                    const float* effort_limit, const float* saturation_effort, const float* velocity_limit,
                    const float* vel, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ENV_H_ */
