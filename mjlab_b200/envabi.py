"""ctypes binding of include/b200env.h (the fused env-step + actuator C ABI).

Enum values are read from the header itself, so the Python side cannot drift from the C side; the struct size is
checked against `b2_sizeof_env()` when the library loads.
"""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from . import engine as E

HEADER = Path(__file__).resolve().parents[1] / "include" / "b200env.h"
_FALLBACK_HEADER = Path(__file__).resolve().parent / "include" / "b200env.h"


def _parse_header() -> dict[str, int]:
  path = HEADER if HEADER.exists() else _FALLBACK_HEADER
  text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
  names: dict[str, int] = {}
  for mac in re.finditer(r"#define\s+(B2E_\w+)\s+(\d+)", text):
    names[mac.group(1)] = int(mac.group(2))
  for body in re.findall(r"enum\s+\w+\s*\{(.*?)\}", text, flags=re.S):
    nxt = 0
    for item in body.split(","):
      item = item.strip()
      if not item:
        continue
      if "=" in item:
        name, expr = (s.strip() for s in item.split("=", 1))
        nxt = int(eval(expr, {"__builtins__": {}}, names))  # sums of earlier names / integers only
      else:
        name = item
      names[name] = nxt
      nxt += 1
  return names


K = _parse_header()
NIP, NFP, NBUF, NLOG, NDRAW, NREW, NTERM, MAXFEET = (K["B2E_IP_COUNT"], K["B2E_FP_COUNT"], K["B2E_B_COUNT"], K["B2E_NLOG"],
                                                      K["B2E_NDRAW"], K["B2E_NREW"], K["B2E_NTERM"], K["B2E_MAXFEET"])


class B2Env(C.Structure):
  _fields_ = [("ip", C.c_int * NIP), ("fp", C.c_float * NFP), ("seed", C.c_ulonglong), ("buf", C.c_void_p * NBUF)]


_bound = False


def load_library() -> C.CDLL:
  """libb200sim.so with the env-layer prototypes attached (raises if the CUDA library is missing)."""
  global _bound
  lib = E.load_library()
  if _bound:
    return lib
  pm, pd, pe, vp = C.POINTER(E.B2Model), C.POINTER(E.B2Data), C.POINTER(B2Env), C.c_void_p
  lib.b2_sizeof_env.restype = C.c_int
  if lib.b2_sizeof_env() != C.sizeof(B2Env):
    raise RuntimeError("libb200sim.so ABI mismatch with include/b200env.h (rebuild the extension)")
  lib.b2_env_check.argtypes = [pm, pd, pe]
  for name in ("b2_env_action", "b2_env_airtime", "b2_env_post"):
    getattr(lib, name).argtypes = [pm, pd, pe, vp]
  lib.b2_env_obs.argtypes = [pm, pd, pe, C.c_int, vp]
  lib.b2_env_substeps.argtypes = [pm, pd, pe, C.c_int, vp]
  lib.b2_env_substeps.restype = C.c_int
  lib.b2_env_reset.argtypes = [pm, pd, pe, vp, vp]
  lib.b2_env_log.argtypes = [pd, pe, vp]
  lib.b2_forward_masked.argtypes = [pm, pd, vp, vp, vp]
  lib.b2_actuator_pd.argtypes = [vp] * 11 + [C.c_longlong, vp]
  lib.b2_delay_step.argtypes = [vp] * 6 + [C.c_int, C.c_int, C.c_int, vp]
  lib.b2_actuator_mlp.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                  C.c_float, vp, vp, vp, vp, C.c_longlong, vp]
  for name in ("b2_env_check", "b2_env_action", "b2_env_airtime", "b2_env_post", "b2_env_obs", "b2_env_log", "b2_env_reset", "b2_forward_masked",
               "b2_actuator_pd", "b2_delay_step", "b2_actuator_mlp"):
    getattr(lib, name).restype = C.c_int
  _bound = True
  return lib


# ---------------------------------------------------------------------------------------------------------
# b2_env construction shared by the CUDA env (torch tensors) and the host-emulation test harness (numpy)
# ---------------------------------------------------------------------------------------------------------
# (enum suffix, attribute, kind, dtype, per-item shape as a function of the resolved cfg)
#   kind "table": one copy shared by all envs, filled from the cfg;  "state": [nworld, *shape], zero-initialised;
#   "scalar": a single device array of `shape`.
BUFFERS = [
  ("JOINT_Q_ADR", "joint_q_adr", "table", "i4", lambda c: (c.nj,)),
  ("JOINT_V_ADR", "joint_v_adr", "table", "i4", lambda c: (c.nj,)),
  ("CTRL_IDS", "ctrl_ids", "table", "i4", lambda c: (c.nj,)),
  ("DEFAULT_JOINT_POS", "default_joint_pos", "table", "f4", lambda c: (c.nj,)),
  ("DEFAULT_JOINT_VEL", "default_joint_vel", "table", "f4", lambda c: (c.nj,)),
  ("SOFT_LO", "soft_lo", "table", "f4", lambda c: (c.nj,)),
  ("SOFT_HI", "soft_hi", "table", "f4", lambda c: (c.nj,)),
  ("ACTION_SCALE", "action_scale", "table", "f4", lambda c: (c.nj,)),
  ("ACTION_OFFSET", "action_offset", "table", "f4", lambda c: (c.nj,)),
  ("STD_STANDING", "std_standing", "table", "f4", lambda c: (c.nj,)),
  ("STD_WALKING", "std_walking", "table", "f4", lambda c: (c.nj,)),
  ("STD_RUNNING", "std_running", "table", "f4", lambda c: (c.nj,)),
  ("NOISE_LO", "noise_lo", "table", "f4", lambda c: (c.n_obs_policy,)),
  ("NOISE_HI", "noise_hi", "table", "f4", lambda c: (c.n_obs_policy,)),
  ("DEFAULT_ROOT_STATE", "default_root_state", "table", "f4", lambda c: (13,)),
  ("TERRAIN_ORIGINS", "terrain_origins", "table", "f4", lambda c: (max(c.terrain_rows, 1), max(c.terrain_cols, 1), 3)),
  ("ENV_ORIGINS", "env_origins", "state", "f4", lambda c: (3,)),
  ("TERRAIN_LEVEL", "terrain_level", "state", "i4", lambda c: ()),
  ("TERRAIN_TYPE", "terrain_type", "state", "i4", lambda c: ()),
  ("ACTION_IN", "action_in", "state", "f4", lambda c: (c.nj,)),
  ("ACTION", "action", "state", "f4", lambda c: (c.nj,)),
  ("PREV_ACTION", "prev_action", "state", "f4", lambda c: (c.nj,)),
  ("PREV_PREV_ACTION", "prev_prev_action", "state", "f4", lambda c: (c.nj,)),
  ("RAW_ACTION", "raw_action", "state", "f4", lambda c: (c.nj,)),
  ("PROCESSED_ACTION", "processed_action", "state", "f4", lambda c: (c.nj,)),
  ("ENCODER_BIAS", "encoder_bias", "state", "f4", lambda c: (c.nj,)),
  ("JOINT_POS_TARGET", "joint_pos_target", "state", "f4", lambda c: (c.nj,)),
  ("COMMAND", "command", "state", "f4", lambda c: (3,)),
  ("HEADING_TARGET", "heading_target", "state", "f4", lambda c: ()),
  ("HEADING_ERROR", "heading_error", "state", "f4", lambda c: ()),
  ("CMD_TIME_LEFT", "cmd_time_left", "state", "f4", lambda c: ()),
  ("CMD_METRICS", "cmd_metrics", "state", "f4", lambda c: (2,)),
  ("CMD_COUNTER", "cmd_counter", "state", "i4", lambda c: ()),
  ("IS_HEADING", "is_heading", "state", "u1", lambda c: ()),
  ("IS_STANDING", "is_standing", "state", "u1", lambda c: ()),
  ("AIR_STATE", "air_state", "state", "f4", lambda c: (4, c.nfeet)),
  ("AIR_LAST_TIME", "air_last_time", "state", "f4", lambda c: ()),
  ("PEAK_HEIGHTS", "peak_heights", "state", "f4", lambda c: (c.nfeet,)),
  ("PUSH_TIME_LEFT", "push_time_left", "state", "f4", lambda c: ()),
  ("EPISODE_LENGTH", "episode_length", "state", "i8", lambda c: ()),
  ("REWARD", "reward", "state", "f4", lambda c: ()),
  ("STEP_REWARD", "step_reward", "state", "f4", lambda c: (NREW,)),
  ("EPISODE_SUMS", "episode_sums", "state", "f4", lambda c: (NREW,)),
  ("TERM_DONES", "term_dones", "state", "u1", lambda c: (NTERM,)),
  ("TERMINATED", "terminated", "state", "u1", lambda c: ()),
  ("TIME_OUTS", "time_outs", "state", "u1", lambda c: ()),
  ("RESET_BUF", "reset_buf", "state", "u1", lambda c: ()),
  ("OBS_POLICY", "obs_policy", "state", "f4", lambda c: (c.n_obs_policy,)),
  ("OBS_CRITIC", "obs_critic", "state", "f4", lambda c: (c.n_obs_critic,)),
  ("OBS_EXTRA", "obs_extra", "state", "f4", lambda c: (6,)),
  ("LOG_ENV", "log_env", "state", "f4", lambda c: (NLOG,)),
  ("LOG", "log", "scalar", "f4", lambda c: (NLOG,)),
  ("STEP_COUNTER", "step_counter", "scalar", "i8", lambda c: (1,)),
  ("ANY_RESET", "any_reset", "scalar", "i4", lambda c: (1,)),
]
assert [K["B2E_B_" + b[0]] for b in BUFFERS] == list(range(len(BUFFERS))) and K["B2E_B_INJECT_U"] == len(BUFFERS)


def table_values(cfg) -> dict:
  """Per-joint / per-observation tables of the resolved velocity cfg, keyed by BUFFERS attribute."""
  import numpy as np
  return dict(
    joint_q_adr=cfg.joint_q_adr, joint_v_adr=cfg.joint_v_adr, ctrl_ids=cfg.ctrl_ids, default_joint_pos=cfg.default_joint_pos,
    default_joint_vel=np.zeros(cfg.nj), soft_lo=cfg.soft_joint_pos_limits[:, 0], soft_hi=cfg.soft_joint_pos_limits[:, 1],
    action_scale=cfg.action_scale, action_offset=cfg.default_joint_pos, std_standing=cfg.std_standing, std_walking=cfg.std_walking,
    std_running=cfg.std_running, noise_lo=cfg.noise_lo, noise_hi=cfg.noise_hi, default_root_state=cfg.default_root_state,
    terrain_origins=cfg.terrain_origins if cfg.terrain_origins is not None else np.zeros((1, 1, 3)),
  )


def fill_params(ce: B2Env, cfg) -> None:
  """Scalar parameters of b2_env from the resolved cfg (mjlab_b200.envcfg.VelocityCfg)."""
  ip, fp = ce.ip, ce.fp
  for name, v in (("NJ", cfg.nj), ("NFEET", cfg.nfeet), ("ROOT_BODY", cfg.root_body), ("UPRIGHT_BODY", cfg.upright_body),
                  ("N_OBS_POLICY", cfg.n_obs_policy), ("N_OBS_CRITIC", cfg.n_obs_critic), ("LINVEL_ADR", cfg.linvel_adr),
                  ("ANGVEL_ADR", cfg.angvel_adr), ("ANGMOM_ADR", cfg.angmom_adr), ("SELFCOL_ADR", cfg.selfcol_adr),
                  ("ILLEGAL_ADR", max(cfg.illegal_adr, 0)), ("ILLEGAL_COUNT", cfg.illegal_count),
                  ("MAX_EPISODE_LENGTH", min(int(cfg.max_episode_length), 2**31 - 1)), ("ENABLE_CORRUPTION", int(cfg.enable_corruption)),
                  ("PUSH_ENABLED", int(cfg.push_enabled)), ("HEADING_COMMAND", int(cfg.heading_command)),
                  ("TERRAIN_MODE", int(cfg.terrain_mode)), ("TERRAIN_ROWS", int(cfg.terrain_rows)), ("TERRAIN_COLS", int(cfg.terrain_cols)),
                  ("TERM_FLAGS", (1 if cfg.min_root_height is not None else 0) | (2 if cfg.nan_termination else 0))):
    ip[K["B2E_IP_" + name]] = int(v)
  for f in range(cfg.nfeet):
    ip[K["B2E_IP_FOOT_SITE"] + f] = int(cfg.foot_site_ids[f])
    ip[K["B2E_IP_FOOT_SITE_BODY"] + f] = int(cfg.foot_site_body[f])
    ip[K["B2E_IP_FOOT_FOUND_ADR"] + f] = int(cfg.foot_found_adr[f])
    ip[K["B2E_IP_FOOT_FORCE_ADR"] + f] = int(cfg.foot_force_adr[f])
  for name, v in (("STEP_DT", cfg.step_dt), ("MAX_EPISODE_LENGTH_S", cfg.max_episode_length_s), ("LIMIT_ANGLE", cfg.limit_angle),
                  ("STD_LIN", cfg.std_lin), ("STD_ANG", cfg.std_ang), ("STD_UPRIGHT", cfg.std_upright),
                  ("WALKING_THRESHOLD", cfg.walking_threshold), ("RUNNING_THRESHOLD", cfg.running_threshold),
                  ("AIR_THR_MIN", cfg.air_thr_min), ("AIR_THR_MAX", cfg.air_thr_max), ("AIR_CMD_THRESHOLD", cfg.air_cmd_threshold),
                  ("TARGET_HEIGHT", cfg.target_height), ("CMD_THRESHOLD", cfg.cmd_threshold), ("REL_STANDING", cfg.rel_standing),
                  ("TERRAIN_HALF_SIZE", cfg.terrain_half_size), ("MIN_ROOT_HEIGHT", cfg.min_root_height if cfg.min_root_height is not None else 0.0),
                  ("REL_HEADING", cfg.rel_heading), ("HEADING_STIFFNESS", cfg.heading_stiffness)):
    fp[K["B2E_FP_" + name]] = float(v)
  for name, pair in (("CMD_RESAMPLE", cfg.cmd_resample), ("CMD_LIN_X", cfg.cmd_lin_x), ("CMD_LIN_Y", cfg.cmd_lin_y),
                     ("CMD_ANG_Z", cfg.cmd_ang_z), ("CMD_HEADING", cfg.cmd_heading), ("RESET_JOINT_POS", cfg.reset_joint_pos),
                     ("RESET_JOINT_VEL", cfg.reset_joint_vel), ("PUSH_INTERVAL", cfg.push_interval)):
    fp[K["B2E_FP_" + name]], fp[K["B2E_FP_" + name] + 1] = float(pair[0]), float(pair[1])
  for name, arr in (("RESET_POSE", cfg.reset_pose), ("RESET_VEL", cfg.reset_vel), ("PUSH_VEL", cfg.push_vel)):
    for k in range(6):
      fp[K["B2E_FP_" + name] + 2 * k], fp[K["B2E_FP_" + name] + 2 * k + 1] = float(arr[k][0]), float(arr[k][1])
  for k in range(NREW):
    fp[K["B2E_FP_REWARD_WEIGHT"] + k] = float(cfg.reward_weights[k])
  ce.seed = int(cfg.seed) & (2**64 - 1)
