"""ctypes binding of include/b200trk.h (fused env step of the motion-tracking task); enum values come from the header."""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from . import engine as E

HEADER = Path(__file__).resolve().parents[1] / "include" / "b200trk.h"


def _parse_header() -> dict[str, int]:
  text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
  names: dict[str, int] = {}
  for mac in re.finditer(r"#define\s+(B2T_\w+)\s+(\d+)", text):
    names[mac.group(1)] = int(mac.group(2))
  for body in re.findall(r"enum\s+\w+\s*\{(.*?)\}", text, flags=re.S):
    nxt = 0
    for item in body.split(","):
      item = item.strip()
      if not item:
        continue
      if "=" in item:
        name, expr = (s.strip() for s in item.split("=", 1))
        nxt = int(eval(expr, {"__builtins__": {}}, names))
      else:
        name = item
      names[name] = nxt
      nxt += 1
  return names


K = _parse_header()
NIP, NFP, NBUF, NLOG, NDRAW, NREW, NTERM, NMETRIC = (K["B2T_IP_COUNT"], K["B2T_FP_COUNT"], K["B2T_B_COUNT"], K["B2T_NLOG"], K["B2T_NDRAW"],
                                                      K["B2T_NREW"], K["B2T_NTERM"], K["B2T_NMETRIC"])


class B2Trk(C.Structure):
  _fields_ = [("ip", C.c_int * NIP), ("fp", C.c_float * NFP), ("seed", C.c_ulonglong), ("buf", C.c_void_p * NBUF)]


_bound = False


def load_library() -> C.CDLL:
  global _bound
  from . import envabi
  lib = envabi.load_library()   # also binds b2_forward_masked and the other shared entry points
  if _bound:
    return lib
  pm, pd, pe, vp = C.POINTER(E.B2Model), C.POINTER(E.B2Data), C.POINTER(B2Trk), C.c_void_p
  lib.b2_sizeof_trk.restype = C.c_int
  if lib.b2_sizeof_trk() != C.sizeof(B2Trk):
    raise RuntimeError("libb200sim.so ABI mismatch with include/b200trk.h (rebuild the extension)")
  lib.b2_trk_check.argtypes = [pm, pd, pe]
  lib.b2_trk_action.argtypes = lib.b2_trk_post.argtypes = [pm, pd, pe, vp]
  lib.b2_trk_reset.argtypes = [pm, pd, pe, vp, vp]
  lib.b2_trk_obs.argtypes = [pm, pd, pe, C.c_int, vp]
  lib.b2_trk_log.argtypes = [pd, pe, vp]
  for name in ("b2_trk_check", "b2_trk_action", "b2_trk_post", "b2_trk_reset", "b2_trk_obs", "b2_trk_log"):
    getattr(lib, name).restype = C.c_int
  _bound = True
  return lib


# (enum suffix, attribute, kind, dtype, shape(cfg)); kinds: table (from cfg), motion (from the clip), state ([nworld, ...]), scalar
BUFFERS = [
  ("JOINT_Q_ADR", "joint_q_adr", "table", "i4", lambda c: (c.nj,)),
  ("JOINT_V_ADR", "joint_v_adr", "table", "i4", lambda c: (c.nj,)),
  ("CTRL_IDS", "ctrl_ids", "table", "i4", lambda c: (c.nj,)),
  ("DEFAULT_JOINT_POS", "default_joint_pos", "table", "f4", lambda c: (c.nj,)),
  ("SOFT_LO", "soft_lo", "table", "f4", lambda c: (c.nj,)),
  ("SOFT_HI", "soft_hi", "table", "f4", lambda c: (c.nj,)),
  ("ACTION_SCALE", "action_scale", "table", "f4", lambda c: (c.nj,)),
  ("ACTION_OFFSET", "action_offset", "table", "f4", lambda c: (c.nj,)),
  ("NOISE_LO", "noise_lo", "table", "f4", lambda c: (c.n_obs_policy,)),
  ("NOISE_HI", "noise_hi", "table", "f4", lambda c: (c.n_obs_policy,)),
  ("BODY_GLOBAL", "body_global", "table", "i4", lambda c: (c.nbody,)),
  ("BODY_MOTION", "body_motion", "table", "i4", lambda c: (c.nbody,)),
  ("M_JOINT_POS", "joint_pos", "motion", "f4", None),
  ("M_JOINT_VEL", "joint_vel", "motion", "f4", None),
  ("M_BODY_POS", "body_pos_w", "motion", "f4", None),
  ("M_BODY_QUAT", "body_quat_w", "motion", "f4", None),
  ("M_BODY_LIN_VEL", "body_lin_vel_w", "motion", "f4", None),
  ("M_BODY_ANG_VEL", "body_ang_vel_w", "motion", "f4", None),
  ("ENV_ORIGINS", "env_origins", "state", "f4", lambda c: (3,)),
  ("ACTION_IN", "action_in", "state", "f4", lambda c: (c.nj,)),
  ("ACTION", "action", "state", "f4", lambda c: (c.nj,)),
  ("PREV_ACTION", "prev_action", "state", "f4", lambda c: (c.nj,)),
  ("PREV_PREV_ACTION", "prev_prev_action", "state", "f4", lambda c: (c.nj,)),
  ("RAW_ACTION", "raw_action", "state", "f4", lambda c: (c.nj,)),
  ("PROCESSED_ACTION", "processed_action", "state", "f4", lambda c: (c.nj,)),
  ("ENCODER_BIAS", "encoder_bias", "state", "f4", lambda c: (c.nj,)),
  ("JOINT_POS_TARGET", "joint_pos_target", "state", "f4", lambda c: (c.nj,)),
  ("TIME_STEPS", "time_steps", "state", "i8", lambda c: ()),
  ("REL_POS", "rel_pos", "state", "f4", lambda c: (c.nbody, 3)),
  ("REL_QUAT", "rel_quat", "state", "f4", lambda c: (c.nbody, 4)),
  ("METRICS", "metrics", "state", "f4", lambda c: (NMETRIC,)),
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
  ("LOG_ENV", "log_env", "state", "f4", lambda c: (NLOG,)),
  ("BIN_FAILED_COUNT", "bin_failed_count", "scalar", "f4", lambda c: (c.bin_count,)),
  ("CURRENT_BIN_FAILED", "current_bin_failed", "scalar", "f4", lambda c: (c.bin_count,)),
  ("LOG", "log", "scalar", "f4", lambda c: (NLOG,)),
  ("STEP_COUNTER", "step_counter", "scalar", "i8", lambda c: (1,)),
  ("ANY_RESET", "any_reset", "scalar", "i4", lambda c: (1,)),
]
assert [K["B2T_B_" + b[0]] for b in BUFFERS] == list(range(len(BUFFERS))) and K["B2T_B_INJECT_U"] == len(BUFFERS)


def table_values(cfg) -> dict:
  return dict(joint_q_adr=cfg.joint_q_adr, joint_v_adr=cfg.joint_v_adr, ctrl_ids=cfg.ctrl_ids, default_joint_pos=cfg.default_joint_pos,
              soft_lo=cfg.soft_joint_pos_limits[:, 0], soft_hi=cfg.soft_joint_pos_limits[:, 1], action_scale=cfg.action_scale,
              action_offset=cfg.default_joint_pos, noise_lo=cfg.noise_lo, noise_hi=cfg.noise_hi, body_global=cfg.body_ids[cfg.body_sel],
              body_motion=cfg.body_sel)


def fill_params(ce: B2Trk, cfg) -> None:
  ip, fp = ce.ip, ce.fp
  mode = {"start": 0, "uniform": 1, "adaptive": 2}[cfg.sampling_mode]
  for name, v in (("NJ", cfg.nj), ("NBODY", cfg.nbody), ("NBODY_MOTION", cfg.nbody_motion), ("ROOT_BODY", cfg.root_body), ("ANCHOR_SEL", cfg.anchor_sel),
                  ("N_OBS_POLICY", cfg.n_obs_policy), ("N_OBS_CRITIC", cfg.n_obs_critic), ("LINVEL_ADR", cfg.linvel_adr), ("ANGVEL_ADR", cfg.angvel_adr),
                  ("SELFCOL_ADR", cfg.selfcol_adr), ("MAX_EPISODE_LENGTH", min(int(cfg.max_episode_length), 2**31 - 1)),
                  ("ENABLE_CORRUPTION", int(cfg.enable_corruption)), ("PUSH_ENABLED", int(cfg.push_enabled)), ("SAMPLING_MODE", mode),
                  ("TIME_STEP_TOTAL", cfg.time_step_total), ("BIN_COUNT", cfg.bin_count), ("KERNEL_SIZE", cfg.adaptive_kernel_size), ("NEE", len(cfg.ee_sel)),
                  ("TERM_FULL3D", int(getattr(cfg, "term_full3d", 0)))):
    ip[K["B2T_IP_" + name]] = int(v)
  for k, s in enumerate(cfg.ee_sel):
    ip[K["B2T_IP_EE_SEL"] + k] = int(s)
  for name, v in (("STEP_DT", cfg.step_dt), ("MAX_EPISODE_LENGTH_S", cfg.max_episode_length_s), ("ANCHOR_POS_THRESHOLD", cfg.anchor_pos_threshold),
                  ("ANCHOR_ORI_THRESHOLD", cfg.anchor_ori_threshold), ("EE_POS_THRESHOLD", cfg.ee_pos_threshold), ("ADAPTIVE_LAMBDA", cfg.adaptive_lambda),
                  ("ADAPTIVE_UNIFORM_RATIO", cfg.adaptive_uniform_ratio), ("ADAPTIVE_ALPHA", cfg.adaptive_alpha)):
    fp[K["B2T_FP_" + name]] = float(v)
  for k, v in enumerate((cfg.std_anchor_pos, cfg.std_anchor_ori, cfg.std_body_pos, cfg.std_body_ori, cfg.std_body_lin_vel, cfg.std_body_ang_vel)):
    fp[K["B2T_FP_STD"] + k] = float(v)
  for name, pair in (("JOINT_POS_RANGE", cfg.joint_position_range), ("PUSH_INTERVAL", cfg.push_interval)):
    fp[K["B2T_FP_" + name]], fp[K["B2T_FP_" + name] + 1] = float(pair[0]), float(pair[1])
  for name, arr in (("POSE_RANGE", cfg.pose_range), ("VELOCITY_RANGE", cfg.velocity_range), ("PUSH_VEL", cfg.push_vel)):
    for k in range(6):
      fp[K["B2T_FP_" + name] + 2 * k], fp[K["B2T_FP_" + name] + 2 * k + 1] = float(arr[k][0]), float(arr[k][1])
  for k in range(NREW):
    fp[K["B2T_FP_REWARD_WEIGHT"] + k] = float(cfg.reward_weights[k])
  ce.seed = int(cfg.seed) & (2**64 - 1)
