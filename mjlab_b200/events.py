"""Event terms that run outside the fused step: model-field domain randomisation.

`randomize_field` follows mjlab's unified randomiser (src/mjlab/envs/mdp/events.py:292-520): "abs" overwrites the target
axes and keeps the others, "scale" / "add" always start from the compiled defaults (`sim.get_default_field`) so repeated
calls do not accumulate (tests/test_domain_randomization.py:182-240).  The field must have been made per-world with
`sim.expand_model_fields` first, as `ManagerBasedRlEnv.__init__` does for every DR field (manager_based_rl_env.py:118-129).
Startup / reset-time work on a handful of geoms: plain torch indexing on the engine's model views, not a hot path.
"""

from __future__ import annotations

import math

import torch

# field -> (entity kind, default axes) (events.py:264-289)
FIELD_SPECS = {
  "dof_armature": ("dof", None), "dof_frictionloss": ("dof", None), "dof_damping": ("dof", None),
  "jnt_range": ("joint", None), "jnt_stiffness": ("joint", None),
  "body_mass": ("body", None), "body_ipos": ("body", [0, 1, 2]), "body_iquat": ("body", [0, 1, 2, 3]), "body_inertia": ("body", None),
  "body_pos": ("body", [0, 1, 2]), "body_quat": ("body", [0, 1, 2, 3]),
  "geom_friction": ("geom", [0]), "geom_pos": ("geom", [0, 1, 2]), "geom_quat": ("geom", [0, 1, 2, 3]),
  "site_pos": ("site", [0, 1, 2]), "site_quat": ("site", [0, 1, 2, 3]), "qpos0": ("qpos", None),
}


def _sample(distribution: str, lo: float, hi: float, shape, device) -> torch.Tensor:
  """One draw of `shape` values (envs/mdp/events.py:497-512 -> utils/lab_api/math.py:1360-1433).  Tests replace this hook to
  replay the draws recorded from the reference's own randomize_field (tools/make_event_golden.py)."""
  if distribution == "uniform":
    return torch.rand(shape, device=device) * (hi - lo) + lo
  if distribution == "log_uniform":
    return torch.exp(torch.rand(shape, device=device) * (math.log(hi) - math.log(lo)) + math.log(lo))
  if distribution == "gaussian":
    # The reference passes one-element tensors as mean / std, so torch.normal returns ONE sample that is broadcast over all
    # selected envs and entities (math.py:1428-1433 with events.py:441-452).  Mirrored, not "fixed": drop-in behaviour.
    return torch.normal(mean=torch.tensor([float(lo)], device=device), std=torch.tensor([float(hi)], device=device))
  raise ValueError(f"Unknown distribution: {distribution}")


def _entity_ids(ix, kind: str, asset_cfg) -> torch.Tensor:
  """events.py:356-378: the entity's global indices of the selected local ids (`slice(None)` = all)."""
  if kind == "geom":
    return ix.geom_ids[asset_cfg.geom_ids]
  if kind == "body":
    return ix.body_ids[asset_cfg.body_ids]
  if kind == "site":
    return ix.site_ids[asset_cfg.site_ids]
  return {"dof": ix.joint_v_adr, "joint": ix.joint_ids, "qpos": ix.joint_q_adr}[kind][asset_cfg.joint_ids]


class _Sel:
  """Stand-in for a resolved SceneEntityCfg built from the *_names convenience arguments."""

  def __init__(self, robot, geom_names, body_names, joint_names, site_names):
    self.name = "robot"
    self.geom_ids = robot.find_geoms(geom_names)[0] if geom_names else slice(None)
    self.body_ids = robot.find_bodies(body_names)[0] if body_names else slice(None)
    self.site_ids = robot.find_sites(site_names)[0] if site_names else slice(None)
    self.joint_ids = robot.find_joints(joint_names)[0] if joint_names else slice(None)


def randomize_field(env, env_ids, field: str, ranges, distribution: str = "uniform", operation: str = "abs", asset_cfg=None,
                    axes=None, geom_names=None, body_names=None, joint_names=None, site_names=None) -> None:
  """Same positional signature as the reference (events.py:292-301); `asset_cfg` is a resolved SceneEntityCfg (`name`,
  `geom_ids` / `body_ids` / `joint_ids` / `site_ids`: local ids, list or slice).  The *_names keywords are a shorthand that
  resolves names on the entity first."""
  if field not in FIELD_SPECS:
    raise ValueError(f"Unknown field '{field}'. Supported fields: {list(FIELD_SPECS)}")
  kind, default_axes = FIELD_SPECS[field]
  from . import engine as _E
  if field in _E.UNCONSUMED_PHYSICS_FIELDS:
    raise ValueError(f"'{field}' is not consumed by the engine (unsupported feature): randomising it would silently do nothing")
  dev = env.device
  robot = env.scene[asset_cfg.name if asset_cfg is not None else "robot"]
  if asset_cfg is None:
    asset_cfg = _Sel(robot, geom_names, body_names, joint_names, site_names)
  ids = torch.as_tensor(_entity_ids(robot.indexing, kind, asset_cfg), device=dev).long()
  model_field = getattr(env.sim.model, field)
  if model_field.shape[0] != env.num_envs or model_field.stride(0) == 0:
    raise ValueError(f"'{field}' is shared by all worlds; call sim.expand_model_fields(('{field}',)) first")
  env_ids = torch.arange(env.num_envs, device=dev) if env_ids is None else env_ids.to(dev)
  eg, ig = torch.meshgrid(env_ids.long(), ids, indexing="ij")
  cur = model_field[eg, ig]
  if axes is not None:
    target = list(axes)
  elif isinstance(ranges, dict):
    target = list(ranges)
  elif default_axes is not None:
    target = default_axes
  else:
    target = list(range(model_field.shape[-1])) if model_field.dim() > 2 else [0]
  if isinstance(ranges, tuple):
    rng = {a: ranges for a in target}
  elif isinstance(ranges, dict):
    missing = set(target) - set(ranges)
    if missing:
      raise ValueError(f"Missing ranges for axes {missing} in field '{field}'. Required axes: {target}")
    rng = {a: ranges[a] for a in target}
  else:
    raise TypeError(f"ranges must be tuple or dict, got {type(ranges)}")
  if operation in ("scale", "add"):
    base = env.sim.get_default_field(field)[ids].unsqueeze(0).expand_as(cur)
  elif operation == "abs":
    base = cur
  else:
    raise ValueError(f"Unknown operation: {operation}")
  rnd = torch.ones_like(cur) if operation == "scale" else torch.zeros_like(cur) if operation == "add" else cur.clone()
  for a in target:
    lo, hi = rng[a]
    if cur.dim() > 2:
      rnd[..., a] = _sample(distribution, lo, hi, (*cur.shape[:-1], 1), dev).squeeze(-1)
    else:
      rnd = _sample(distribution, lo, hi, cur.shape, dev).expand_as(cur).to(cur.dtype)
  if operation == "add":
    model_field[eg, ig] = base + rnd
  elif operation == "scale":
    model_field[eg, ig] = base * rnd
  else:
    model_field[eg, ig] = rnd


# ---- actuator / wrench / encoder events (envs/mdp/events.py:226-252, 519-790) -------------------------------------------------
def _env_ids(env, env_ids) -> torch.Tensor:
  if env_ids is None:
    return torch.arange(env.num_envs, device=env.device, dtype=torch.long)
  return torch.as_tensor(env_ids, device=env.device).long()


def _selected_actuators(env, asset_cfg):
  """events.py:556-567: the entity's actuator objects selected by `asset_cfg.actuator_ids`, DelayedActuators unwrapped."""
  asset = env.scene[asset_cfg.name if asset_cfg is not None else "robot"]
  ids = getattr(asset_cfg, "actuator_ids", slice(None)) if asset_cfg is not None else slice(None)
  if isinstance(ids, list):
    acts = [asset.actuators[i] for i in ids]
  elif isinstance(ids, slice):
    acts = list(asset.actuators[ids])
  else:
    acts = [asset.actuators[ids]]
  return [getattr(a, "base_actuator", a) for a in acts]


def _per_world(env, *fields: str) -> None:
  for f in fields:
    t = getattr(env.sim.model, f)
    if t.shape[0] != env.num_envs or t.stride(0) == 0:
      raise ValueError(f"'{f}' is shared by all worlds; call sim.expand_model_fields(('{f}',)) first")


def randomize_pd_gains(env, env_ids, kp_range, kd_range, asset_cfg=None, distribution: str = "uniform", operation: str = "scale") -> None:
  """Stiffness / damping randomisation (events.py:519-624).  Builtin position actuators live in the model: gainprm[0] = kp,
  biasprm[1] = -kp, biasprm[2] = -kd per world (the physics kernels read them every step); objects with `set_gains`
  (IdealPdActuator and its subclasses) keep per-env gain tensors.  "scale" always starts from the compiled / default gains, so
  repeated calls do not accumulate.  Draw order as the reference: per actuator, kp then kd, one value per (env, ctrl)."""
  if operation not in ("scale", "abs"):
    raise ValueError(f"Unknown operation: {operation}")
  ids = _env_ids(env, env_ids)
  dev = env.device
  for act in _selected_actuators(env, asset_cfg):
    ctrl = act.ctrl_ids
    kp = _sample(distribution, kp_range[0], kp_range[1], (len(ids), len(ctrl)), dev)
    kd = _sample(distribution, kd_range[0], kd_range[1], (len(ids), len(ctrl)), dev)
    if hasattr(act, "set_gains"):
      if operation == "scale":
        act.set_gains(ids, kp=act.default_stiffness[ids] * kp, kd=act.default_damping[ids] * kd)
      else:
        act.set_gains(ids, kp=kp, kd=kd)
      continue
    _per_world(env, "actuator_gainprm", "actuator_biasprm")
    gain, bias = env.sim.model.actuator_gainprm, env.sim.model.actuator_biasprm
    if operation == "scale":
      g0, b0 = env.sim.get_default_field("actuator_gainprm"), env.sim.get_default_field("actuator_biasprm")
      gain[ids[:, None], ctrl, 0] = g0[ctrl, 0] * kp
      bias[ids[:, None], ctrl, 1] = b0[ctrl, 1] * kp
      bias[ids[:, None], ctrl, 2] = b0[ctrl, 2] * kd
    else:
      gain[ids[:, None], ctrl, 0] = kp
      bias[ids[:, None], ctrl, 1] = -kp
      bias[ids[:, None], ctrl, 2] = -kd


def randomize_effort_limits(env, env_ids, effort_limit_range, asset_cfg=None, distribution: str = "uniform", operation: str = "scale") -> None:
  """Effort-limit randomisation (events.py:627-706): `actuator_forcerange` = -/+ limit for builtin actuators, `force_limit` for
  the PD classes.  As in the reference, "scale" multiplies the CURRENT limits (it does accumulate over repeated calls)."""
  if operation not in ("scale", "abs"):
    raise ValueError(f"Unknown operation: {operation}")
  ids = _env_ids(env, env_ids)
  for act in _selected_actuators(env, asset_cfg):
    ctrl = act.ctrl_ids
    eff = _sample(distribution, effort_limit_range[0], effort_limit_range[1], (len(ids), len(ctrl)), env.device)
    if hasattr(act, "set_effort_limit"):
      act.set_effort_limit(ids, effort_limit=act.force_limit[ids].clone() * eff if operation == "scale" else eff)
      continue
    _per_world(env, "actuator_forcerange")
    fr = env.sim.model.actuator_forcerange
    if operation == "scale":
      fr[ids[:, None], ctrl, 0] *= eff
      fr[ids[:, None], ctrl, 1] *= eff
    else:
      fr[ids[:, None], ctrl, 0] = -eff
      fr[ids[:, None], ctrl, 1] = eff


def randomize_encoder_bias(env, env_ids, bias_range, asset_cfg=None) -> None:
  """Joint-encoder calibration error (events.py:709-744): U(bias_range) per (env, joint) into `entity.data.encoder_bias`, which
  the observation kernels add to `joint_pos_rel(biased=True)`."""
  asset = env.scene[asset_cfg.name if asset_cfg is not None else "robot"]
  ids = _env_ids(env, env_ids)
  joint_ids = getattr(asset_cfg, "joint_ids", slice(None)) if asset_cfg is not None else slice(None)
  n = asset.num_joints if isinstance(joint_ids, slice) else len(joint_ids)
  bias = _sample("uniform", bias_range[0], bias_range[1], (len(ids), n), env.device)
  if isinstance(joint_ids, slice):
    asset.data.encoder_bias[ids] = bias
  else:
    asset.data.encoder_bias[ids[:, None], torch.as_tensor(joint_ids, device=env.device)] = bias


def apply_external_force_torque(env, env_ids, force_range, torque_range, asset_cfg=None) -> None:
  """U(force_range) / U(torque_range) per (env, body, axis) into `xfrc_applied` of the selected bodies (events.py:226-245);
  the wrench stays until overwritten (the position / velocity-stage reuse of the env step sees a nonzero wrench and recomputes)."""
  asset = env.scene[asset_cfg.name if asset_cfg is not None else "robot"]
  ids = _env_ids(env, env_ids)
  body_ids = getattr(asset_cfg, "body_ids", slice(None)) if asset_cfg is not None else slice(None)
  nb = len(body_ids) if isinstance(body_ids, list) else asset.num_bodies
  size = (len(ids), nb, 3)
  forces = _sample("uniform", force_range[0], force_range[1], size, env.device)
  torques = _sample("uniform", torque_range[0], torque_range[1], size, env.device)
  asset.write_external_wrench_to_sim(forces, torques, env_ids=ids, body_ids=body_ids)


def sync_actuator_delays(env, env_ids, lag_range, asset_cfg=None) -> None:
  """One lag per env, applied to every delayed actuator of the entity (events.py:747-790)."""
  asset = env.scene[asset_cfg.name if asset_cfg is not None else "robot"]
  ids = _env_ids(env, env_ids)
  sel = getattr(asset_cfg, "actuator_ids", slice(None)) if asset_cfg is not None else slice(None)
  acts = [asset.actuators[i] for i in sel] if isinstance(sel, list) else list(asset.actuators[sel]) if isinstance(sel, slice) else [asset.actuators[sel]]
  delayed = [a for a in acts if hasattr(a, "set_lags")]
  if not delayed:
    return
  lags = torch.randint(int(lag_range[0]), int(lag_range[1]) + 1, (len(ids),), device=env.device)
  for a in delayed:
    a.set_lags(lags, ids)


# ---- a consumed ManagerBasedRlEnvCfg's startup events (envcfg.from_manager_cfg -> extra["startup_dr"]) -------------------------
STARTUP_ACTUATOR_EVENTS = {"randomize_pd_gains": ("actuator_gainprm", "actuator_biasprm"), "randomize_effort_limits": ("actuator_forcerange",),
                           "randomize_encoder_bias": ()}


def apply_startup_event(env, ev: dict) -> None:
  """One startup event of the cfg, applied the way `ManagerBasedRlEnv.__init__` does (manager_based_rl_env.py:118-129, 219):
  the model fields it writes get one row per world first, then the term function runs once over all envs."""
  from .managers import SceneEntityCfg
  ac = SceneEntityCfg(**{k: v for k, v in ev["asset_cfg"].items() if v is not None})
  ac.resolve(env.scene)
  kind = ev.get("kind", "randomize_field")
  if kind == "randomize_field":
    if ev["field"] not in FIELD_SPECS:
      from .envcfg import UnsupportedTermError
      raise UnsupportedTermError(f"startup event '{ev['name']}': randomize_field does not know the field '{ev['field']}'")
    env.sim.expand_model_fields((ev["field"],))
    randomize_field(env, None, ev["field"], ev["ranges"], ev["distribution"], ev["operation"], asset_cfg=ac, axes=ev["axes"])
    return
  env.sim.expand_model_fields(STARTUP_ACTUATOR_EVENTS[kind])
  {"randomize_pd_gains": randomize_pd_gains, "randomize_effort_limits": randomize_effort_limits,
   "randomize_encoder_bias": randomize_encoder_bias}[kind](env, None, asset_cfg=ac, **ev["params"])
