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
