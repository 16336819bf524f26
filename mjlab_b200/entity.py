"""Entity / EntityData tensor views over the engine's world-batched buffers.

Host-side mirror of `mjlab.entity.data.EntityData` (src/mjlab/entity/data.py:34-533) and the parts of
`mjlab.entity.entity.Entity` the env step touches (index maps :754-813, `find_*` :215-262, state writes :531-640).
Same property names, shapes, frames and staleness rules: reads come from the engine's derived buffers and are only
current after `sim.forward()`; writes go straight to `qpos` / `qvel` / `ctrl` / `xfrc_applied` / `mocap_*`.

These are generic torch views for user code and tests; the shipped velocity task does not go
through them on its hot path - the fused env kernels (csrc/b2_env.cuh) read the same buffers directly.
"""

from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch


# ---- quaternion helpers, (w, x, y, z) (utils/lab_api/math.py:527-670) ------------------------------------
def quat_apply(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
  xyz = q[..., 1:]
  t = torch.cross(xyz, v, dim=-1) * 2
  return v + q[..., 0:1] * t + torch.cross(xyz, t, dim=-1)


def quat_apply_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
  xyz = q[..., 1:]
  t = torch.cross(xyz, v, dim=-1) * 2
  return v - q[..., 0:1] * t + torch.cross(xyz, t, dim=-1)


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
  w1, x1, y1, z1 = a.unbind(-1)
  w2, x2, y2, z2 = b.unbind(-1)
  return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                      w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)


def quat_from_matrix(mat: torch.Tensor) -> torch.Tensor:
  """Rotation matrices [..., 3, 3] -> unit quaternions with w >= 0 branch selection (math.py:318-372)."""
  m = mat.reshape(mat.shape[:-2] + (9,)) if mat.shape[-1] == 3 else mat
  m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
  q_abs = torch.sqrt(torch.clamp(torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1), min=0))
  cand = torch.stack([
    torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
    torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
    torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
    torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
  cand = cand / (2 * q_abs[..., None].clamp(min=0.1))
  best = q_abs.argmax(-1)
  return torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)


def velocity_from_cvel(pos: torch.Tensor, subtree_com: torch.Tensor, cvel: torch.Tensor) -> torch.Tensor:
  """[lin_w, ang_w] from the engine's com-based spatial velocity (entity/data.py:20-31)."""
  ang = cvel[..., 0:3]
  return torch.cat([cvel[..., 3:6] - torch.cross(ang, subtree_com - pos, dim=-1), ang], -1)


@dataclass
class EntityIndexing:
  """Global ids of one entity inside the compiled scene (entity/entity.py:60-93)."""

  body_ids: torch.Tensor
  geom_ids: torch.Tensor
  site_ids: torch.Tensor
  joint_ids: torch.Tensor
  ctrl_ids: torch.Tensor
  joint_q_adr: torch.Tensor
  joint_v_adr: torch.Tensor
  free_joint_q_adr: torch.Tensor
  free_joint_v_adr: torch.Tensor
  mocap_id: int | None = None

  @property
  def root_body_id(self) -> int:
    return int(self.body_ids[0])


def _resolve(keys, names: list[str], preserve_order: bool = False) -> tuple[list[int], list[str]]:
  """resolve_matching_names (utils/lab_api/string.py:180-271): regex full-match of every key against the name list."""
  keys = [keys] if isinstance(keys, str) else list(keys)
  hits: list[int] = []
  seen = set()
  used = [False] * len(keys)
  order = []
  for i, n in enumerate(names):
    for k, key in enumerate(keys):
      if re.fullmatch(key, n):
        if i in seen:
          raise ValueError(f"Multiple matches for '{n}': {keys}")
        seen.add(i)
        used[k] = True
        order.append((k, i))
  if not all(used):
    raise ValueError(f"Not all regular expressions are matched: {[k for k, u in zip(keys, used) if not u]} in {names}")
  if preserve_order:
    order.sort(key=lambda t: t[0])
  hits = [i for _, i in order]
  return hits, [names[i] for i in hits]


class EntityData:
  """Read properties and write methods of one entity (entity/data.py)."""

  POS_DIM, QUAT_DIM, ROOT_POSE_DIM, ROOT_VEL_DIM, ROOT_STATE_DIM = 3, 4, 7, 6, 13

  def __init__(self, indexing: EntityIndexing, data, model, device: str, *, default_root_state, default_joint_pos, default_joint_vel,
               joint_pos_limits, soft_joint_pos_limits, is_fixed_base=False, is_articulated=True, is_actuated=True,
               joint_pos_target=None, joint_vel_target=None, joint_effort_target=None, encoder_bias=None):
    self.indexing, self.data, self.model, self.device = indexing, data, model, device
    n = data.nworld
    nj = len(indexing.joint_ids)
    z = lambda: torch.zeros(n, nj, device=device)  # noqa: E731
    self.default_root_state, self.default_joint_pos, self.default_joint_vel = default_root_state, default_joint_pos, default_joint_vel
    self.default_joint_pos_limits = joint_pos_limits.clone()
    self.joint_pos_limits, self.soft_joint_pos_limits = joint_pos_limits, soft_joint_pos_limits
    self.gravity_vec_w = torch.tensor([0.0, 0.0, -1.0], device=device).repeat(n, 1)
    self.forward_vec_b = torch.tensor([1.0, 0.0, 0.0], device=device).repeat(n, 1)
    self.is_fixed_base, self.is_articulated, self.is_actuated = is_fixed_base, is_articulated, is_actuated
    self.joint_pos_target = z() if joint_pos_target is None else joint_pos_target
    self.joint_vel_target = z() if joint_vel_target is None else joint_vel_target
    self.joint_effort_target = z() if joint_effort_target is None else joint_effort_target
    self.encoder_bias = z() if encoder_bias is None else encoder_bias

  # ---- writes (entity/data.py:79-210) ---------------------------------------------------------------
  @staticmethod
  def _env(env_ids):
    if env_ids is None:
      return slice(None)
    return env_ids[:, None] if isinstance(env_ids, torch.Tensor) else env_ids

  def write_root_state(self, root_state, env_ids=None):
    if self.is_fixed_base:
      raise ValueError("Cannot write root state for fixed-base entity.")
    assert root_state.shape[-1] == self.ROOT_STATE_DIM
    self.write_root_pose(root_state[:, :7], env_ids)
    self.write_root_velocity(root_state[:, 7:], env_ids)

  def write_root_pose(self, pose, env_ids=None):
    if self.is_fixed_base:
      raise ValueError("Cannot write root pose for fixed-base entity.")
    assert pose.shape[-1] == self.ROOT_POSE_DIM
    self.data.qpos[self._env(env_ids), self.indexing.free_joint_q_adr] = pose

  def write_root_velocity(self, velocity, env_ids=None):
    """World-frame [lin, ang] -> qvel; the angular part is stored in the body frame of the quaternion currently in qpos."""
    if self.is_fixed_base:
      raise ValueError("Cannot write root velocity for fixed-base entity.")
    assert velocity.shape[-1] == self.ROOT_VEL_DIM
    e = self._env(env_ids)
    quat = self.data.qpos[e, self.indexing.free_joint_q_adr[3:7]]
    self.data.qvel[e, self.indexing.free_joint_v_adr] = torch.cat([velocity[:, :3], quat_apply_inverse(quat, velocity[:, 3:])], -1)

  def write_joint_state(self, position, velocity, joint_ids=None, env_ids=None):
    if not self.is_articulated:
      raise ValueError("Cannot write joint state for non-articulated entity.")
    self.write_joint_position(position, joint_ids, env_ids)
    self.write_joint_velocity(velocity, joint_ids, env_ids)

  def write_joint_position(self, position, joint_ids=None, env_ids=None):
    if not self.is_articulated:
      raise ValueError("Cannot write joint position for non-articulated entity.")
    self.data.qpos[self._env(env_ids), self.indexing.joint_q_adr[slice(None) if joint_ids is None else joint_ids]] = position

  def write_joint_velocity(self, velocity, joint_ids=None, env_ids=None):
    if not self.is_articulated:
      raise ValueError("Cannot write joint velocity for non-articulated entity.")
    self.data.qvel[self._env(env_ids), self.indexing.joint_v_adr[slice(None) if joint_ids is None else joint_ids]] = velocity

  def write_external_wrench(self, force, torque, body_ids=None, env_ids=None):
    e = self._env(env_ids)
    gb = self.indexing.body_ids[slice(None) if body_ids is None else body_ids]
    if force is not None:
      self.data.xfrc_applied[e, gb, 0:3] = force
    if torque is not None:
      self.data.xfrc_applied[e, gb, 3:6] = torque

  def write_ctrl(self, ctrl, ctrl_ids=None, env_ids=None):
    if not self.is_actuated:
      raise ValueError("Cannot write control for non-actuated entity.")
    self.data.ctrl[self._env(env_ids), self.indexing.ctrl_ids[slice(None) if ctrl_ids is None else ctrl_ids]] = ctrl

  def write_mocap_pose(self, pose, env_ids=None):
    if self.indexing.mocap_id is None:
      raise ValueError("Cannot write mocap pose for non-mocap entity.")
    assert pose.shape[-1] == self.ROOT_POSE_DIM
    e = self._env(env_ids)
    self.data.mocap_pos[e, self.indexing.mocap_id] = pose[:, 0:3].unsqueeze(1)
    self.data.mocap_quat[e, self.indexing.mocap_id] = pose[:, 3:7].unsqueeze(1)

  def clear_state(self, env_ids=None):
    if self.is_actuated:
      e = slice(None) if env_ids is None else env_ids
      self.joint_pos_target[e] = 0.0
      self.joint_vel_target[e] = 0.0
      self.joint_effort_target[e] = 0.0

  # ---- reads (entity/data.py:212-533) ----------------------------------------------------------------
  def _root(self, field):
    return getattr(self.data, field)[:, self.indexing.root_body_id]

  def _vel(self, pos, body_ids):
    com = self._root("subtree_com")
    return velocity_from_cvel(pos, com.unsqueeze(1) if pos.dim() == 3 else com, self.data.cvel[:, body_ids])

  @property
  def root_link_pose_w(self):
    return torch.cat([self._root("xpos"), self._root("xquat")], -1)

  @property
  def root_link_vel_w(self):
    return self._vel(self._root("xpos"), self.indexing.root_body_id)

  @property
  def root_com_pose_w(self):
    iq = self.model.body_iquat[:, self.indexing.root_body_id]
    return torch.cat([self._root("xipos"), quat_mul(self._root("xquat"), iq)], -1)

  @property
  def root_com_vel_w(self):
    return self._vel(self._root("xipos"), self.indexing.root_body_id)

  @property
  def body_link_pose_w(self):
    b = self.indexing.body_ids
    return torch.cat([self.data.xpos[:, b], self.data.xquat[:, b]], -1)

  @property
  def body_link_vel_w(self):
    b = self.indexing.body_ids
    return self._vel(self.data.xpos[:, b], b)

  @property
  def body_com_pose_w(self):
    b = self.indexing.body_ids
    return torch.cat([self.data.xipos[:, b], quat_mul(self.data.xquat[:, b], self.model.body_iquat[:, b])], -1)

  @property
  def body_com_vel_w(self):
    b = self.indexing.body_ids
    return self._vel(self.data.xipos[:, b], b)

  @property
  def body_external_wrench(self):
    return self.data.xfrc_applied[:, self.indexing.body_ids]

  @property
  def geom_pose_w(self):
    g = self.indexing.geom_ids
    return torch.cat([self.data.geom_xpos[:, g], quat_from_matrix(self.data.geom_xmat[:, g])], -1)

  @property
  def geom_vel_w(self):
    g = self.indexing.geom_ids
    return self._vel(self.data.geom_xpos[:, g], self._static_ids("geom_bodyid")[g])

  @property
  def site_pose_w(self):
    s = self.indexing.site_ids
    return torch.cat([self.data.site_xpos[:, s], quat_from_matrix(self.data.site_xmat[:, s])], -1)

  @property
  def site_vel_w(self):
    s = self.indexing.site_ids
    return self._vel(self.data.site_xpos[:, s], self._static_ids("site_bodyid")[s])

  def _static_ids(self, name):
    v = getattr(self.model, name)
    v = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v
    if v.dim() == 2:  # bridge views carry a leading world axis
      v = v[0]
    return v.to(device=self.device, dtype=torch.long)

  @property
  def joint_pos(self):
    return self.data.qpos[:, self.indexing.joint_q_adr]

  @property
  def joint_pos_biased(self):
    return self.joint_pos + self.encoder_bias

  @property
  def joint_vel(self):
    return self.data.qvel[:, self.indexing.joint_v_adr]

  @property
  def joint_acc(self):
    return self.data.qacc[:, self.indexing.joint_v_adr]

  @property
  def joint_torques(self):
    raise NotImplementedError("Joint torques are not available; use 'actuator_force' or 'generalized_force'.")

  @property
  def actuator_force(self):
    return self.data.actuator_force[:, self.indexing.ctrl_ids]

  @property
  def generalized_force(self):
    return self.data.qfrc_applied[:, self.indexing.free_joint_v_adr]

  @property
  def projected_gravity_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.gravity_vec_w)

  @property
  def heading_w(self):
    f = quat_apply(self.root_link_quat_w, self.forward_vec_b)
    return torch.atan2(f[:, 1], f[:, 0])

  @property
  def root_link_lin_vel_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.root_link_lin_vel_w)

  @property
  def root_link_ang_vel_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.root_link_ang_vel_w)

  @property
  def root_com_lin_vel_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.root_com_lin_vel_w)

  @property
  def root_com_ang_vel_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.root_com_ang_vel_w)


def _component(name: str, src: str, lo: int, hi: int):
  setattr(EntityData, name, property(lambda self: getattr(self, src)[..., lo:hi]))


# component accessors (entity/data.py:376-496): <prefix>_{pos,quat}_w from *_pose_w, <prefix>_{lin,ang}_vel_w from *_vel_w
for _p in ("root_link", "root_com", "body_link", "body_com", "geom", "site"):
  _component(f"{_p}_pos_w", f"{_p}_pose_w", 0, 3)
  _component(f"{_p}_quat_w", f"{_p}_pose_w", 3, 7)
  _component(f"{_p}_lin_vel_w", f"{_p}_vel_w", 0, 3)
  _component(f"{_p}_ang_vel_w", f"{_p}_vel_w", 3, 6)
_component("body_external_force", "body_external_wrench", 0, 3)
_component("body_external_torque", "body_external_wrench", 3, 6)


class Entity:
  """One articulated robot inside the compiled scene: names, index maps, `data`, and the `write_*_to_sim` API."""

  def __init__(self, name: str, sim, cfg=None, prefix: str | None = None, device: str | None = None,
               joint_pos_target=None, encoder_bias=None, actuator_groups=None):
    m = sim.mj_model
    device = device or sim.device
    pre = f"{name}/" if prefix is None else prefix
    self.name, self._sim = name, sim
    sel = lambda names: [i for i, n in enumerate(names) if n.startswith(pre)]  # noqa: E731
    body_ids, geom_ids, site_ids = [i for i in sel(m.body_names) if i != 0], sel(m.geom_names), sel(m.site_names)   # never the world body
    jall = [j for j in range(m.njnt) if m.jnt_names[j].startswith(pre)]
    free = [j for j in jall if m.jnt_type[j] == 0]
    joints = [j for j in jall if m.jnt_type[j] != 0]
    self.body_names = [m.body_names[i].removeprefix(pre) for i in body_ids]
    self.geom_names = [m.geom_names[i].removeprefix(pre) for i in geom_ids]
    self.site_names = [m.site_names[i].removeprefix(pre) for i in site_ids]
    self.joint_names = [m.jnt_names[j].removeprefix(pre) for j in joints]
    act_of_joint = {int(m.actuator_trnid[a, 0]): a for a in range(m.nu)}
    ctrl = [act_of_joint[j] for j in joints if j in act_of_joint]
    self.actuator_names = [self.joint_names[k] for k, j in enumerate(joints) if j in act_of_joint]
    L = lambda a: torch.as_tensor(np.asarray(a, np.int64), device=device)  # noqa: E731
    fq = int(m.jnt_qposadr[free[0]]) if free else 0
    fv = int(m.jnt_dofadr[free[0]]) if free else 0
    self.indexing = EntityIndexing(
      body_ids=L(body_ids), geom_ids=L(geom_ids), site_ids=L(site_ids), joint_ids=L(joints), ctrl_ids=L(ctrl),
      joint_q_adr=L(m.jnt_qposadr[joints]), joint_v_adr=L(m.jnt_dofadr[joints]),
      free_joint_q_adr=L(np.arange(fq, fq + 7) if free else []), free_joint_v_adr=L(np.arange(fv, fv + 6) if free else []))
    n = sim.num_envs
    kq = np.asarray(m.key_qpos[0] if getattr(m, "nkey", len(m.key_qpos)) else m.qpos0, np.float32)
    T = lambda a: torch.as_tensor(np.asarray(a, np.float32), device=device)  # noqa: E731
    root = np.concatenate([kq[fq:fq + 7], np.zeros(6, np.float32)]) if free else np.zeros(13, np.float32)
    lim = np.asarray(m.jnt_range[joints], np.float32)
    mean, rng = 0.5 * (lim[:, 0] + lim[:, 1]), lim[:, 1] - lim[:, 0]
    soft = np.stack([mean - 0.45 * rng, mean + 0.45 * rng], 1)  # soft_joint_pos_limit_factor 0.9 (entity/entity.py:409-462)
    self.data = EntityData(
      self.indexing, sim.data, sim.model, device, default_root_state=T(root).repeat(n, 1),
      default_joint_pos=T(kq[m.jnt_qposadr[joints]]).repeat(n, 1), default_joint_vel=torch.zeros(n, len(joints), device=device),
      joint_pos_limits=T(lim).repeat(n, 1, 1), soft_joint_pos_limits=T(soft).repeat(n, 1, 1), is_fixed_base=not free,
      is_articulated=bool(joints), is_actuated=bool(ctrl), joint_pos_target=joint_pos_target, encoder_bias=encoder_bias)
    # `entity.actuators` (entity/entity.py:160-206): one object per actuator cfg.  `actuator_groups`: the joint-name patterns of
    # every cfg, in cfg order (robots.g1_cfg / go1_cfg); None: all builtin actuators as one group.
    from .actuators import BuiltinPositionActuator
    from .robots import resolve_matching_names
    local_of = {n: k for k, n in enumerate(self.joint_names)}
    ctrl_of = {n: c for n, c in zip(self.actuator_names, ctrl)}
    groups = [tuple(g) if not isinstance(g, str) else (g,) for g in actuator_groups] if actuator_groups else [(".*",)]
    self.actuators = []
    for g in groups:
      _, names = resolve_matching_names(g, self.actuator_names)
      if names:
        self.actuators.append(BuiltinPositionActuator(L([ctrl_of[n] for n in names]), L([local_of[n] for n in names]), names))

  # properties (entity/entity.py:209-262)
  @property
  def is_fixed_base(self) -> bool:
    return self.data.is_fixed_base

  @property
  def is_articulated(self) -> bool:
    return self.data.is_articulated

  @property
  def is_actuated(self) -> bool:
    return self.data.is_actuated

  @property
  def num_joints(self) -> int:
    return len(self.joint_names)

  @property
  def num_bodies(self) -> int:
    return len(self.body_names)

  @property
  def num_geoms(self) -> int:
    return len(self.geom_names)

  @property
  def num_sites(self) -> int:
    return len(self.site_names)

  @property
  def num_actuators(self) -> int:
    return len(self.actuator_names)

  def find_bodies(self, keys, preserve_order=False):
    return _resolve(keys, self.body_names, preserve_order)

  def find_joints(self, keys, joint_subset: Sequence[str] | None = None, preserve_order=False):
    return _resolve(keys, list(joint_subset) if joint_subset is not None else self.joint_names, preserve_order)

  def find_geoms(self, keys, preserve_order=False):
    return _resolve(keys, self.geom_names, preserve_order)

  def find_sites(self, keys, preserve_order=False):
    return _resolve(keys, self.site_names, preserve_order)

  def find_actuators(self, keys, preserve_order=False):
    return _resolve(keys, self.actuator_names, preserve_order)

  # state writes (entity/entity.py:553-640)
  def write_root_state_to_sim(self, root_state, env_ids=None):
    self.data.write_root_state(root_state, env_ids)

  def write_root_link_pose_to_sim(self, pose, env_ids=None):
    self.data.write_root_pose(pose, env_ids)

  def write_root_link_velocity_to_sim(self, velocity, env_ids=None):
    self.data.write_root_velocity(velocity, env_ids)

  def write_joint_state_to_sim(self, position, velocity, joint_ids=None, env_ids=None):
    self.data.write_joint_state(position, velocity, joint_ids, env_ids)

  def write_joint_position_to_sim(self, position, joint_ids=None, env_ids=None):
    self.data.write_joint_position(position, joint_ids, env_ids)

  def write_joint_velocity_to_sim(self, velocity, joint_ids=None, env_ids=None):
    self.data.write_joint_velocity(velocity, joint_ids, env_ids)

  def write_external_wrench_to_sim(self, forces, torques, env_ids=None, body_ids=None):
    self.data.write_external_wrench(forces, torques, body_ids, env_ids)

  def write_mocap_pose_to_sim(self, pose, env_ids=None):
    self.data.write_mocap_pose(pose, env_ids)

  def set_joint_position_target(self, position, joint_ids=None, env_ids=None):
    e = slice(None) if env_ids is None else env_ids
    j = slice(None) if joint_ids is None else joint_ids
    self.data.joint_pos_target[e, j] = position

  def set_joint_velocity_target(self, velocity, joint_ids=None, env_ids=None):
    e = slice(None) if env_ids is None else env_ids
    j = slice(None) if joint_ids is None else joint_ids
    self.data.joint_vel_target[e, j] = velocity

  def set_joint_effort_target(self, effort, joint_ids=None, env_ids=None):
    e = slice(None) if env_ids is None else env_ids
    j = slice(None) if joint_ids is None else joint_ids
    self.data.joint_effort_target[e, j] = effort

  def write_data_to_sim(self) -> None:
    """Builtin position actuators: ctrl <- joint position target (actuator/builtin_group.py:68-81, entity.py:815-825)."""
    if self.is_actuated:
      self.data.write_ctrl(self.data.joint_pos_target)

  def reset(self, env_ids=None) -> None:
    self.data.clear_state(env_ids)
