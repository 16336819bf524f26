"""Sensor view objects over the engine's `sensordata`.

  BuiltinSensor   src/mjlab/sensor/builtin_sensor.py:327-340   `.data` = the sensor's slice of sensordata, [B, dim] (a view)
  ContactSensor   src/mjlab/sensor/contact_sensor.py:231-367   `.data` = ContactData(found [B, N*slots], force [B, N*slots, 3], ...
                                                                + air-time state), `compute_first_contact / compute_first_air`,
                                                                `reset`, `update`

Contact matching, reduction and the per-contact force decode are done by the physics kernels (k_post) into sensordata; the
sensors of a compiled scene are discovered by name, `<cfg.name>_<primary>_<field>` (contact_sensor.py:150-204).  Inside
`ManagerBasedRlEnv` the air-time state lives in the env's buffers and is advanced by the fused `k_env_airtime` kernel after every
physics step, so `update()` is a no-op there; a stand-alone sensor (only a `Simulation`) tracks it with the reference's torch
update rule.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

SENS_CONTACT = 42
_FIELD_DIMS = {"found": 1, "force": 3, "torque": 3, "dist": 1, "pos": 3, "normal": 3, "tangent": 3}
_BIT_FIELD = {0: "found", 1: "force", 2: "torque", 3: "dist", 4: "pos", 5: "normal", 6: "tangent"}
_REDUCE = {0: "none", 1: "mindist", 2: "maxforce", 3: "netforce"}


class BuiltinSensor:
  """One model sensor (gyro, velocimeter, accelerometer, subtreeangmom, joint / frame sensors, ...) by name."""

  def __init__(self, name: str, sim):
    names = list(sim.mj_model.sensor_names)
    if name not in names:
      raise KeyError(f"sensor '{name}' not found in the model; available: {names}")
    i = names.index(name)
    self.name = name
    self._adr, self._dim = int(sim.mj_model.sensor_adr[i]), int(sim.mj_model.sensor_dim[i])
    self._view = sim.data.sensordata[:, self._adr: self._adr + self._dim]

  @property
  def data(self) -> torch.Tensor:
    return self._view

  def reset(self, env_ids=None) -> None:
    pass

  def update(self, dt: float) -> None:
    pass


@dataclass
class ContactData:
  """Only the requested fields are populated (contact_sensor.py:121-147)."""

  found: torch.Tensor | None = None
  force: torch.Tensor | None = None
  torque: torch.Tensor | None = None
  dist: torch.Tensor | None = None
  pos: torch.Tensor | None = None
  normal: torch.Tensor | None = None
  tangent: torch.Tensor | None = None
  current_air_time: torch.Tensor | None = None
  last_air_time: torch.Tensor | None = None
  current_contact_time: torch.Tensor | None = None
  last_contact_time: torch.Tensor | None = None


class ContactSensor:
  def __init__(self, name: str, sim, track_air_time: bool = False, air_state: torch.Tensor | None = None,
               air_last_time: torch.Tensor | None = None):
    """`air_state` [B, 4, N] / `air_last_time` [B]: buffers owned by an env whose fused kernel advances them."""
    m = sim.mj_model
    self.name, self._sim = name, sim
    slots: list[tuple[str, str, int, int]] = []   # (primary, field, adr, dim) in model order
    self.num_slots, self.reduce = 1, "maxforce"
    for i, sn in enumerate(m.sensor_names):
      if not (sn.startswith(name + "_") and int(m.sensor_type[i]) == SENS_CONTACT):
        continue
      bits, red, num = (int(v) for v in np.asarray(m.sensor_intprm).reshape(-1, 3)[i])
      field = _BIT_FIELD[bits.bit_length() - 1]
      prim = sn[len(name) + 1: -len(field) - 1]
      slots.append((prim, field, int(m.sensor_adr[i]), int(m.sensor_dim[i])))
      self.num_slots, self.reduce = max(num, 1), _REDUCE[red]
    if not slots:
      raise RuntimeError(f"There was an error initializing contact sensor '{name}'")
    self._slots = slots
    self.primaries = list(dict.fromkeys(p for p, _, _, _ in slots))
    self.fields = tuple(dict.fromkeys(f for _, f, _, _ in slots))
    self._fused = air_state is not None
    self.track_air_time = track_air_time or self._fused
    B, N = sim.num_envs, len(self.primaries)
    if self._fused:
      self._air, self._last_time = air_state, air_last_time
    elif self.track_air_time:
      self._air = torch.zeros((B, 4, N), device=sim.device)
      self._last_time = torch.zeros(B, device=sim.device)
    else:
      self._air = self._last_time = None

  def _extract_sensor_data(self) -> ContactData:
    sd = self._sim.data.sensordata
    chunks: dict[str, list[torch.Tensor]] = {f: [] for f in self.fields}
    for _, field, adr, dim in self._slots:
      chunks[field].append(sd[:, adr: adr + dim].view(sd.shape[0], self.num_slots, _FIELD_DIMS[field]))
    out = ContactData()
    for field, ch in chunks.items():
      cat = torch.cat(ch, dim=1)
      setattr(out, field, cat.squeeze(-1) if cat.size(-1) == 1 else cat)
    return out

  @property
  def data(self) -> ContactData:
    out = self._extract_sensor_data()
    if self._air is not None:
      out.current_air_time, out.last_air_time = self._air[:, 0], self._air[:, 1]
      out.current_contact_time, out.last_contact_time = self._air[:, 2], self._air[:, 3]
    return out

  def reset(self, env_ids=None) -> None:
    if self._air is None:
      return
    ids = slice(None) if env_ids is None else env_ids
    self._air[ids] = 0.0
    self._last_time[ids] = self._sim.data.time[ids]

  def update(self, dt: float) -> None:
    """scene.update(dt) after every physics step (manager_based_rl_env.py:275-280).  No-op when an env's fused kernel owns the
    state; otherwise the reference's rule (contact_sensor.py:327-367)."""
    if self._air is None or self._fused or "found" not in self.fields:
      return
    found = self._extract_sensor_data().found
    now = self._sim.data.time
    el = (now - self._last_time).unsqueeze(-1)
    con = found > 0
    cur_air, last_air, cur_con, last_con = (self._air[:, k] for k in range(4))
    first_contact, first_detached = (cur_air > 0) & con, (cur_con > 0) & ~con
    last_air[:] = torch.where(first_contact, cur_air + el, last_air)
    cur_air[:] = torch.where(~con, cur_air + el, torch.zeros_like(cur_air))
    last_con[:] = torch.where(first_detached, cur_con + el, last_con)
    cur_con[:] = torch.where(con, cur_con + el, torch.zeros_like(cur_con))
    self._last_time[:] = now

  def _need_air(self, what: str):
    if self._air is None:
      raise RuntimeError(f"Sensor '{self.name}' must have track_air_time=True to use {what}")

  def compute_first_contact(self, dt: float, abs_tol: float = 1.0e-8) -> torch.Tensor:
    """[B, N] bool: contacts established within the last dt seconds (contact_sensor.py:260-269)."""
    self._need_air("compute_first_contact")
    t = self._air[:, 2]
    return (t > 0.0) & (t < (dt + abs_tol))

  def compute_first_air(self, dt: float, abs_tol: float = 1.0e-8) -> torch.Tensor:
    self._need_air("compute_first_air")
    t = self._air[:, 0]
    return (t > 0.0) & (t < (dt + abs_tol))
