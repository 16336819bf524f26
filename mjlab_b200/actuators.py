"""Custom (torch-side) actuator models on the B200 engine.

Host-side mirror of mjlab's explicit actuators: same class names, constructor meaning, `compute(cmd)` contract
(`ActuatorCmd` in, effort [num_envs, num_joints] out), gain setters and reset behaviour:

  IdealPdActuator      src/mjlab/actuator/pd_actuator.py:60-131      -> b2_actuator_pd
  DcMotorActuator      src/mjlab/actuator/dc_actuator.py:60-162      -> b2_actuator_pd (with the torque-speed clip)
  DelayBuffer          src/mjlab/utils/buffers/delay_buffer.py:100-303 -> b2_delay_step
  DelayedActuator      src/mjlab/actuator/delayed_actuator.py:68-175
  LearnedMlpActuator   src/mjlab/actuator/learned_actuator.py:69-207 -> b2_actuator_mlp

The arithmetic runs in the CUDA kernels behind the C ABI (include/b200env.h); torch only owns the buffers.  The shipped
velocity tasks use MuJoCo-side position actuators (`BuiltinPositionActuator`), which live inside k_smooth, so none of this
is on their hot path.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Literal, Sequence

import torch

from . import engine as E
from . import envabi as A


@dataclass
class ActuatorCmd:
  """Targets and current joint state, all [num_envs, num_joints] (actuator/actuator.py:52-69)."""

  position_target: torch.Tensor
  velocity_target: torch.Tensor
  effort_target: torch.Tensor
  joint_pos: torch.Tensor
  joint_vel: torch.Tensor


def _ptr(t: torch.Tensor | None):
  return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
  return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


def _stream(dev) -> int:
  return torch.cuda.current_stream(dev).cuda_stream


# Every random draw of this module goes through these two hooks, in the same order and with the same shapes as the
# reference's DelayBuffer (utils/buffers/delay_buffer.py:158-166,213-222,276-301).  tests/test_gpu_actuators.py replaces them
# to replay the draws recorded from the reference's own class (tools/make_actuator_golden.py).
def _randint(low: int, high: int, shape, device, generator=None) -> torch.Tensor:
  return torch.randint(low, high, shape, dtype=torch.long, device=device, generator=generator)


def _rand(shape, device, generator=None) -> torch.Tensor:
  return torch.rand(shape, dtype=torch.float32, device=device, generator=generator)


class BuiltinPositionActuator:
  """View of one group of MuJoCo-builtin position actuators (actuator/builtin_actuator.py:26-76): the torque law runs inside
  the physics kernels (`ctrl` -> `actuator_gainprm` / `actuator_biasprm` / `actuator_forcerange`), so the object only names the
  group - `ctrl_ids` (global actuator ids), `joint_ids` (entity-local), `joint_names` - for the event terms that edit its gains
  and limits in the model (`events.randomize_pd_gains`, `events.randomize_effort_limits`)."""

  def __init__(self, ctrl_ids: torch.Tensor, joint_ids: torch.Tensor, joint_names: Sequence[str]):
    self.ctrl_ids, self.joint_ids, self.joint_names = ctrl_ids, joint_ids, list(joint_names)

  def __repr__(self) -> str:
    return f"BuiltinPositionActuator({self.joint_names})"


class IdealPdActuator:
  """tau = clamp(kp (q* - q) + kd (qd* - qd) + tau_ff, +-effort_limit), per-env gains (pd_actuator.py:94-131)."""

  def __init__(self, num_envs: int, num_joints: int, stiffness: float, damping: float, effort_limit: float = float("inf"),
               device: str = "cuda:0"):
    if not str(device).startswith("cuda"):
      raise RuntimeError("mjlab_b200 actuators require a CUDA device (there is no CPU fallback)")
    self._lib = A.load_library()
    self.device = torch.device(device)
    full = lambda v: torch.full((num_envs, num_joints), float(v), dtype=torch.float32, device=self.device)  # noqa: E731
    self.stiffness, self.damping, self.force_limit = full(stiffness), full(damping), full(effort_limit)
    self.default_stiffness, self.default_damping = self.stiffness.clone(), self.damping.clone()
    self.default_force_limit = self.force_limit.clone()
    self._effort = torch.zeros((num_envs, num_joints), dtype=torch.float32, device=self.device)
    self.num_envs, self.num_joints = num_envs, num_joints
    self.ctrl_ids = torch.arange(num_joints, device=self.device)   # overwritten by the owner when the group is a subset of an entity's actuators

  def _dc(self) -> tuple[torch.Tensor | None, torch.Tensor | None]:
    return None, None

  def compute(self, cmd: ActuatorCmd) -> torch.Tensor:
    sat, vlim = self._dc()
    args = [_f32c(t) for t in (cmd.position_target, cmd.velocity_target, cmd.effort_target, cmd.joint_pos, cmd.joint_vel)]
    E.check_call(self._lib, self._lib.b2_actuator_pd(
      self._effort.data_ptr(), *[t.data_ptr() for t in args], self.stiffness.data_ptr(), self.damping.data_ptr(),
      self.force_limit.data_ptr(), _ptr(sat), _ptr(vlim), self._effort.numel(), _stream(self.device)))
    return self._effort

  def set_gains(self, env_ids, kp: torch.Tensor | None = None, kd: torch.Tensor | None = None) -> None:
    """kp / kd: (num_env_ids, num_joints), or (num_env_ids,) = one gain per env (pd_actuator.py:111-135)."""
    if kp is not None:
      self.stiffness[env_ids] = kp.unsqueeze(-1) if kp.ndim == 1 else kp
    if kd is not None:
      self.damping[env_ids] = kd.unsqueeze(-1) if kd.ndim == 1 else kd

  def set_effort_limit(self, env_ids, effort_limit: torch.Tensor) -> None:
    """effort_limit: (num_env_ids, num_joints) or (num_env_ids,) (pd_actuator.py:137-151)."""
    self.force_limit[env_ids] = effort_limit.unsqueeze(-1) if effort_limit.ndim == 1 else effort_limit

  def reset(self, env_ids=None) -> None:
    pass


class DcMotorActuator(IdealPdActuator):
  """PD torque clipped by the DC-motor torque-speed line, then by the continuous limit (dc_actuator.py:136-162)."""

  def __init__(self, num_envs: int, num_joints: int, stiffness: float, damping: float, effort_limit: float,
               saturation_effort: float, velocity_limit: float, device: str = "cuda:0"):
    super().__init__(num_envs, num_joints, stiffness, damping, effort_limit, device)
    full = lambda v: torch.full((num_envs, num_joints), float(v), dtype=torch.float32, device=self.device)  # noqa: E731
    self.saturation_effort, self.velocity_limit_motor = full(saturation_effort), full(velocity_limit)

  def _dc(self):
    return self.saturation_effort, self.velocity_limit_motor


class DelayBuffer:
  """Per-env lagged read of a signal; lag resampling policy as in utils/buffers/delay_buffer.py:229-303."""

  def __init__(self, min_lag: int, max_lag: int, batch_size: int, device: str = "cuda:0", per_env: bool = True,
               hold_prob: float = 0.0, update_period: int = 0, per_env_phase: bool = True, generator: torch.Generator | None = None):
    if min_lag < 0 or max_lag < min_lag:
      raise ValueError(f"Invalid lag range [{min_lag}, {max_lag}]")
    if not 0.0 <= hold_prob <= 1.0:
      raise ValueError("hold_prob must be in [0, 1]")
    self._lib = A.load_library()
    self.min_lag, self.max_lag, self.batch_size, self.device = min_lag, max_lag, batch_size, torch.device(device)
    self.per_env, self.hold_prob, self.update_period, self.per_env_phase, self.generator = per_env, hold_prob, update_period, per_env_phase, generator
    self._cap = max_lag + 1 if max_lag > 0 else 1
    z = lambda dt: torch.zeros(batch_size, dtype=dt, device=self.device)  # noqa: E731
    self._count, self._head, self._lag32 = z(torch.int32), z(torch.int32), z(torch.int32)
    self._current_lags, self._step_count = z(torch.long), z(torch.long)
    self._phase = self._rand_phase() if (update_period > 0 and per_env_phase) else z(torch.long)
    self._ring: torch.Tensor | None = None
    self._out: torch.Tensor | None = None
    self._pending: torch.Tensor | None = None

  def _rand_phase(self) -> torch.Tensor:
    return _randint(0, self.update_period, (self.batch_size,), self.device, self.generator)

  @property
  def is_initialized(self) -> bool:
    return self._ring is not None

  @property
  def current_lags(self) -> torch.Tensor:
    return self._current_lags

  def set_lags(self, lags: torch.Tensor, batch_ids=None) -> None:
    idx = slice(None) if batch_ids is None else batch_ids
    self._current_lags[idx] = torch.as_tensor(lags, device=self.device).clamp(self.min_lag, self.max_lag)

  def reset(self, batch_ids=None) -> None:
    idx = slice(None) if batch_ids is None else batch_ids
    self._count[idx] = 0
    self._current_lags[idx] = 0
    self._step_count[idx] = 0
    if self._ring is not None:
      self._ring[idx] = 0.0
    if self.update_period > 0 and self.per_env_phase:
      self._phase[idx] = self._rand_phase()[idx]

  def append(self, data: torch.Tensor) -> None:
    if data.shape[0] != self.batch_size:
      raise ValueError(f"Expected batch size {self.batch_size}, got {data.shape[0]}")
    self._pending = _f32c(data)
    if self._ring is None:
      width = self._pending[0].numel()
      self._ring = torch.zeros((self.batch_size, self._cap, width), dtype=torch.float32, device=self.device)
      self._out = torch.zeros_like(self._pending)

  def _update_lags(self) -> None:
    if self.update_period > 0:
      should = (self._step_count + self._phase) % self.update_period == 0
    else:
      should = torch.ones(self.batch_size, dtype=torch.bool, device=self.device)
    if self.per_env:
      cand = _randint(self.min_lag, self.max_lag + 1, (self.batch_size,), self.device, self.generator)
    else:
      cand = _randint(self.min_lag, self.max_lag + 1, (1,), self.device, self.generator).expand(self.batch_size)
    if self.hold_prob > 0.0:
      should = should & (_rand((self.batch_size,), self.device, self.generator) >= self.hold_prob)
    self._current_lags = torch.where(should, cand, self._current_lags)
    self._step_count += 1

  def compute(self) -> torch.Tensor:
    """Push the pending sample and return, per env, the sample `lag` pushes ago (clamped to what has been pushed)."""
    if self._ring is None or self._pending is None:
      raise RuntimeError("Buffer not initialized. Call append() first.")
    self._update_lags()
    self._lag32.copy_(self._current_lags)
    assert self._out is not None
    E.check_call(self._lib, self._lib.b2_delay_step(
      self._out.data_ptr(), self._ring.data_ptr(), self._count.data_ptr(), self._head.data_ptr(), self._lag32.data_ptr(),
      self._pending.data_ptr(), self.batch_size, self._cap, self._ring.shape[2], _stream(self.device)))
    self._pending = None
    return self._out


class DelayedActuator:
  """Delays the chosen target(s) before the wrapped actuator sees them (delayed_actuator.py:68-150)."""

  def __init__(self, base_actuator, delay_target: str | Sequence[str] = "position", delay_min_lag: int = 0, delay_max_lag: int = 0,
               delay_hold_prob: float = 0.0, delay_update_period: int = 0, delay_per_env_phase: bool = True):
    self.base_actuator = base_actuator
    targets = (delay_target,) if isinstance(delay_target, str) else tuple(delay_target)
    for t in targets:
      if t not in ("position", "velocity", "effort"):
        raise ValueError(f"Unknown delay target '{t}'")
    self._delay_buffers = {t: DelayBuffer(delay_min_lag, delay_max_lag, base_actuator.num_envs, str(base_actuator.device),
                                          hold_prob=delay_hold_prob, update_period=delay_update_period,
                                          per_env_phase=delay_per_env_phase) for t in targets}

  def compute(self, cmd: ActuatorCmd) -> torch.Tensor:
    t = {"position": cmd.position_target, "velocity": cmd.velocity_target, "effort": cmd.effort_target}
    for name, buf in self._delay_buffers.items():
      buf.append(t[name])
      t[name] = buf.compute().clone()
    return self.base_actuator.compute(ActuatorCmd(t["position"], t["velocity"], t["effort"], cmd.joint_pos, cmd.joint_vel))

  def set_lags(self, lags: torch.Tensor, env_ids=None) -> None:
    for buf in self._delay_buffers.values():
      buf.set_lags(lags, env_ids)

  def reset(self, env_ids=None) -> None:
    for buf in self._delay_buffers.values():
      buf.reset(env_ids)
    self.base_actuator.reset(env_ids)


_ACT = {"elu": 0, "tanh": 1, "relu": 2, "softsign": 3, "identity": 4}


class LearnedMlpActuator(DcMotorActuator):
  """MLP on the histories of position error and velocity, one evaluation per (env, joint), DC-motor clipped output
  (learned_actuator.py:144-207).  `network` is a sequence of torch.nn.Linear layers (activations between them are given by
  `activation`) or a path to a TorchScript / pickled module from which the Linear parameters are taken in order."""

  def __init__(self, num_envs: int, num_joints: int, network, pos_scale: float, vel_scale: float, torque_scale: float,
               effort_limit: float, saturation_effort: float, velocity_limit: float, input_order: Literal["pos_vel", "vel_pos"] = "pos_vel",
               history_length: int = 3, activation: str = "softsign", device: str = "cuda:0"):
    super().__init__(num_envs, num_joints, 0.0, 0.0, effort_limit, saturation_effort, velocity_limit, device)
    if input_order not in ("pos_vel", "vel_pos"):
      raise ValueError(f"Invalid input order: {input_order}. Must be 'pos_vel' or 'vel_pos'.")
    if isinstance(network, str):
      network = torch.jit.load(network, map_location="cpu")
    params = [p.detach() for _, p in network.named_parameters()]
    ws = [p for p in params if p.dim() == 2]
    bs = [p for p in params if p.dim() == 1]
    if len(bs) not in (0, len(ws)):
      raise ValueError("every Linear layer must either have a bias or none of them")
    self._w = [w.to(self.device, torch.float32).contiguous() for w in ws]
    self._b = [b.to(self.device, torch.float32).contiguous() for b in bs]
    if self._w[0].shape[1] != 2 * history_length:
      raise ValueError(f"network expects {self._w[0].shape[1]} inputs, history_length={history_length} provides {2 * history_length}")
    self.pos_scale, self.vel_scale, self.torque_scale = pos_scale, vel_scale, torque_scale
    self.input_order, self.history_length, self._act = input_order, history_length, _ACT[activation]
    self._pos_hist = torch.zeros((num_envs, num_joints, history_length), dtype=torch.float32, device=self.device)
    self._vel_hist = torch.zeros_like(self._pos_hist)
    self._pushes = torch.zeros(num_envs, dtype=torch.long, device=self.device)
    n = len(self._w)
    self._wp = (C.c_void_p * n)(*[w.data_ptr() for w in self._w])
    self._bp = (C.c_void_p * n)(*[b.data_ptr() for b in self._b]) if self._b else None
    self._lin = (C.c_int * n)(*[w.shape[1] for w in self._w])
    self._lout = (C.c_int * n)(*[w.shape[0] for w in self._w])

  def reset(self, env_ids=None) -> None:
    idx = slice(None) if env_ids is None else env_ids
    self._pos_hist[idx] = 0.0
    self._vel_hist[idx] = 0.0
    self._pushes[idx] = 0

  def _push(self, hist: torch.Tensor, value: torch.Tensor, first: torch.Tensor) -> None:
    """index 0 = newest; the first push after a reset back-fills every slot (circular_buffer.py:190-214)."""
    hist[:, :, 1:] = hist[:, :, :-1].clone()
    hist[:, :, 0] = value
    if bool(first.any()):
      hist[first] = value[first].unsqueeze(-1).expand(-1, -1, self.history_length)

  def compute(self, cmd: ActuatorCmd) -> torch.Tensor:
    first = self._pushes == 0
    self._push(self._pos_hist, cmd.position_target - cmd.joint_pos, first)
    self._push(self._vel_hist, cmd.joint_vel, first)
    self._pushes += 1
    p = (self._pos_hist * self.pos_scale).reshape(-1, self.history_length)
    v = (self._vel_hist * self.vel_scale).reshape(-1, self.history_length)
    x = torch.cat([p, v] if self.input_order == "pos_vel" else [v, p], dim=1).contiguous()
    vel = _f32c(cmd.joint_vel)
    E.check_call(self._lib, self._lib.b2_actuator_mlp(
      self._effort.data_ptr(), x.data_ptr(), self._wp, self._bp, self._lin, self._lout, len(self._w), self._act,
      float(self.torque_scale), self.force_limit.data_ptr(), self.saturation_effort.data_ptr(), self.velocity_limit_motor.data_ptr(),
      vel.data_ptr(), self._effort.numel(), _stream(self.device)))
    return self._effort
