"""rsl_rl vec-env adapter for the B200 env.

Mirrors `mjlab.rl.RslRlVecEnvWrapper` (src/mjlab/rl/vecenv_wrapper.py:11-111): action clamp, observation dict with the
env batch size, `dones` as int64, `extras["time_outs"]`, `episode_length_buf` get/set passthrough (rsl_rl randomises it
at start-up, scripts/train.py:167-169), and the reset at construction.  `tensordict` / `rsl_rl` are optional: when
`tensordict` is importable the observations are wrapped in a TensorDict, otherwise in a plain dict subclass carrying
`batch_size`.
"""

from __future__ import annotations

from typing import Any

import torch

try:  # pragma: no cover - optional dependency
  from tensordict import TensorDict as _TensorDict
except Exception:  # noqa: BLE001
  _TensorDict = None


class ObsDict(dict):
  """Minimal TensorDict stand-in: a dict of tensors with a `batch_size` and `.to()`."""

  def __init__(self, data: dict[str, Any], batch_size: list[int]):
    super().__init__(data)
    self.batch_size = torch.Size(batch_size)

  def to(self, *args, **kwargs) -> "ObsDict":
    return ObsDict({k: v.to(*args, **kwargs) for k, v in self.items()}, list(self.batch_size))


def _wrap(obs: dict[str, Any], n: int):
  if _TensorDict is not None:
    return _TensorDict(obs, batch_size=[n])
  return ObsDict(obs, [n])


class RslRlVecEnvWrapper:
  def __init__(self, env, clip_actions: float | None = None):
    self.env = env
    self.clip_actions = clip_actions
    self.num_envs = env.num_envs
    self.device = torch.device(env.device)
    self.max_episode_length = env.max_episode_length
    self.num_actions = env.action_manager.total_action_dim
    self.env.reset()  # rsl_rl does not call reset itself

  cfg = property(lambda self: self.env.cfg)
  render_mode = property(lambda self: self.env.render_mode)
  unwrapped = property(lambda self: self.env)

  @classmethod
  def class_name(cls) -> str:
    return cls.__name__

  @property
  def episode_length_buf(self) -> torch.Tensor:
    return self.env.episode_length_buf

  @episode_length_buf.setter
  def episode_length_buf(self, value: torch.Tensor) -> None:
    self.env.episode_length_buf = value

  def seed(self, seed: int = -1) -> int:
    return self.env.seed(seed)

  def get_observations(self):
    return _wrap(self.env.observation_manager.compute(), self.num_envs)

  def reset(self):
    obs, extras = self.env.reset()
    return _wrap(obs, self.num_envs), extras

  def step(self, actions: torch.Tensor):
    if self.clip_actions is not None:
      actions = torch.clamp(actions, -self.clip_actions, self.clip_actions)
    obs, rew, terminated, truncated, extras = self.env.step(actions)
    dones = (terminated | truncated).to(dtype=torch.long)
    if not getattr(self.env.cfg, "is_finite_horizon", False):
      extras["time_outs"] = truncated
    return _wrap(obs, self.num_envs), rew, dones, extras

  def close(self) -> None:
    return self.env.close()
