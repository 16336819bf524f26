"""NaN guard of the physics step: a rolling buffer of (qpos, qvel) snapshots that is written to disk the first time a
world's state turns NaN / Inf.  Same contract as the reference's `NanGuard` (src/mjlab/utils/nan_guard.py:16-181): off by
default and then free; `watch(data)` brackets `Simulation.step`; the check looks at qpos, qvel, qacc and qacc_warmstart;
the dump keeps the last `buffer_size` steps of at most `max_envs_to_dump` offending worlds.

Without MuJoCo there is no `mj_getState` / `.mjb` file: the dump stores `[time?, qpos, qvel]` per world - the layout of
MuJoCo's mjSTATE_PHYSICS for models without activations or plugins - and the model as the `.npz` that
`mjlab_b200.mjcf.ModelDesc.save` writes, which `ModelDesc.load` reads back.
"""

from __future__ import annotations

from collections import deque
from contextlib import contextmanager
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Iterator

import numpy as np
import torch


@dataclass
class NanGuardCfg:
  """src/mjlab/utils/nan_guard.py:16-23."""

  enabled: bool = False
  buffer_size: int = 100
  output_dir: str = "/tmp/mjlab/nan_dumps"
  max_envs_to_dump: int = 5


class NanGuard:
  def __init__(self, cfg: NanGuardCfg, num_envs: int, mj_model) -> None:
    self.enabled = cfg.enabled
    self.num_envs = num_envs
    if not self.enabled:
      return
    self.buffer_size = cfg.buffer_size
    self.output_dir = Path(cfg.output_dir)
    self.max_envs_to_dump = cfg.max_envs_to_dump
    self.buffer: deque = deque(maxlen=self.buffer_size)
    self.step_counter = 0
    self._dumped = False
    self.mj_model = mj_model
    self.state_size = int(mj_model.nq + mj_model.nv)      # mjSTATE_PHYSICS = qpos, qvel (na == 0, no plugins)

  def capture(self, data) -> None:
    if not self.enabled:
      return
    self.buffer.append({"step": self.step_counter, "qpos": data.qpos.clone(), "qvel": data.qvel.clone()})
    self.step_counter += 1

  @contextmanager
  def watch(self, data) -> Iterator[None]:
    """`with nan_guard.watch(data): <one physics step>`: snapshot before, check after."""
    self.capture(data)
    yield
    self.check_and_dump(data)

  @staticmethod
  def detect_nans(data) -> torch.Tensor:
    """Per-world flag: NaN / Inf anywhere in qpos, qvel, qacc, qacc_warmstart (nan_guard.py:80-99)."""
    mask = torch.zeros(data.qpos.shape[0], dtype=torch.bool, device=data.qpos.device)
    for t in (data.qpos, data.qvel, data.qacc, data.qacc_warmstart):
      mask |= ~torch.isfinite(t).all(dim=-1)
    return mask

  def check_and_dump(self, data) -> bool:
    if not self.enabled or self._dumped:
      return False
    mask = self.detect_nans(data)
    if bool(mask.any()):
      self._dump_buffer(torch.where(mask)[0].cpu().numpy().tolist())
      self._dumped = True
      return True
    return False

  def _dump_buffer(self, nan_env_ids: list[int]) -> None:
    self.output_dir.mkdir(parents=True, exist_ok=True)
    stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
    filename = self.output_dir / f"nan_dump_{stamp}.npz"
    model_filename = self.output_dir / f"model_{stamp}.npz"
    envs = nan_env_ids[: self.max_envs_to_dump]
    out = {}
    for item in self.buffer:
      q, v = item["qpos"][envs].cpu().numpy(), item["qvel"][envs].cpu().numpy()
      out[f"states_step_{item['step']:06d}"] = np.concatenate([q, v], axis=1).astype(np.float64)
    out["_metadata"] = np.array({
      "num_envs_total": self.num_envs, "num_envs_dumped": len(envs), "nan_env_ids": nan_env_ids, "dumped_env_ids": list(envs),
      "state_size": self.state_size, "buffer_size": len(self.buffer), "detection_step": self.step_counter, "timestamp": stamp,
      "model_file": model_filename.name,
      "note": "rows are [qpos, qvel] per dumped world (MuJoCo's mjSTATE_PHYSICS layout for this model); the model file is a "
              "mjlab_b200.mjcf.ModelDesc npz",
    }, dtype=object)
    np.savez_compressed(filename, **out)
    self.mj_model.save(model_filename)
    for link, target in ((self.output_dir / "nan_dump_latest.npz", filename), (self.output_dir / "model_latest.npz", model_filename)):
      link.unlink(missing_ok=True)
      link.symlink_to(target.name)
    print(f"[NanGuard] NaN/Inf at step {self.step_counter} in envs {nan_env_ids[:10]}; dumped {len(self.buffer)} states of envs {envs} to {filename}")
