"""`CircularBuffer`: batched ring of the last `max_len` frames, newest first (src/mjlab/utils/buffers/circular_buffer.py).

Host-side bookkeeping in plain torch (a handful of index operations per control step, used by observation history and by
the learned actuator's input window; the per-physics-step delay path is the `b2_delay_step` kernel behind
`actuators.DelayBuffer`).  Semantics the reference's tests pin (tests/test_circular_buffer.py, docs
observation_history_delay.md:110-118): the first `append` after a (re)set back-fills every slot of that batch row, `buf[lag]`
clamps the lag to the number of pushes - 1, `buffer` returns the frames oldest -> newest.
"""

from __future__ import annotations

from typing import Sequence

import torch


class CircularBuffer:
  def __init__(self, max_len: int, batch_size: int, device: str = "cuda:0") -> None:
    if max_len < 1:
      raise ValueError(f"max_len must be >= 1, got {max_len}")
    self._max_len, self._batch_size, self._device = max_len, batch_size, torch.device(device)
    self._all = torch.arange(batch_size, device=self._device)
    self._num_pushes = torch.zeros(batch_size, dtype=torch.long, device=self._device)
    self._pointer = -1
    self._buffer: torch.Tensor | None = None   # [max_len, batch, ...]

  batch_size = property(lambda self: self._batch_size)
  device = property(lambda self: str(self._device))
  max_length = property(lambda self: self._max_len)
  is_initialized = property(lambda self: self._buffer is not None)

  @property
  def current_length(self) -> torch.Tensor:
    """Frames stored per batch row, <= max_len (circular_buffer.py:120-128)."""
    return torch.minimum(self._num_pushes, torch.full_like(self._num_pushes, self._max_len))

  @property
  def buffer(self) -> torch.Tensor:
    """[batch, max_len, ...], oldest -> newest (circular_buffer.py:130-147)."""
    if self._buffer is None:
      raise RuntimeError("Buffer not initialized. Call append() first.")
    return torch.roll(self._buffer, shifts=self._max_len - self._pointer - 1, dims=0).transpose(0, 1)

  def reset(self, batch_ids: Sequence[int] | torch.Tensor | slice | None = None) -> None:
    ids = slice(None) if batch_ids is None else batch_ids
    self._num_pushes[ids] = 0
    if self._buffer is not None:
      self._buffer[:, ids] = 0.0

  def append(self, data: torch.Tensor) -> None:
    if data.shape[0] != self._batch_size:
      raise ValueError(f"Expected batch size {self._batch_size}, got {data.shape[0]}")
    data = data.to(self._device)
    if self._buffer is None:
      self._pointer = -1
      self._buffer = torch.empty((self._max_len, *data.shape), dtype=data.dtype, device=self._device)
    self._pointer = (self._pointer + 1) % self._max_len
    self._buffer[self._pointer] = data
    first = self._num_pushes == 0   # newly (re)set rows: the whole history becomes the first frame
    self._buffer[:, first] = data[first]
    self._num_pushes += 1

  def __getitem__(self, key: torch.Tensor | int) -> torch.Tensor:
    if self._buffer is None:
      raise RuntimeError("Buffer not initialized. Call append() first.")
    if isinstance(key, int):
      key = torch.full((self._batch_size,), key, dtype=torch.long, device=self._device)
    else:
      key = (key.expand(self._batch_size) if key.ndim == 0 else key).to(device=self._device, dtype=torch.long)
    if key.numel() != self._batch_size:
      raise ValueError(f"Expected {self._batch_size} lags, got {key.numel()}")
    valid = torch.minimum(key, self._num_pushes.clamp_min(1) - 1).clamp_min(0)
    idx = torch.remainder(self._pointer - valid, self._max_len)
    return self._buffer[idx, self._all]
