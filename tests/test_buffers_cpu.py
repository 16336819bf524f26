"""`mjlab_b200.buffers.CircularBuffer` against the reference's own class (golden: tools/make_actuator_golden.py runs
src/mjlab/utils/buffers/circular_buffer.py): back-fill on the first push after a (re)set, lag clamping, chronological
`buffer` view, partial reset.  Plus the error behaviour the reference's tests pin (tests/test_circular_buffer.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from mjlab_b200.buffers import CircularBuffer

G = np.load(Path(__file__).resolve().parent / "golden" / "actuators.npz")


def test_circular_buffer_matches_reference_class():
  x = torch.as_tensor(G["delay_in"])
  B = x.shape[1]
  cb = CircularBuffer(max_len=4, batch_size=B, device="cpu")
  for t in range(8):
    if t == 5:
      cb.reset([0, 3])
    cb.append(x[t])
    for lag in range(6):
      np.testing.assert_array_equal(cb[torch.full((B,), lag, dtype=torch.long)].numpy(), G["circ_out"][t, lag])
      np.testing.assert_array_equal(cb[lag].numpy(), G["circ_out"][t, lag])
    np.testing.assert_array_equal(cb.buffer.numpy(), G["circ_buffer"][t])
    np.testing.assert_array_equal(cb.current_length.numpy(), G["circ_len"][t])


def test_circular_buffer_errors():
  with pytest.raises(ValueError):
    CircularBuffer(max_len=0, batch_size=2, device="cpu")
  cb = CircularBuffer(max_len=2, batch_size=2, device="cpu")
  assert not cb.is_initialized
  with pytest.raises(RuntimeError):
    cb[0]
  with pytest.raises(RuntimeError):
    cb.buffer
  with pytest.raises(ValueError):
    cb.append(torch.zeros(3, 1))
  cb.append(torch.ones(2, 1))
  with pytest.raises(ValueError):
    cb[torch.zeros(3, dtype=torch.long)]
  assert cb.is_initialized and cb.max_length == 2 and cb.batch_size == 2
