"""NaN guard (src/mjlab/utils/nan_guard.py): detection mask, rolling buffer, dump contents.  CPU: the guard only touches
torch tensors, so plain CPU tensors stand in for `sim.data`."""
from types import SimpleNamespace as NS

import numpy as np
import torch

import mjlab_b200
from mjlab_b200.mjcf import ModelDesc
from mjlab_b200.nan_guard import NanGuard, NanGuardCfg


def _data(m, B):
  return NS(qpos=torch.zeros(B, m.nq), qvel=torch.zeros(B, m.nv), qacc=torch.zeros(B, m.nv), qacc_warmstart=torch.zeros(B, m.nv))


def test_disabled_guard_is_a_no_op(tmp_path):
  m = mjlab_b200.load_model("go1_flat")
  g = NanGuard(NanGuardCfg(enabled=False, output_dir=str(tmp_path)), 4, m)
  d = _data(m, 4)
  d.qpos[1, 0] = float("nan")
  with g.watch(d):
    pass
  assert not g.check_and_dump(d) and not list(tmp_path.iterdir())


def test_detects_and_dumps_the_history(tmp_path):
  m = mjlab_b200.load_model("go1_flat")
  B = 6
  g = NanGuard(NanGuardCfg(enabled=True, buffer_size=5, output_dir=str(tmp_path), max_envs_to_dump=2), B, m)
  d = _data(m, B)
  for step in range(8):                      # more steps than the buffer holds
    with g.watch(d):
      d.qpos += 1.0
      d.qvel += 0.5
      if step == 7:
        d.qvel[4, 3] = float("inf")
        d.qacc[1, 0] = float("nan")
        d.qacc_warmstart[5, 2] = float("nan")
  assert NanGuard.detect_nans(d).tolist() == [False, True, False, False, True, True]
  dump = np.load(tmp_path / "nan_dump_latest.npz", allow_pickle=True)
  meta = dump["_metadata"].item()
  assert meta["nan_env_ids"] == [1, 4, 5] and meta["dumped_env_ids"] == [1, 4] and meta["buffer_size"] == 5
  assert meta["state_size"] == m.nq + m.nv and meta["detection_step"] == 8
  keys = sorted(k for k in dump.files if k.startswith("states_step_"))
  assert keys == [f"states_step_{s:06d}" for s in range(3, 8)]          # the last 5 snapshots, taken BEFORE each step
  last = dump["states_step_000007"]
  assert last.shape == (2, m.nq + m.nv) and np.all(last[:, : m.nq] == 7.0) and np.all(last[:, m.nq:] == 3.5)
  again = ModelDesc.load(tmp_path / meta["model_file"])
  assert again.nq == m.nq and np.array_equal(again.qpos0, m.qpos0)
  # one dump per guard: a second detection does not write again
  n = len(list(tmp_path.iterdir()))
  with g.watch(d):
    pass
  assert len(list(tmp_path.iterdir())) == n
