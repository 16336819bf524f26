"""The env step's reuse of the forward pass (b2_substeps_after_forward, include/b200sim.h): after ManagerBasedRlEnv.step has
run `sim.forward()` for the reset envs (manager_based_rl_env.py:306-311), the next step's first sub-step finds kinematics, mass
matrix, contacts and bias forces of the SAME state in the data arrays, and only the ctrl-dependent tail is redone.  Here on the
host build of the kernel sources: the reuse is taken exactly where it may be, and every array is bit-identical to the plain path."""
import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import engine as E
from tests import util
from tests.hostemu.harness import HostSim

FIELDS = [a for _, a, _ in E.DATA_FLOAT_FIELDS if not a.startswith("memo_")] + [a for _, a, _ in E.DATA_INT_FIELDS if a != "smooth_memo"]


def _same(a: HostSim, b: HostSim, where: str) -> None:
  for f in FIELDS:
    x, y = getattr(a, f), getattr(b, f)
    assert x.tobytes() == y.tobytes(), (where, f, np.abs(x.astype(np.float64) - y).max())


@pytest.mark.parametrize("model", ["g1_flat", "go1_flat"])
def test_first_substep_reuses_forward_bit_identically(model):
  m = mjlab_b200.load_model(model)
  nw, dec = 6, 4
  qpos, qvel, ctrl, _ = util.random_states(m, nw, seed=3)
  sims = [HostSim(m, nw, 48, 300), HostSim(m, nw, 48, 300)]
  for s in sims:
    s.qpos[:], s.qvel[:], s.ctrl[:] = qpos, qvel, ctrl
    s.forward()
  plain, memo = sims
  rng = np.random.default_rng(0)
  expect_hits = 0
  for it in range(4):
    c = ctrl + rng.normal(size=ctrl.shape) * 0.3          # the new action: ctrl differs from the forward pass's
    touched = set()
    if it == 1:                                           # an interval event pushes world 2 AFTER the forward pass
      for s in sims:
        s.qvel[2, :3] += np.float32(0.25)
      touched.add(2)
    if it == 2:                                           # an external wrench on world 4 needs the kinematic quantities
      for s in sims:
        s.xfrc_applied[4, 3, :3] = (1.0, -2.0, 0.5)
      touched.add(4)
    if it == 3:                                           # model edited without telling: the owner must clear the flag
      for s in sims:
        s.xfrc_applied[:] = 0
      memo.smooth_memo[1] = 0
      touched.add(1)
    h0 = memo.memo_hits()
    for k in range(dec):
      for s in sims:
        s.ctrl[:] = c
      plain.step()
      memo.step(after_forward=(k == 0))
      _same(plain, memo, f"step {it}.{k}")
    expect_hits = nw - len(touched)
    assert memo.memo_hits() - h0 == expect_hits, (it, memo.memo_hits() - h0, expect_hits)   # later sub-steps never hit: the state moved on
    # resets of some worlds, then the forward pass over all worlds (reference forward mode)
    for s in sims:
      s.qpos[it % nw], s.qvel[it % nw] = qpos[it % nw], 0
      s.forward()
    _same(plain, memo, f"forward {it}")
  assert expect_hits > 0


def test_no_forward_no_reuse():
  """Without a pass on the current state (masked forward mode: only reset worlds are refreshed) nothing is reused."""
  m = mjlab_b200.load_model("go1_flat")
  qpos, qvel, ctrl, _ = util.random_states(m, 3, seed=1)
  a, b = HostSim(m, 3, 48, 300), HostSim(m, 3, 48, 300)
  for s in (a, b):
    s.qpos[:], s.qvel[:], s.ctrl[:] = qpos, qvel, ctrl
  h0 = b.memo_hits()
  for k in range(3):
    a.step()
    b.step(after_forward=True)
    _same(a, b, f"step {k}")
  assert b.memo_hits() == h0
