"""The compiled CPU build used for bench.py's cpu_baseline (oracle/cport) reproduces the numpy oracle, so the baseline
times the same arithmetic the parity tests check."""
import numpy as np

from oracle import mjstep as O
from oracle.cport import CpuSim
from tests.util import random_states


def test_cport_matches_numpy_oracle(go1):
  m = go1
  m.opt_timestep, m.opt_iterations, m.opt_ls_iterations = 0.005, 10, 20
  nworld, nstep = 6, 3
  qpos, qvel, ctrl, xfrc = random_states(m, nworld, seed=4)
  sim = CpuSim(m, nworld, 35, 300)
  sim.qpos[:], sim.qvel[:], sim.ctrl[:] = qpos, qvel, ctrl
  sim.step(nstep, 3)   # three threads over six worlds
  for w in range(nworld):
    d = O.Data(m)
    d.qpos[:], d.qvel[:], d.ctrl[:] = qpos[w], qvel[w], ctrl[w]
    for _ in range(nstep):
      O.step(m, d, 35, 300)
    assert int(sim.ncon[w]) == len(d.contacts)
    np.testing.assert_allclose(sim.qpos[w], d.qpos, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sim.qvel[w], d.qvel, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(sim.sensordata[w], d.sensordata, rtol=1e-7, atol=1e-9)
