"""Friction-loss constraint rows (dof_frictionloss > 0; the `frictionloss` of the reference's actuator cfgs,
src/mjlab/actuator/actuator.py:28, written to the joint by utils/spec.py:117, and the `dof_frictionloss` randomisation field,
envs/mdp/events.py:267).  MuJoCo's rule (mj_instantiateFriction / mj_constraintUpdate / PrimalEval): one row per dof with
frictionloss > 0, ahead of the limit and contact rows, cost quadratic for |J a - aref| < R f and linear beyond with the force
saturated at -/+ f.

1. analytic pins of the oracle: a hinge under a torque below the bound does not move, above the bound it accelerates with
   (tau - f) / I, a spinning free hinge decelerates with f / I and then stops for good;
2. the kernel source (float64 host build) against the oracle on Go1 with friction loss on every actuated dof, Newton and CG:
   row order, row counts, forces and accelerations;
3. the model compiler carries `frictionloss`, `solreffriction`, `solimpfriction` into dof_frictionloss / dof_solref / dof_solimp.
"""
import copy

import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import engine as E
from mjlab_b200 import mjcf
from oracle import mjstep as O
from tests import util
from tests.hostemu.harness import HostSim

HINGE = """
<mujoco>
  <option timestep="0.002" gravity="0 0 0"/>
  <worldbody>
    <body name="arm" pos="0 0 1">
      <joint name="j" type="hinge" axis="0 1 0" frictionloss="0.5" solreffriction="0.01 1" solimpfriction="0.95 0.99 0.001 0.5 2"/>
      <geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.03" mass="1.5" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
  <actuator><motor name="m" joint="j"/></actuator>
</mujoco>
"""


def _hinge():
  return mjcf.Spec.from_string(HINGE).compile()


def test_compiler_carries_friction_fields():
  m = _hinge()
  assert m.nv == 1 and m.dof_frictionloss[0] == 0.5
  np.testing.assert_allclose(m.dof_solref[0], [0.01, 1.0])
  np.testing.assert_allclose(m.dof_solimp[0], [0.95, 0.99, 0.001, 0.5, 2.0])
  assert E.count_friction_dofs(m) == 1
  E.check_supported(m)
  g = mjlab_b200.load_model("go1_flat")      # model files written before the fields existed: MuJoCo's defaults
  assert g.dof_solref.shape == (g.nv, 2) and g.dof_solimp.shape == (g.nv, 5) and E.count_friction_dofs(g) == 0
  g.dof_frictionloss = -np.ones(g.nv)
  with pytest.raises(ValueError, match="frictionloss"):
    E.check_supported(g)


def test_oracle_stick_slip_analytic():
  m = _hinge()
  d = O.Data(m)
  O.forward(m, d)
  inertia = float(d.qM[0, 0])
  f = 0.5
  # below the bound: the row balances the torque.  The constraint is soft (MuJoCo's rows are springs-dampers in acceleration
  # space): in the steady state qacc = 0 the row force D (aref - 0) = -D b v equals -tau, i.e. the joint creeps with
  # v = tau R / b, R = invweight (1 - imp) / imp with imp = dmin at pos = 0, b = 2 / (dmax timeconst)
  d = O.Data(m)
  d.ctrl[:] = 0.3
  for _ in range(400):
    O.step(m, d)
  assert d.nf == 1 and d.nefc == 1
  R = (1.0 / inertia) * (1 - 0.95) / 0.95
  b = 2.0 / (0.99 * 0.01)
  np.testing.assert_allclose(d.efc_R[0], R, rtol=1e-3)    # dof_invweight0 of a single hinge = 1 / I
  np.testing.assert_allclose(d.qvel[0], 0.3 * d.efc_R[0] / b, rtol=1e-6)
  np.testing.assert_allclose(d.efc_force[0], -0.3, rtol=1e-6)
  assert abs(d.qvel[0]) < 2e-3        # against 0.3 / I * t = 6 rad/s after 0.8 s without the row
  # above the bound: the force saturates at f and the joint accelerates with (tau - f) / I
  d = O.Data(m)
  d.ctrl[:] = 0.9
  O.forward(m, d)
  np.testing.assert_allclose(d.efc_force[0], -f, rtol=1e-12)
  np.testing.assert_allclose(d.qacc[0], (0.9 - f) / inertia, rtol=1e-9)
  # spinning freely: decelerates with f / I, then sticks (velocity does not change sign and run away)
  d = O.Data(m)
  d.qvel[:] = 1.0
  O.forward(m, d)
  np.testing.assert_allclose(d.qacc[0], -f / inertia, rtol=1e-9)
  t_stop = 1.0 * inertia / f
  n = int(t_stop / m.opt_timestep) + 50
  for _ in range(n):
    O.step(m, d)
  assert abs(d.qvel[0]) < 1e-3
  q_stop = d.qpos[0]
  np.testing.assert_allclose(q_stop, 0.5 * 1.0 * t_stop, rtol=2e-2)   # v0^2 / (2 a) = v0 t_stop / 2
  for _ in range(100):
    O.step(m, d)
  assert abs(d.qpos[0] - q_stop) < 1e-3


def _go1_with_friction(solver):
  m = copy.deepcopy(mjlab_b200.load_model("go1_flat"))
  fl = np.zeros(m.nv)
  fl[6:] = np.linspace(0.05, 0.6, m.nv - 6)    # actuated dofs; joints at rest, sliding and in between all occur below
  fl[8] = 0.0                                   # a dof without a row in the middle: row -> dof map is not the identity
  m.dof_frictionloss = fl
  m.opt_solver, m.opt_iterations, m.opt_ls_iterations = solver, 100, 50
  return m


@pytest.mark.parametrize("solver", [2, 1])
def test_kernel_source_matches_oracle_with_friction_rows(solver):
  m = _go1_with_friction(solver)
  nfr = E.count_friction_dofs(m)
  assert nfr == m.nv - 7
  nw = 6
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=11)
  qvel[:, 6:] *= np.array([0.0, 0.02, 1.0, 5.0, 0.2, 1.0])[:, None]   # world 0: joints at rest (rows quadratic), world 3: all sliding
  hs = HostSim(m, nw, 35, 300, real="double")
  hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
  hs.forward()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 35, 300, integrate_last=False)
  zones = set()
  for w, d in enumerate(ds):
    assert d.nf == nfr and hs.nf[w] == nfr and hs.nl[w] == np.sum(d.efc_type == 0)
    assert hs.nefc[w] == d.nefc and hs.solver_niter[w] == d.solver_niter, (w, hs.solver_niter[w], d.solver_niter)
    np.testing.assert_allclose(hs.efc_D[w, : d.nefc], d.efc_D, rtol=1e-12)
    np.testing.assert_allclose(hs.efc_aref[w, : d.nefc], d.efc_aref, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(hs.qacc[w], d.qacc, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(hs.efc_force[w, : d.nefc], d.efc_force, rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(hs.qfrc_constraint[w], d.qfrc_constraint, rtol=1e-7, atol=1e-7)
    ff = d.efc_force[:nfr]
    bound = d.efc_frictionloss[:nfr]
    assert np.all(np.abs(ff) <= bound * (1 + 1e-12))
    zones |= {"sat"} if np.any(np.abs(ff) == bound) else set()
    zones |= {"quad"} if np.any(np.abs(ff) < bound) else set()
  assert zones == {"sat", "quad"}


def test_friction_rows_trajectory_kernel_source_vs_oracle():
  """20 steps (implicitfast) with friction rows, limits and contacts together; float64 host build against the oracle."""
  m = _go1_with_friction(2)
  nw = 3
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=3)
  hs = HostSim(m, nw, 35, 300, real="double")
  hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
  for _ in range(20):
    hs.step()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 20, 35, 300)
  for w, d in enumerate(ds):
    np.testing.assert_allclose(hs.qpos[w], d.qpos, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(hs.qvel[w], d.qvel, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("njmax,arena", [(8, 0), (14, 0), (300, 600)])
def test_friction_rows_under_row_and_arena_overflow(njmax, arena):
  """Friction-loss rows come first, so `njmax` truncation reaches them last: njmax below the friction-row count cuts friction rows,
  a little above it drops limits / contacts; a deliberately small solver arena (host build knob) drops contact rows from the tail.
  Same truncation in the oracle and the kernel source (njmax cases); the arena case must still agree on the rows it keeps."""
  m = _go1_with_friction(2)
  nfr = E.count_friction_dofs(m)
  nw = 3
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=8, drop=0.06)
  hs = HostSim(m, nw, 35, njmax, real="double")
  if arena:
    hs.lib.b2h_set_jpool_budget(arena)
  try:
    hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
    hs.forward()
  finally:
    hs.lib.b2h_set_jpool_budget(0)
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 35, njmax, integrate_last=False)
  for w, d in enumerate(ds):
    assert hs.nf[w] == min(nfr, njmax)
    if not arena:
      assert hs.nf[w] == d.nf and hs.nefc[w] == d.nefc <= njmax
      np.testing.assert_allclose(hs.qacc[w], d.qacc, rtol=1e-8, atol=1e-8)
      np.testing.assert_allclose(hs.efc_force[w, : d.nefc], d.efc_force, rtol=1e-7, atol=1e-7)
    else:
      assert nfr <= hs.nefc[w] <= d.nefc and np.all(np.isfinite(hs.qacc[w]))
  if arena:
    assert hs.nefc_dropped.sum() > 0
