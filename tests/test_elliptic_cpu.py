"""Elliptic friction cones (`MujocoCfg.cone = "elliptic"`, /root/reference/src/mjlab/sim/sim.py:51-56 -> mjCONE_ELLIPTIC).

1. the oracle's cone terms are self-consistent: force = -d cost / d residual and the cone Hessian block = d2 cost / d residual2
   (central differences), in all three zones;
2. the oracle is pinned analytically: a box pushed in ANY direction of the contact plane sticks below mu N and slides above it
   with a friction force of exactly mu N (the pyramid's capacity depends on the direction: mu N / sqrt 2 on the diagonal);
3. the kernel source (float64 host build) reproduces the oracle digit for digit - rows, D, aref, iteration counts, forces,
   accelerations - for Newton and CG, on Go1 / G1 states with many contacts, one step and a trajectory."""
import copy
import math

import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import mjcf
from oracle import mjstep as O
from tests import util
from tests.hostemu.harness import HostSim


def _cone_cost(x, D, mu, fri):
  cost, _, _ = O._row_update(x.copy(), D, 1.0 / D, np.zeros(0), 0, [(0, len(x), mu, fri)])
  return float(np.sum(cost))


@pytest.mark.parametrize("x,zone", [((0.5, 0.1, -0.1), 0), ((-1.0, 0.05, 0.02), 1), ((-0.1, 0.4, -0.3), 2), ((0.2, 0.5, 0.1), 2)])
def test_cone_force_and_hessian_are_derivatives_of_the_cost(x, zone):
  mu, fri = 0.6, np.array([0.8, 0.8])
  D = np.array([200.0, 50.0, 50.0])
  x = np.array(x, float)
  U = x * np.concatenate(([mu], fri))
  assert O._cone_zone(U[0], float(np.linalg.norm(U[1:])), mu) == zone
  hc = []
  cost, force, act = O._row_update(x.copy(), D, 1.0 / D, np.zeros(0), 0, [(0, 3, mu, fri)], hc)
  h = 1e-6
  g = np.array([(_cone_cost(x + h * e, D, mu, fri) - _cone_cost(x - h * e, D, mu, fri)) / (2 * h) for e in np.eye(3)])
  np.testing.assert_allclose(force, -g, rtol=1e-6, atol=1e-7)
  H = np.zeros((3, 3))
  for i, e in enumerate(np.eye(3)):
    fp = O._row_update(x + h * e, D, 1.0 / D, np.zeros(0), 0, [(0, 3, mu, fri)])[1]
    fm = O._row_update(x - h * e, D, 1.0 / D, np.zeros(0), 0, [(0, 3, mu, fri)])[1]
    H[i] = -(fp - fm) / (2 * h)
  want = {0: np.zeros((3, 3)), 1: np.diag(D)}.get(zone)
  if zone == 2:
    assert len(hc) == 1 and not act.any()
    want = hc[0][2]
    assert np.all(np.linalg.eigvalsh(want) > -1e-9 * np.abs(want).max())     # convex
  else:
    assert not hc and act.all() == (zone == 1)
  np.testing.assert_allclose(H, want, rtol=1e-5, atol=1e-6 * D.max())
  # the line-search evaluator agrees with the update along a ray (cost, slope, curvature)
  v = np.array([0.3, -0.2, 0.5])
  qg = (0.0, 0.0, 0.0)

  def ev(a):
    return O._ls_eval(a, qg, x.copy(), v, D, 1.0 / D, np.zeros(0), 0, [(0, 3, mu, fri)])

  for a in (0.0, 0.05, 0.3):
    p = ev(a)
    assert abs(p.cost - _cone_cost(x + a * v, D, mu, fri)) < 1e-9 * max(1.0, p.cost)
    if O._cone_zone(*((lambda Ua: (Ua[0], float(np.linalg.norm(Ua[1:]))))((x + a * v) * np.concatenate(([mu], fri)))), mu) \
       == O._cone_zone(*((lambda Ua: (Ua[0], float(np.linalg.norm(Ua[1:]))))((x + (a + 1e-5) * v) * np.concatenate(([mu], fri)))), mu):
      d1 = (ev(a + 1e-6).cost - ev(a - 1e-6).cost) / 2e-6 if a > 0 else (ev(1e-6).cost - p.cost) / 1e-6
      assert abs(p.d1 - d1) < 1e-4 * max(1.0, abs(p.d1))


SLIDER = """
<mujoco>
  <option timestep="0.002" integrator="Euler" cone="{cone}"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1" friction="{mu} 0.005 0.0001"/>
    <body name="box" pos="0 0 0.0495">
      <freejoint/>
      <geom name="b" type="box" size="0.3 0.3 0.05" mass="3.0" friction="{mu} 0.005 0.0001"/>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.mark.parametrize("ang", [0.0, math.pi / 4, 2.0])
def test_elliptic_cone_is_isotropic_where_the_pyramid_is_not(ang):
  mu, mass, g = 0.5, 3.0, 9.81
  Fmax = mu * mass * g

  def run(cone, f):
    m = mjcf.Spec.from_string(SLIDER.format(mu=mu, cone=cone)).compile()
    d = O.Data(m)
    for _ in range(300):
      O.step(m, d)
    d.qacc_warmstart[:] = 0
    d.xfrc_applied[1, :3] = (f * Fmax * math.cos(ang), f * Fmax * math.sin(ang), 0.0)   # at the COM, 5 cm above the plane
    vs, ratios = [], []
    for _ in range(100):
      O.step(m, d)
      vs.append(math.hypot(d.qvel[0], d.qvel[1]))
      ft = sum((O.contact_force(m, d, c) for c in d.contacts), np.zeros(3))
      if ft[0] > 0:
        ratios.append(math.hypot(ft[1], ft[2]) / ft[0])
    return (vs[99] - vs[19]) / (80 * m.opt_timestep), vs[99], float(np.median(ratios[20:]))

  a, v, _ = run("elliptic", 0.9)
  assert abs(a) < 0.002 * Fmax / mass and v < 0.01                      # sticks in every direction up to mu N
  a, v, ratio = run("elliptic", 1.5)
  np.testing.assert_allclose(ratio, mu, rtol=1e-3)                      # sliding: |sum f_t| = mu sum f_n in any direction (four contacts)
  np.testing.assert_allclose(a, 0.5 * Fmax / mass, rtol=0.11)
  if abs(ang - math.pi / 4) < 1e-9:                                     # the pyramid gives way at mu N / sqrt 2 on its diagonal
    a, v, ratio = run("pyramidal", 0.9)
    assert a > 0.1 * Fmax / mass
    np.testing.assert_allclose(ratio, mu / math.sqrt(2), rtol=1e-3)


@pytest.mark.parametrize("name,solver,fl", [("go1_flat", 2, 0.0), ("g1_flat", 2, 0.0), ("go1_flat", 1, 0.0), ("go1_flat", 2, 0.3), ("go1_flat", 1, 0.3)])
def test_elliptic_kernel_source_matches_oracle(name, solver, fl):
  m = copy.deepcopy(mjlab_b200.load_model(name))
  m.opt_cone, m.opt_solver, m.opt_impratio = 1, solver, 2.0
  if fl > 0:                                      # friction-loss rows next to the cone contacts (k_solve_t<*, true, true>)
    m.dof_frictionloss = np.concatenate((np.zeros(6), np.linspace(0.05, fl, m.nv - 6)))
  m.opt_iterations, m.opt_ls_iterations = (100, 50) if solver == 1 else (20, 50)
  nw = 6
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=11)
  hs = HostSim(m, nw, 35, 300, real="double")
  hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
  hs.forward()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 35, 300, integrate_last=False)
  nmid = 0
  for w, d in enumerate(ds):
    assert hs.nefc[w] == d.nefc, (w, hs.nefc[w], d.nefc)
    n = d.nefc
    if n:
      np.testing.assert_allclose(hs.efc_J[w, :n], d.efc_J, rtol=1e-9, atol=1e-11)
      np.testing.assert_allclose(hs.efc_D[w, :n], d.efc_D, rtol=1e-10)
      np.testing.assert_allclose(hs.efc_aref[w, :n], d.efc_aref, rtol=1e-8, atol=1e-9)
    assert hs.solver_niter[w] == d.solver_niter, (w, hs.solver_niter[w], d.solver_niter)
    np.testing.assert_allclose(hs.qacc[w], d.qacc, rtol=1e-8, atol=1e-8)
    if n:
      np.testing.assert_allclose(hs.efc_force[w, :n], d.efc_force, rtol=1e-7, atol=1e-8)
    for a, dim, mu, fri in d.efc_cones:                                # every contact force lies in its cone
      f = d.efc_force[a:a + dim]
      assert f[0] >= -1e-9 and np.linalg.norm(f[1:] / fri) <= f[0] * (1 + 1e-6) + 1e-9
      nmid += f[0] > 0 and np.linalg.norm(f[1:] / fri) > 0.999 * f[0]
    np.testing.assert_allclose(hs.sensordata[w], d.sensordata, rtol=1e-7, atol=1e-8)
  assert sum(len(d.efc_cones) for d in ds) >= 10 and nmid >= 1        # sliding contacts (cone surface) are in the sample


def test_elliptic_trajectory_kernel_source_matches_oracle():
  m = copy.deepcopy(mjlab_b200.load_model("go1_flat"))
  m.opt_cone = 1
  nw, nstep = 3, 12
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=3)
  hs = HostSim(m, nw, 35, 300, real="double")
  hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
  for _ in range(nstep):
    hs.step()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, nstep, 35, 300)
  for w, d in enumerate(ds):
    np.testing.assert_allclose(hs.qpos[w], d.qpos, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(hs.qvel[w], d.qvel, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("njmax,arena", [(10, 0), (17, 0), (300, 500)])
def test_elliptic_rows_under_row_and_arena_overflow(njmax, arena):
  """`njmax` truncation keeps whole contacts (three rows each) exactly as the oracle does; a deliberately small solver arena (host
  build knob) drops contacts from the tail: the kept rows form a prefix of whole contacts and the solve stays finite."""
  m = copy.deepcopy(mjlab_b200.load_model("go1_flat"))
  m.opt_cone = 1
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
    if not arena:
      assert hs.nefc[w] == d.nefc <= njmax
      np.testing.assert_array_equal(hs.contact_efc_address[w, : len(d.contacts)], [c.efc_address for c in d.contacts])
      np.testing.assert_allclose(hs.qacc[w], d.qacc, rtol=1e-8, atol=1e-8)
      np.testing.assert_allclose(hs.efc_force[w, : d.nefc], d.efc_force, rtol=1e-7, atol=1e-7)
    else:
      assert hs.nefc[w] <= d.nefc and np.all(np.isfinite(hs.qacc[w]))
      kept = hs.contact_efc_address[w, : len(d.contacts)]
      assert (hs.nefc[w] - hs.nl[w]) == 3 * int(np.sum(kept >= 0))       # whole contacts only
  assert hs.nefc_dropped.sum() > 0
