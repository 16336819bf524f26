"""Analytic pins of the physics oracle (oracle/mjstep.py).  MuJoCo / mujoco_warp cannot be installed in this image or on the GPU
box (profiles/r02_reference_install_attempt.log), so the physics stays "parity unpinned" against the reference's binaries; what
CAN be pinned are closed forms that follow from MuJoCo's documented constraint model (Computation chapter: impedance d(r),
reference acceleration aref = -b v - k r, regulariser R = (1 - d) / d * diag(A), so that one isolated row obeys
a1 = (1 - d) a0 + d aref) and from classical mechanics - each computed here independently of the oracle's code:

  * the joint-limit response of an isolated hinge: acceleration at a violated limit, and the critically damped return
  * the friction slip threshold of a pyramidal contact at mu * N
  * the normal-mode periods of a double pendulum, omega^2 = (g / l)(2 -+ sqrt 2)
  * the resting force split of a capsule lying on a plane: lever rule over its two contacts
"""
import math

import numpy as np
import pytest

from mjlab_b200 import mjcf
from oracle import mjstep as O


def _compile(xml):
  return mjcf.Spec.from_string(xml).compile()


HINGE = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="0.001" gravity="0 0 0" integrator="Euler"/>
  <worldbody>
    <body name="arm" pos="0 0 1">
      <joint name="j" type="hinge" axis="0 1 0" range="-0.5 0.5" limited="true" solreflimit="{tc} {dr}" solimplimit="{d0} {d1} {w} 0.5 2"/>
      <inertial pos="0 0 0" mass="2.0" diaginertia="0.3 0.3 0.3"/>
      <geom type="sphere" size="0.05" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
</mujoco>
"""


def _impedance(r, d0, d1, w, mid=0.5, power=2):
  x = min(abs(r) / w, 1.0)
  y = x ** power / mid ** (power - 1) if x < mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
  return d0 + y * (d1 - d0)


@pytest.mark.parametrize("tc,dr,viol,vel", [(0.02, 1.0, 0.03, 0.4), (0.05, 1.0, 0.002, -0.2), (0.02, 0.7, 0.01, 0.0)])
def test_isolated_joint_limit_row_obeys_the_documented_law(tc, dr, viol, vel):
  """One hinge beyond its upper limit, nothing else: a1 = (1 - d) a0 + d aref with a0 = 0 (no gravity, no actuator),
  aref = -b v - k r in constraint space (r = -viol < 0, v = -qvel for the upper limit), k = d(r) / (dmax^2 tc^2 dr^2), b = 2 / (dmax tc)."""
  d0, d1, w = 0.9, 0.95, 0.001 if viol > 0.005 else 0.01
  m = _compile(HINGE.format(tc=tc, dr=dr, d0=d0, d1=d1, w=w))
  d = O.Data(m)
  d.qpos[0], d.qvel[0] = 0.5 + viol, vel
  O.forward(m, d)
  assert d.nefc == 1
  r, v = -viol, -vel
  imp = _impedance(r, d0, d1, w)
  k, b = 1.0 / (d1 * d1 * tc * tc * dr * dr), 2.0 / (d1 * tc)
  aref = -b * v - k * imp * r
  a1_constraint = imp * aref                      # (1 - d) * 0 + d * aref
  if a1_constraint < 0:                           # the row would pull INTO the limit: inactive, no force
    assert abs(d.qacc[0]) < 1e-12 and d.efc_force[0] == 0
    return
  np.testing.assert_allclose(d.qacc[0], -a1_constraint, rtol=1e-9)     # joint space: J = -1 for the upper limit
  np.testing.assert_allclose(d.efc_force[0], a1_constraint * 0.3, rtol=1e-9)   # f = I * a


def test_joint_limit_return_is_critically_damped_while_the_row_pushes():
  """Released from rest 0.02 rad beyond the limit (impedance width tiny -> d = dmax throughout): while the row is active the
  violation obeys r'' = d (-b r' - k d r) = -(2 / tc) r' - r / tc^2, i.e. the critically damped r(t) = r0 (1 + t / tc) exp(-t / tc).
  That solution asks for a PULLING force from t = tc on (r'' changes sign); a limit row is unilateral, so from there the joint
  coasts with the velocity it has, v = -r0 / (e tc)."""
  tc, d1 = 0.04, 0.95
  m = _compile(HINGE.format(tc=tc, dr=1.0, d0=d1, d1=d1, w=1e-6))
  m.opt_timestep = 2e-4
  d = O.Data(m)
  r0 = 0.02
  d.qpos[0] = 0.5 + r0
  ts, rs, fs = [], [], []
  for i in range(int(2 * tc / m.opt_timestep)):
    O.step(m, d)
    ts.append(d.time); rs.append(d.qpos[0] - 0.5); fs.append(d.efc_force[0] if d.nefc else 0.0)
  ts, rs, fs = np.array(ts), np.array(rs), np.array(fs)
  w = 1.0 / tc
  push = ts <= tc
  want = r0 * (1 + w * ts) * np.exp(-w * ts)
  assert np.abs(rs - want)[push].max() < 0.01 * r0                # first-order integrator at dt = tc / 200
  assert np.all(fs[ts < 0.95 * tc] > 0) and np.all(fs[ts > 1.05 * tc] == 0)     # the row lets go at t = tc
  r1, v1 = r0 * 2 / math.e, -r0 * w / math.e
  coast = ts > 1.05 * tc
  assert np.abs(rs - (r1 + v1 * (ts - tc)))[coast].max() < 0.01 * r0


SLIDER = """
<mujoco>
  <option timestep="0.002" integrator="Euler"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1" friction="{mu} 0.005 0.0001"/>
    <body name="box" pos="0 0 0.0495">
      <freejoint/>
      <geom name="b" type="box" size="0.3 0.3 0.05" mass="3.0" friction="{mu} 0.005 0.0001"/>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.mark.parametrize("mu", [0.4, 0.7])
def test_friction_slip_threshold_is_mu_times_normal_force(mu):
  """A flat box resting on the plane, pushed along a contact-frame axis in the contact plane (the applied torque cancels the lever
  arm of the force).  Below mu N it stays put (the soft constraint only creeps: millimetres per second); above, it slides with
  (F - mu N) / m.  Pyramidal cone: along a frame axis the pyramid's capacity equals the cone's, mu N.  Accelerations are measured
  over 0.16 s of sliding, not from one forward pass: the instant the push starts the soft normal rows still rebalance."""
  m = _compile(SLIDER.format(mu=mu))
  mass, g = 3.0, 9.81
  d = O.Data(m)
  for _ in range(300):                       # settle
    O.step(m, d)
  assert len(d.contacts) == 4
  N = sum(O.contact_force(m, d, c)[0] for c in d.contacts)
  np.testing.assert_allclose(N, mass * g, rtol=2e-3)
  q0, v0, z = d.qpos.copy(), d.qvel.copy(), d.qpos[2]
  Fmax = mu * mass * g

  def slide(f):
    d.qpos[:], d.qvel[:], d.time = q0, v0, 0.0
    d.qacc_warmstart[:] = 0
    d.xfrc_applied[:] = 0
    d.xfrc_applied[1, 0] = f * Fmax
    d.xfrc_applied[1, 4] = -f * Fmax * z
    vs = []
    for _ in range(100):
      O.step(m, d)
      vs.append(d.qvel[0])
    return (vs[99] - vs[19]) / (80 * m.opt_timestep), vs[99]

  for f in (0.5, 0.9):
    a, v = slide(f)
    assert abs(a) < 0.002 * f * Fmax / mass and abs(v) < 0.01      # sticks: no acceleration, creep below 1 cm/s
  for f in (1.3, 1.6):
    a, v = slide(f)
    np.testing.assert_allclose(a, (f - 1.0) * Fmax / mass, rtol=0.11)


PENDULUM = """
<mujoco>
  <option timestep="0.0005" integrator="Euler"/>
  <worldbody>
    <body name="l1" pos="0 0 2">
      <joint name="j1" type="hinge" axis="0 1 0"/>
      <inertial pos="0 0 -0.5" mass="1.0" diaginertia="1e-9 1e-9 1e-9"/>
      <body name="l2" pos="0 0 -0.5">
        <joint name="j2" type="hinge" axis="0 1 0"/>
        <inertial pos="0 0 -0.5" mass="1.0" diaginertia="1e-9 1e-9 1e-9"/>
      </body>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.mark.parametrize("mode", [0, 1])
def test_double_pendulum_normal_mode_periods(mode):
  """Two equal point masses on massless links of length l: omega^2 = (g / l)(2 -+ sqrt 2), mode shapes phi2 = +-sqrt 2 phi1 in
  ABSOLUTE angles.  Started in a mode shape at small amplitude the motion is a pure sinusoid of that frequency."""
  m = _compile(PENDULUM)
  l, g = 0.5, 9.81
  sgn = 1.0 if mode == 0 else -1.0
  omega = math.sqrt(g / l * (2 - sgn * math.sqrt(2)))
  d = O.Data(m)
  amp = 1e-3
  d.qpos[0] = amp
  d.qpos[1] = sgn * math.sqrt(2) * amp - amp           # relative joint angle = absolute(2) - absolute(1)
  ts, th = [], []
  nper = 3
  for _ in range(int(nper * 2 * math.pi / omega / m.opt_timestep)):
    O.step(m, d)
    ts.append(d.time); th.append(d.qpos[0])
  ts, th = np.array(ts), np.array(th)
  up = np.flatnonzero((th[:-1] < 0) & (th[1:] >= 0))   # upward zero crossings, linearly interpolated
  tz = ts[up] - th[up] * (ts[up + 1] - ts[up]) / (th[up + 1] - th[up])
  period = np.diff(tz).mean()
  np.testing.assert_allclose(period, 2 * math.pi / omega, rtol=2e-3)
  # a pure mode: the ratio of the two absolute angles stays at +-sqrt 2
  ratio = (d.qpos[0] + d.qpos[1]) / d.qpos[0] if abs(d.qpos[0]) > 0.2 * amp else sgn * math.sqrt(2)
  np.testing.assert_allclose(ratio, sgn * math.sqrt(2), rtol=0.05)


CAPSULE = """
<mujoco>
  <option timestep="0.002"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <body name="cap" pos="0 0 0.0495" euler="0 90 0">
      <freejoint/>
      <inertial pos="0 0 {off}" mass="2.0" diaginertia="0.02 0.02 0.002"/>
      <geom name="c" type="capsule" size="0.05 0.3"/>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.mark.parametrize("off", [0.0, 0.1, -0.2])
def test_resting_capsule_splits_its_weight_by_the_lever_rule(off):
  """A capsule lying on the plane touches it at its two end spheres (half length h = 0.3 apart from the centre).  With the centre
  of mass shifted by `off` along the axis, statics gives N1, N2 = m g (1/2 -+ off / (2 h)); both positive, summing to the weight."""
  m = _compile(CAPSULE.format(off=off))
  mass, g, h = 2.0, 9.81, 0.3
  d = O.Data(m)
  for _ in range(1500):
    O.step(m, d)
  assert len(d.contacts) == 2 and np.abs(d.qvel).max() < 1e-4
  ax = d.geom_xmat[1][:, 2]                            # capsule axis in the world
  forces = {}
  for c in d.contacts:
    s = float(np.dot(c.pos - d.geom_xpos[1], ax))      # +h or -h end
    forces[round(s / h)] = O.contact_force(m, d, c)[0]
  np.testing.assert_allclose(forces[1] + forces[-1], mass * g, rtol=1e-3)
  np.testing.assert_allclose(forces[1], mass * g * (0.5 + off / (2 * h)), rtol=5e-3)
  np.testing.assert_allclose(forces[-1], mass * g * (0.5 - off / (2 * h)), rtol=5e-3)
