"""The oracle is 'parity unpinned' (no reference golden vectors exist), so pin it
with invariants the domain offers: conservation laws, closed forms, and the qualitative assertions
the reference's own tests make at the physics boundary (tests/test_sim.py, test_contact_sensor.py,
test_builtin_sensor.py, test_pd_actuator.py)."""
import numpy as np
import pytest

from mjlab_b200 import mjcf
from oracle import mjstep as O


def _strip_actuators(m):
  m.nu = 0
  for k, v in list(m.__dict__.items()):
    if k.startswith("actuator_") and isinstance(v, np.ndarray):
      m.__dict__[k] = v[:0]
  m.key_ctrl = m.key_ctrl[:, :0]
  return m


def _momenta(m, d):
  O.kinematics(m, d); O.com_pos(m, d); O.com_vel(m, d); O.subtree_vel(m, d)
  root = 2  # robot root body (0 world, 1 terrain)
  return d.subtree_linvel[root].copy(), d.subtree_angmom[root].copy()


def test_free_flight_momentum(g1):
  """Internal PD torques cannot change momentum: dv_com = g*T, angular momentum about the COM constant."""
  import copy
  drift = []
  for dt in (1e-3, 5e-4):
    m = copy.deepcopy(g1)
    m.opt_timestep = dt
    rng = np.random.default_rng(0)
    d = O.Data(m)
    d.qpos[:] = m.key_qpos[0]; d.qpos[2] = 10.0
    d.qvel[:] = rng.normal(size=m.nv)
    d.ctrl[:] = m.key_ctrl[0] + rng.normal(size=m.nu) * 0.3
    v0, L0 = _momenta(m, d)
    n = int(round(0.05 / dt))
    for _ in range(n):
      O.step(m, d)
    v1, L1 = _momenta(m, d)
    assert len(d.contacts) == 0
    np.testing.assert_allclose(v1 - v0, m.opt_gravity * n * dt, atol=2e-3)
    drift.append(np.linalg.norm(L1 - L0) / np.linalg.norm(L0))
  assert drift[0] < 2e-2
  assert drift[1] < 0.7 * drift[0]  # first-order integrator: error shrinks with dt


def test_energy_drift_first_order(g1):
  import copy
  errs = []
  for dt in (1e-3, 5e-4):
    m = _strip_actuators(copy.deepcopy(g1))
    m.opt_timestep, m.opt_integrator = dt, mjcf.INT_EULER
    m.dof_armature[:] = 0.01
    m.jnt_limited[:] = 0
    d = O.Data(m)
    d.qpos[:] = m.key_qpos[0]; d.qpos[2] = 10.0
    d.qvel[:] = np.random.default_rng(1).normal(size=m.nv) * 0.5

    def energy():
      O.kinematics(m, d); O.com_pos(m, d); O.crb(m, d)
      ke = 0.5 * d.qvel @ d.qM @ d.qvel
      return ke - np.sum(m.body_mass[:, None] * d.xipos * m.opt_gravity), ke
    e0, k0 = energy()
    for _ in range(int(round(0.05 / dt))):
      O.step(m, d)
    e1, _ = energy()
    errs.append(abs(e1 - e0) / k0)
  assert errs[0] < 0.1
  assert 0.4 < errs[1] / errs[0] < 0.6


BALL = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <geom name="floor" type="plane" size="0 0 .01"/>
  <body name="ball" pos="0 0 0.5"><freejoint/><geom name="ball" type="sphere" size="0.1" mass="1"/></body>
</worldbody>
<sensor><accelerometer name="acc" site="s"/></sensor>
</mujoco>
"""


def test_free_fall_closed_form():
  xml = BALL.replace('<sensor><accelerometer name="acc" site="s"/></sensor>', "")
  m = mjcf.Spec.from_string(xml).compile()
  d = O.Data(m)
  n = 100
  for _ in range(n):
    O.step(m, d)
  t = n * m.opt_timestep
  # semi-implicit Euler: v_k = -g t, z_n = z0 - g dt^2 n(n+1)/2
  assert abs(d.qvel[2] + 9.81 * t) < 1e-9
  assert abs(d.qpos[2] - (0.5 - 9.81 * m.opt_timestep**2 * n * (n + 1) / 2)) < 1e-9


def test_ball_rests_on_plane_with_weight_force():
  xml = BALL.replace('<sensor><accelerometer name="acc" site="s"/></sensor>', "")
  spec = mjcf.Spec.from_string(xml)
  spec.add_sensor(name="touch_found", type=mjcf.SENS_CONTACT, objtype=mjcf.OBJ_GEOM, objname="ball", intprm=(1, 3, 1))
  spec.add_sensor(name="touch_force", type=mjcf.SENS_CONTACT, objtype=mjcf.OBJ_GEOM, objname="ball", intprm=(2, 3, 1))
  m = spec.compile()
  d = O.Data(m)
  d.qpos[2] = 0.11
  for _ in range(1500):
    O.step(m, d)
  assert len(d.contacts) == 1 and d.contacts[0].geom == (0, 1)
  assert abs(d.qvel[2]) < 1e-3
  assert 0.09 < d.qpos[2] < 0.1  # small penetration of a soft contact
  assert d.sensordata[0] == 1.0
  # netforce, primary = ball (geom2 of the pair): force from ball onto floor = -weight
  np.testing.assert_allclose(d.sensordata[1:4], [0, 0, -9.81], atol=2e-2)
  # in the air: found == 0 and force == 0 (reference tests/test_contact_sensor.py:657-686)
  d.qpos[2] = 1.0
  O.forward(m, d)
  assert np.all(d.sensordata == 0)


def test_accelerometer_reads_gravity_at_rest(go1):
  d = O.Data(go1)
  d.qpos[:] = go1.key_qpos[0]
  d.ctrl[:] = go1.key_ctrl[0]
  for _ in range(400):
    O.step(go1, d)
  acc = d.sensordata[6:9]
  assert abs(np.linalg.norm(acc) - 9.81) < 0.3
  # all contacts together carry the weight; the feet carry most of it
  total = sum((c.frame.T @ O.contact_force(go1, d, c))[2] for c in d.contacts)
  assert abs(total - go1.body_subtreemass[2] * 9.81) < 2.0
  feet = -sum(d.sensordata[12 + 4 * k + 3] for k in range(4))
  assert 0.8 * total < feet <= total + 1e-9
  assert all(d.sensordata[12 + 4 * k] >= 1 for k in range(4))


def test_position_actuator_is_pd(go1):
  """IdealPD == builtin <position> (reference tests/test_pd_actuator.py:53-131): force = kp*(ctrl-q) - kd*qd, clamped."""
  d = O.Data(go1)
  rng = np.random.default_rng(3)
  d.qpos[:] = go1.key_qpos[0]
  d.qpos[7:] += rng.normal(size=12) * 0.2
  d.qvel[6:] = rng.normal(size=12)
  d.ctrl[:] = go1.key_ctrl[0]
  O.forward(go1, d)
  kp, kd = go1.actuator_gainprm[:, 0], -go1.actuator_biasprm[:, 2]
  q = d.qpos[go1.jnt_qposadr[go1.actuator_trnid[:, 0]]]
  v = d.qvel[go1.jnt_dofadr[go1.actuator_trnid[:, 0]]]
  want = np.clip(kp * (d.ctrl - q) - kd * v, go1.actuator_forcerange[:, 0], go1.actuator_forcerange[:, 1])
  np.testing.assert_allclose(d.actuator_force, want, rtol=1e-12)
  np.testing.assert_allclose(d.qfrc_actuator[go1.jnt_dofadr[go1.actuator_trnid[:, 0]]], want, rtol=1e-12)


def test_qM_matches_body_jacobian_assembly(g1):
  """CRB (oracle) vs sum_b J^T I J (model compiler's independent assembly) at qpos0."""
  d = O.Data(g1)
  O.kinematics(g1, d); O.com_pos(g1, d); O.crb(g1, d)
  np.testing.assert_allclose(d.qM, g1.qM0, atol=1e-12)


def test_solver_satisfies_kkt(g1):
  """At the solution the gradient M(a - a_smooth) - J^T f vanishes and forces are non-negative."""
  d = O.Data(g1)
  d.qpos[:] = g1.key_qpos[0]; d.qpos[2] -= 0.05
  d.ctrl[:] = g1.key_ctrl[0]
  import copy
  m = copy.deepcopy(g1)
  m.opt_iterations = 50
  O.forward(m, d)
  assert d.nefc > 0
  grad = d.qM @ d.qacc - d.qfrc_smooth - d.efc_J.T @ d.efc_force
  assert np.linalg.norm(grad) < 1e-6 * max(1.0, np.linalg.norm(d.qfrc_smooth))
  assert np.all(d.efc_force >= 0)
  jar = d.efc_J @ d.qacc - d.efc_aref
  np.testing.assert_allclose(d.efc_force, np.where(jar < 0, -d.efc_D * jar, 0), rtol=1e-9, atol=1e-9)


def test_joint_limit_rows(go1):
  d = O.Data(go1)
  d.qpos[:] = go1.key_qpos[0]; d.qpos[2] = 5.0
  j = 3  # FR_calf: range [-2.818, -0.888]
  d.qpos[go1.jnt_qposadr[j]] = -0.5
  O.forward(go1, d)
  assert d.nefc == 1 and d.efc_type[0] == 0
  assert d.efc_J[0, go1.jnt_dofadr[j]] == -1.0
  assert d.efc_pos[0] == pytest.approx(-0.888 + 0.5)
  assert d.efc_force[0] > 0
