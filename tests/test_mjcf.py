"""Model compiler: sizes, tree tables, collision filtering, actuator / keyframe assembly (reference:
tests/test_g1_constants.py, test_go1_constants.py, test_entity.py semantics at the compiled-model level)."""
import numpy as np

from mjlab_b200 import engine as E
from mjlab_b200 import mjcf


def test_g1_sizes_and_order(g1):
  assert (g1.nq, g1.nv, g1.nu, g1.nbody, g1.ngeom, g1.nsensordata) == (36, 35, 29, 32, 69, 21)
  assert g1.body_names[:3] == ["world", "terrain", "robot/pelvis"]
  assert g1.geom_names[0] == "terrain" and g1.geom_type[0] == mjcf.GEOM_PLANE
  # actuators are grouped by actuator cfg, not by joint order (entity.py:153-168)
  assert g1.actuator_names[0].endswith("left_shoulder_pitch_joint")
  assert np.all(g1.actuator_biasprm[:, 1] == -g1.actuator_gainprm[:, 0])
  assert np.all(g1.actuator_forcelimited == 1) and np.all(g1.actuator_ctrllimited == 0)
  # keyframe: KNEES_BENT (g1_constants.py:207-220)
  assert g1.key_qpos[0, 2] == 0.76
  jid = g1.jnt_names.index("robot/left_knee_joint")
  assert g1.key_qpos[0, g1.jnt_qposadr[jid]] == 0.669


def test_go1_collision_filter(go1):
  # contype=1, conaffinity=0 on the robot: only robot-vs-plane pairs survive (go1_constants.py:120-128)
  assert go1.npair == 30
  assert np.all(go1.pair_geom[:, 0] == 0)
  foot = go1.geom_names.index("robot/FR_foot_collision")
  assert go1.geom_condim[foot] == 3 and go1.geom_priority[foot] == 1 and go1.geom_friction[foot, 0] == 0.6
  assert np.allclose(go1.geom_solimp[foot, :3], (0.9, 0.95, 0.023))
  thigh = go1.geom_names.index("robot/FR_thigh_collision1")
  assert go1.geom_condim[thigh] == 1


def test_g1_self_collision_excludes(g1):
  pairs = {(g1.geom_names[a], g1.geom_names[b]) for a, b in g1.pair_geom}
  bodies = lambda a, b: (g1.body_names[g1.geom_bodyid[g1.geom_names.index(a)]], g1.body_names[g1.geom_bodyid[g1.geom_names.index(b)]])
  for a, b in pairs:
    ba, bb = bodies(a, b)
    assert ba != bb
    assert {ba, bb} != {"robot/pelvis", "robot/left_hip_roll_link"}  # <exclude>
  # parent-child pairs are filtered (hip_roll -> hip_yaw)
  assert ("robot/left_hip_collision", "robot/left_thigh_collision") not in pairs
  # foot capsules reach the plane pair list with the plane first
  assert ("terrain", "robot/left_foot1_collision") in pairs


def test_fromto_capsule_and_inferred_inertia():
  xml = """<mujoco><worldbody><body name="b" pos="0 0 1"><freejoint/>
    <geom name="c" type="capsule" size="0.1" fromto="0 0 0 1 0 0"/>
    <geom name="s" type="box" size="0.1 0.2 0.3" pos="0 0 0" mass="2"/></body></worldbody></mujoco>"""
  m = mjcf.Spec.from_string(xml).compile()
  g = m.geom_names.index("c")
  assert np.allclose(m.geom_pos[g], [0.5, 0, 0]) and np.isclose(m.geom_size[g, 1], 0.5)
  R = mjcf.quat_to_mat(m.geom_quat[g])
  assert np.allclose(np.abs(R[:, 2]), [1, 0, 0])
  vol = np.pi * 0.1**2 * 1.0 + 4 / 3 * np.pi * 0.1**3
  assert np.isclose(m.body_mass[1], 1000 * vol + 2.0)
  assert m.nq == 7 and m.nv == 6 and np.allclose(m.qpos0, [0, 0, 1, 1, 0, 0, 0])


def test_invweight_of_a_free_body():
  xml = """<mujoco><worldbody><body name="b"><freejoint/><geom type="sphere" size="0.1" mass="2"/></body></worldbody></mujoco>"""
  m = mjcf.Spec.from_string(xml).compile()
  I = 0.4 * 2 * 0.01
  assert np.allclose(m.body_invweight0[1], [0.5, 1 / I])
  assert np.allclose(m.dof_invweight0, [0.5] * 3 + [1 / I] * 3)
  assert np.isclose(m.stat_meaninertia, (3 * 2 + 3 * I) / 6)


def test_tree_tables(g1):
  t = E.derived_int_tables(g1)
  order, adr = t["level_body"], t["level_adr"]
  assert order[0] == 0 and adr[0] == 0 and adr[-1] == g1.nbody
  level = {int(b): lv for lv in range(t["nlevel"]) for b in order[adr[lv]:adr[lv + 1]]}
  for b in range(1, g1.nbody):
    assert level[b] == level[int(g1.body_parentid[b])] + 1
  # dof masks: ancestors' dofs including own
  mask = t["body_dofmask"].view(np.uint32)
  b = g1.body_names.index("robot/left_ankle_roll_link")
  full = int(mask[b, 0]) | (int(mask[b, 1]) << 32)
  assert bin(full).count("1") == 12  # 6 free + 6 leg joints
  assert E.model_scalars(g1, t)["maxchain"] == 23


def test_model_roundtrip(tmp_path, go1):
  p = tmp_path / "m.npz"
  go1.save(p)
  m2 = mjcf.ModelDesc.load(p)
  assert m2.body_names == go1.body_names and np.array_equal(m2.pair_geom, go1.pair_geom)
  assert m2.opt_timestep == go1.opt_timestep
