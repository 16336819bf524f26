"""This repo's model compiler (mjlab_b200/mjcf.py + robots.py: MJCF -> arrays without MuJoCo) held to the reference's robot constants:
a mirror of the reference's `tests/test_g1_constants.py:24-141` and `tests/test_go1_constants.py:20-95`, which assert these facts
about the MjModel MuJoCo compiles.  The expected numbers are dumped from the reference's own `g1_constants.py` / `go1_constants.py`
(tools/make_robot_constants_golden.py -> tests/golden/robot_constants.json)."""
import json
import re
from pathlib import Path

import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import envcfg

G = json.loads((Path(__file__).resolve().parent / "golden" / "robot_constants.json").read_text())
FOOT = {"g1": r"^(left|right)_foot[1-7]_collision$", "go1": r"^[FR][LR]_foot_collision$"}


def _local(names):
  return [n.split("/")[-1] for n in names]


def _first(table: dict, name: str, default=None):
  """First matching pattern wins, `re.match` from the start of the name (NOT fullmatch: `.*_collision` also selects
  `FR_thigh_collision1`), as the reference resolves its {pattern: value} dicts (utils/string.py:5-21)."""
  for pat, v in table.items():
    if re.match(pat, name):
      return v
  return default


@pytest.fixture(scope="module", params=["g1", "go1"])
def robot(request):
  return request.param, mjlab_b200.load_model(f"{request.param}_flat"), G[request.param]


def test_actuator_parameters(robot):
  """test_g1_constants.py:37-49 / test_go1_constants.py:33-45: gainprm[0] = kp, biasprm = (0, -kp, -kd), forcerange = +-effort_limit;
  plus the joint armature every actuator cfg carries."""
  name, m, ref = robot
  act_names = _local(m.actuator_names)
  seen = 0
  for cfg in ref["actuators"]:
    for i, an in enumerate(act_names):
      if not any(re.match(p, an) for p in cfg["joint_names_expr"]):
        continue
      seen += 1
      assert m.actuator_gainprm[i, 0] == pytest.approx(cfg["stiffness"], rel=1e-12)
      assert m.actuator_biasprm[i, 1] == pytest.approx(-cfg["stiffness"], rel=1e-12)
      assert m.actuator_biasprm[i, 2] == pytest.approx(-cfg["damping"], rel=1e-12)
      assert tuple(m.actuator_forcerange[i]) == (-cfg["effort_limit"], cfg["effort_limit"])
      dof = int(m.jnt_dofadr[int(m.actuator_trnid[i, 0])])
      assert m.dof_armature[dof] == pytest.approx(cfg["armature"], rel=1e-12)
      assert m.dof_frictionloss[dof] == cfg["frictionloss"] == 0.0
  assert seen == m.nu == ref["num_actuators"]


def test_actuators_are_force_limited_not_ctrl_limited(robot):
  """test_g1_constants.py:128-141."""
  _, m, _ = robot
  assert not np.any(np.asarray(m.actuator_ctrllimited)) and np.all(np.asarray(m.actuator_forcelimited) == 1)


def test_keyframe(robot):
  """test_g1_constants.py:52-78 / test_go1_constants.py:48-66: base pose and joint positions of the init keyframe."""
  name, m, ref = robot
  kf = ref["keyframe"]
  np.testing.assert_array_equal(m.key_qpos[0, :3], kf["pos"])
  np.testing.assert_array_equal(m.key_qpos[0, 3:7], kf["rot"])
  jn = _local(m.jnt_names)
  n = 0
  for j, jname in enumerate(jn):
    if int(m.jnt_type[j]) == 0:    # the free joint
      continue
    want = _first(kf["joint_pos"], jname, 0.0)
    np.testing.assert_allclose(m.key_qpos[0, int(m.jnt_qposadr[j])], want, rtol=1e-5, atol=1e-12, err_msg=jname)
    n += 1
  assert n == ref["num_joints"]


def test_collision_geoms(robot):
  """test_g1_constants.py:81-118 / test_go1_constants.py:69-87: foot geoms condim 3, priority 1, friction 0.6 (Go1: its solimp too);
  every other `_collision` geom condim 1; 14 / 4 foot geoms; everything that is not a `_collision` geom does not collide
  (CollisionCfg.disable_other_geoms)."""
  name, m, ref = robot
  col = ref["collision"]
  gn = _local(m.geom_names)
  robot_geoms = [i for i, n in enumerate(m.geom_names) if n.startswith("robot/")]
  feet = 0
  for i in robot_geoms:
    g = gn[i]
    if any(re.match(p, g) for p in col["geom_names_expr"]):    # utils/string.py:24-29 filter_exp
      assert int(m.geom_condim[i]) == _first(col["condim"], g), g
      assert int(m.geom_priority[i]) == _first(col["priority"], g, 0), g
      assert int(m.geom_contype[i]) == col["contype"] and int(m.geom_conaffinity[i]) == col["conaffinity"], g
      fr = _first(col["friction"], g)
      if fr is not None:
        assert m.geom_friction[i, 0] == fr[0], g
      si = _first(col["solimp"] or {}, g)
      if si is not None:
        np.testing.assert_array_equal(m.geom_solimp[i, : len(si)], si)
      feet += bool(re.match(FOOT[name], g))
    else:
      assert int(m.geom_contype[i]) == 0 and int(m.geom_conaffinity[i]) == 0, g
  assert feet == ref["foot_geoms"]


def test_action_scale_and_soft_limits(robot):
  """`G1_ACTION_SCALE` / `GO1_ACTION_SCALE` (0.25 * effort_limit / stiffness per joint pattern) and the soft joint limit factor, as
  the env's resolved table carries them."""
  name, m, ref = robot
  c = envcfg.velocity_cfg(m, name)
  jn = _local(c.joint_names)
  want = np.array([_first(ref["action_scale"], j) for j in jn], float)
  np.testing.assert_allclose(c.action_scale, want, rtol=1e-12)
  lim = np.asarray(m.jnt_range)[c.joint_ids]
  mid, half = lim.mean(1), 0.5 * (lim[:, 1] - lim[:, 0]) * ref["soft_joint_pos_limit_factor"]
  np.testing.assert_allclose(c.soft_joint_pos_limits, np.stack([mid - half, mid + half], 1), rtol=1e-12, atol=1e-12)
