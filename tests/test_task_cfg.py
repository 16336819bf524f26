"""The kernels' built-in task tables (`envcfg.velocity_cfg`, `env.VelocityEnvCfg`) against THE REFERENCE'S OWN configuration
objects: tests/golden/task_cfg/*.json are dumps of `load_env_cfg(task, play)` (tools/make_cfg_golden.py runs the reference's
tasks/registry.py under tools/refstub.py).  `envcfg.from_manager_cfg` - the path a user's ManagerBasedRlEnvCfg takes - must
resolve each dump to exactly the built-in configuration: every reward weight and parameter, noise range, termination, reset /
push / command range, curriculum stage, episode length.  Also: unsupported terms raise instead of being ignored."""
import copy
import json
from pathlib import Path

import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import envcfg

CFG = Path(__file__).resolve().parent / "golden" / "task_cfg"
CASES = [("Mjlab-Velocity-Flat-Unitree-G1", "g1", "flat"), ("Mjlab-Velocity-Flat-Unitree-Go1", "go1", "flat"),
         ("Mjlab-Velocity-Rough-Unitree-G1", "g1", "rough"), ("Mjlab-Velocity-Rough-Unitree-Go1", "go1", "rough")]
SCALARS = ("decimation", "step_dt", "max_episode_length", "max_episode_length_s", "limit_angle", "std_lin", "std_ang", "std_upright",
           "walking_threshold", "running_threshold", "air_thr_min", "air_thr_max", "air_cmd_threshold", "target_height", "cmd_threshold",
           "cmd_resample", "cmd_lin_x", "cmd_lin_y", "cmd_ang_z", "cmd_heading", "rel_standing", "rel_heading", "heading_stiffness",
           "reset_joint_pos", "reset_joint_vel", "enable_corruption", "push_enabled", "heading_command", "terrain_mode",
           "upright_body", "upright_body_local")
ARRAYS = ("reward_weights", "std_standing", "std_walking", "std_running", "noise_lo", "noise_hi", "action_scale", "reset_pose", "reset_vel")


def _load(task, play):
  return json.loads((CFG / f"{task}{'-play' if play else ''}.json").read_text())


@pytest.mark.parametrize("task,robot,terrain", CASES)
@pytest.mark.parametrize("play", [False, True])
def test_builtin_task_tables_equal_the_reference_cfg(task, robot, terrain, play):
  d = _load(task, play)
  m = mjlab_b200.load_model(f"{robot}_{terrain}" + ("_play" if (play and terrain == "rough") else ""))
  got = envcfg.from_manager_cfg(d, m, robot)
  want = envcfg.velocity_cfg(m, robot, play=play)
  if play and robot == "g1" and terrain == "flat":        # env.ManagerBasedRlEnv applies the G1 play override of the command ranges (config/g1/env_cfgs.py:167-173)
    want.cmd_lin_x, want.cmd_ang_z = (-1.5, 2.0), (-0.7, 0.7)
  for k in SCALARS:
    a, b = getattr(got, k), getattr(want, k)
    if isinstance(b, tuple):
      np.testing.assert_allclose(a, b, rtol=1e-12, err_msg=k)
    else:
      assert a == pytest.approx(b, rel=1e-12), k
  for k in ARRAYS:
    np.testing.assert_allclose(getattr(got, k), getattr(want, k), rtol=1e-12, atol=0, err_msg=k)
  if want.push_enabled:
    np.testing.assert_allclose(got.push_vel, want.push_vel)
    np.testing.assert_allclose(got.push_interval, want.push_interval)
  if terrain == "rough":
    assert got.terrain_half_size == want.terrain_half_size and got.max_init_terrain_level == want.max_init_terrain_level
  # env-level settings: foot friction DR, velocity curriculum stages, sim options
  from mjlab_b200.env import VelocityEnvCfg, _default_sim_cfg
  ecfg = VelocityEnvCfg(robot=robot, terrain=terrain, play=play)
  ff = got.extra["foot_friction"]
  assert ff is not None and tuple(ff["ranges"]) == tuple(ecfg.foot_friction_range)
  prims = [f"{s}_foot{i}_collision" for s in ("left", "right") for i in range(1, 8)] if robot == "g1" else [f"{leg}_foot_collision" for leg in ("FR", "FL", "RR", "RL")]
  assert sorted(ff["geom_names"]) == sorted(prims)
  stages = got.extra["velocity_stages"]
  assert len(stages) == len(ecfg.velocity_stages)
  for s, w in zip(stages, ecfg.velocity_stages):
    assert s["step"] == w["step"] and all(tuple(s[k]) == tuple(w[k]) for k in w if k != "step") and set(s) == set(w)
  sc = _default_sim_cfg()
  assert got.extra["sim"] == dict(timestep=sc.mujoco.timestep, iterations=sc.mujoco.iterations, ls_iterations=sc.mujoco.ls_iterations,
                                  integrator=sc.mujoco.integrator, nconmax=sc.nconmax, njmax=sc.njmax)
  # no term asks for clip / scale / delay / history in the shipped tasks: the post-processing pipeline is the identity
  for group in ("policy", "critic"):
    assert all(t["clip"] is None and t["scale"] is None and t["delay_max_lag"] == 0 and t["history_length"] == 0 for t in got.extra["obs_post"][group])


def test_unsupported_terms_raise():
  m = mjlab_b200.load_model("g1_flat")
  base = _load("Mjlab-Velocity-Flat-Unitree-G1", False)

  def bad(edit, match):
    d = copy.deepcopy(base)
    edit(d)
    with pytest.raises(envcfg.UnsupportedTermError, match=match):
      envcfg.from_manager_cfg(d, m, "g1")

  bad(lambda d: d["rewards"].update(my_term=dict(func="my_reward", weight=1.0, params={})), "my_term")
  bad(lambda d: d["terminations"].update(x=dict(func="my_termination", time_out=False, params={})), "my_termination")
  bad(lambda d: d["terminations"].update(x=dict(func="nan_detection", time_out=True, params={})), "time_out=True")
  bad(lambda d: d["rewards"].update(jv=dict(func="joint_vel_l2", weight=-1.0, params={"asset_cfg": {"name": "robot", "joint_names": (".*_knee_joint",)}})),
      "all actuated joints")
  bad(lambda d: d["observations"]["policy"]["terms"].update(h=dict(d["observations"]["policy"]["terms"]["joint_vel"], func="body_height", params={})),
      "term bank")
  bad(lambda d: d["observations"]["policy"]["terms"].update(h=dict(d["observations"]["policy"]["terms"]["joint_vel"], func="base_lin_vel", params={})),
      "carries no noise")
  # parameters are either consumed or must name what the kernels are wired to: never silently ignored
  bad(lambda d: d["rewards"]["foot_slip"]["params"]["asset_cfg"].update(site_names=["left_foot"]), "foot sites")
  bad(lambda d: d["rewards"]["body_ang_vel"]["params"]["asset_cfg"].update(body_names=["pelvis"]), "flat_orientation term")
  bad(lambda d: d["rewards"]["track_linear_velocity"]["params"].update(command_name="other"), "command_name")
  bad(lambda d: d["rewards"]["air_time"]["params"].update(sensor_name="hands_contact"), "sensor_name")
  bad(lambda d: d["rewards"]["action_rate_l2"]["params"].update(gain=2.0), "parameter 'gain' is not consumed")
  bad(lambda d: d["terminations"]["fell_over"]["params"].update(asset_cfg=dict(name="robot", body_names=["torso_link"])), "all actuated joints")
  bad(lambda d: d["observations"]["critic"]["terms"]["foot_height"]["params"]["asset_cfg"].update(site_names=["right_foot"]), "foot sites")
  bad(lambda d: d["events"]["reset_robot_joints"]["params"]["asset_cfg"].update(joint_names=[".*_knee_joint"]), "all actuated joints")
  bad(lambda d: d["observations"].pop("critic"), "'policy' and 'critic' are required")
  bad(lambda d: d["observations"]["policy"]["terms"]["joint_pos"]["params"].update(biased=True), "biased=True")
  bad(lambda d: (d["observations"]["critic"].update(enable_corruption=True),
                 d["observations"]["critic"]["terms"]["joint_vel"]["noise"].update(n_max=9.0)), "one noise range per term function")
  bad(lambda d: d["observations"]["critic"]["terms"]["foot_height"].update(noise=dict(__type__="UniformNoiseCfg", operation="add", n_min=-1, n_max=1)),
      "carry no noise")
  bad(lambda d: d["events"].update(x=dict(func="apply_external_force_torque", mode="reset", params={}, interval_range_s=None)), "apply_external_force_torque")
  bad(lambda d: d["commands"]["twist"].update(init_velocity_prob=0.5), "init_velocity_prob")
  bad(lambda d: d["sim"]["mujoco"].update(solver="pgs"), "supported subset")
  bad(lambda d: d["observations"]["policy"]["terms"]["joint_pos"].update(noise=dict(__type__="GaussianNoiseCfg", operation="add", mean=0, std=1)), "UniformNoiseCfg")


def test_reweighted_and_removed_terms_are_consumed():
  m = mjlab_b200.load_model("go1_flat")
  d = _load("Mjlab-Velocity-Flat-Unitree-Go1", False)
  d["rewards"]["foot_slip"]["weight"] = -0.7
  d["rewards"].pop("soft_landing")
  d["rewards"]["track_linear_velocity"]["params"]["std"] = 0.3
  d["events"].pop("push_robot")
  d["observations"]["policy"]["terms"]["joint_vel"]["noise"]["n_max"] = 2.5
  d["observations"]["policy"]["terms"]["base_ang_vel"]["delay_max_lag"] = 2
  d["observations"]["critic"]["history_length"] = 3
  # terms of the generic library (envs/mdp/rewards.py, terminations.py) that the shipped tasks do not use
  d["rewards"]["alive"] = dict(func="is_alive", weight=0.25, params={})
  d["rewards"]["torques"] = dict(func="joint_torques_l2", weight=-2e-4, params={"asset_cfg": {"name": "robot", "joint_names": None, "body_names": None}})
  d["rewards"]["act_acc"] = dict(func="action_acc_l2", weight=-0.01, params={})
  d["terminations"]["too_low"] = dict(func="root_height_below_minimum", time_out=False, params={"minimum_height": 0.17})
  d["terminations"]["nans"] = dict(func="nan_detection", time_out=False, params={})
  # observation groups are gathers from the kernels' term bank: drop a policy term (the usual sim-to-real edit), reorder, add a group
  pt = d["observations"]["policy"]["terms"]
  pt.pop("base_lin_vel")
  d["observations"]["policy"]["terms"] = {"command": pt["command"], **{k2: v for k2, v in pt.items() if k2 != "command"}}
  d["observations"]["feet"] = dict(d["observations"]["critic"], terms={k2: d["observations"]["critic"]["terms"][k2] for k2 in ("foot_contact", "foot_height")})
  d["observations"]["critic"]["terms"]["root_lin_vel"] = dict(d["observations"]["critic"]["terms"]["foot_height"], func="base_lin_vel", params={})
  c = envcfg.from_manager_cfg(d, m, "go1")
  k = envcfg.REWARD_NAMES.index
  lay = c.extra["obs_layout"]
  assert lay["critic"][-1] == ("root_lin_vel", 3, "extra", 0)        # envs/mdp/observations.py base_lin_vel: the kernels' side buffer
  assert c.extra["obs_custom"] and [t[0] for t in lay["policy"]] == ["command", "base_ang_vel", "projected_gravity", "joint_pos", "joint_vel", "actions"]
  assert lay["policy"][0][1:] == (3, "critic", 9 + 3 * c.nj)          # no noise on the command: the clean copy
  assert lay["policy"][1][1:] == (3, "policy", 3)                     # noisy imu_ang_vel
  assert [t[0] for t in lay["feet"]] == ["foot_contact", "foot_height"] and lay["feet"][0][2:] == ("critic", 12 + 3 * c.nj + 2 * c.nfeet)
  assert c.reward_weights[k("is_alive")] == 0.25 and c.reward_weights[k("joint_torques_l2")] == -2e-4 and c.reward_weights[k("action_acc_l2")] == -0.01
  assert c.reward_weights[k("joint_acc_l2")] == 0.0 and c.min_root_height == 0.17 and c.nan_termination
  assert c.reward_weights[k("foot_slip")] == -0.7 and c.reward_weights[k("soft_landing")] == 0.0 and c.std_lin == 0.3
  assert not c.push_enabled and c.noise_hi[9 + c.nj] == 2.5 and c.noise_lo[9 + c.nj] == -1.5
  post = c.extra["obs_post"]
  assert {t["name"]: t for t in post["policy"]}["base_ang_vel"]["delay_max_lag"] == 2 and all(t["history_length"] == 3 for t in post["critic"])


@pytest.mark.parametrize("play", [False, True])
def test_tracking_cfg_from_the_reference_cfg_equals_the_builtin_table(play):
  """`envcfg.tracking_from_manager_cfg(load_env_cfg("Mjlab-Tracking-Flat-Unitree-G1", play))` (its JSON dump) resolves to the very
  table the built-in tracking task uses: weights, stds, thresholds, bodies, ranges, sampling, noise, action scale, push event."""
  import dataclasses
  m = mjlab_b200.load_model("g1_flat")
  d = _load("Mjlab-Tracking-Flat-Unitree-G1", play)
  got = envcfg.tracking_from_manager_cfg(d, m, 500)
  want = envcfg.tracking_cfg(m, 500, play=play)
  for f in dataclasses.fields(want):
    if f.name == "extra":
      continue
    x, y = getattr(want, f.name), getattr(got, f.name)
    if isinstance(x, (str, type(None))):
      assert x == y, f.name
    else:
      np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=f.name)
  assert got.extra["encoder_bias_range"] == (-0.01, 0.01)
  assert [(e["field"], e["operation"]) for e in got.extra["startup_dr"]] == [("body_ipos", "add"), ("geom_friction", "abs")]


def test_tracking_cfg_edits_and_unsupported_terms():
  m = mjlab_b200.load_model("g1_flat")
  base = _load("Mjlab-Tracking-Flat-Unitree-G1", False)
  d = copy.deepcopy(base)
  d["rewards"]["motion_body_pos"].update(weight=2.0)
  d["rewards"]["motion_body_pos"]["params"]["std"] = 0.2
  d["rewards"].pop("self_collisions")
  d["terminations"].pop("anchor_ori")
  d["terminations"]["ee_body_pos"]["params"].update(threshold=0.4, body_names=["left_ankle_roll_link", "right_ankle_roll_link"])
  d["commands"]["motion"].update(sampling_mode="uniform", joint_position_range=[-0.2, 0.2])
  d["events"].pop("push_robot")
  d["episode_length_s"] = 6.0
  c = envcfg.tracking_from_manager_cfg(d, m, 500)
  k = envcfg.TRACKING_REWARD_NAMES.index
  assert c.reward_weights[k("motion_body_pos")] == 2.0 and c.std_body_pos == 0.2 and c.reward_weights[k("self_collisions")] == 0.0
  assert c.anchor_ori_threshold > 1e8 and c.ee_pos_threshold == 0.4 and len(c.ee_sel) == 2
  assert c.sampling_mode == "uniform" and c.joint_position_range == (-0.2, 0.2) and not c.push_enabled and c.max_episode_length == 300

  def bad(edit, match):
    x = copy.deepcopy(base)
    edit(x)
    with pytest.raises(envcfg.UnsupportedTermError, match=match):
      envcfg.tracking_from_manager_cfg(x, m, 500)

  bad(lambda x: x["rewards"].update(my=dict(func="my_reward", weight=1.0, params={})), "my")
  bad(lambda x: x["terminations"].update(t=dict(func="bad_orientation", time_out=False, params={"limit_angle": 1.0})), "bad_orientation")
  bad(lambda x: x["observations"]["policy"]["terms"].update(h=dict(x["observations"]["policy"]["terms"]["joint_vel"], func="foot_height", params={})), "term bank")
  bad(lambda x: x["observations"].pop("critic"), "'policy' and 'critic' are required")
  bad(lambda x: x["events"].update(e=dict(func="reset_joints_by_offset", mode="reset", params={}, interval_range_s=None)), "reset_joints_by_offset")
  bad(lambda x: x["observations"]["policy"]["terms"]["joint_pos"]["params"].update(biased=False), "biased")
  bad(lambda x: x["observations"]["critic"]["terms"]["joint_pos"]["params"].update(biased=True), "biased")
  bad(lambda x: x["rewards"]["motion_body_pos"]["params"].update(command_name="other"), "command_name")
  bad(lambda x: x["rewards"]["motion_body_pos"]["params"].update(body_names=["pelvis"]), "parameter 'body_names' is not consumed")


def test_startup_actuator_events_are_consumed():
  """randomize_pd_gains / randomize_effort_limits / randomize_encoder_bias (envs/mdp/events.py:519-744) as STARTUP events of a
  user's cfg come back in extra["startup_dr"] for `events.apply_startup_event`; the same functions in reset mode would need
  per-reset host work outside the step graph and are refused by name."""
  m = mjlab_b200.load_model("go1_flat")
  d = _load("Mjlab-Velocity-Flat-Unitree-Go1", False)
  ac = dict(name="robot")   # all actuator objects (actuator_ids index `entity.actuators`, one object per actuator cfg)
  d["events"]["gains"] = dict(func="randomize_pd_gains", mode="startup", interval_range_s=None,
                              params=dict(kp_range=[0.8, 1.2], kd_range=[0.9, 1.1], operation="scale", distribution="uniform", asset_cfg=ac))
  d["events"]["effort"] = dict(func="randomize_effort_limits", mode="startup", interval_range_s=None,
                               params=dict(effort_limit_range=[0.7, 1.0], asset_cfg=ac))
  d["events"]["enc"] = dict(func="randomize_encoder_bias", mode="startup", interval_range_s=None, params=dict(bias_range=[-0.01, 0.01]))
  c = envcfg.from_manager_cfg(d, m, "go1")
  got = {e["name"]: e for e in c.extra["startup_dr"] if e.get("kind")}
  assert set(got) == {"gains", "effort", "enc"}
  assert got["gains"]["kind"] == "randomize_pd_gains" and got["gains"]["params"] == dict(kp_range=(0.8, 1.2), kd_range=(0.9, 1.1), operation="scale", distribution="uniform")
  assert got["gains"]["asset_cfg"] == dict(name="robot", joint_names=None, actuator_names=None)
  assert got["effort"]["params"] == dict(effort_limit_range=(0.7, 1.0)) and got["enc"]["params"] == dict(bias_range=(-0.01, 0.01))
  names = [e["name"] for e in c.extra["startup_dr"]]
  assert names.index("gains") < names.index("effort") < names.index("enc")     # cfg order is kept
  d["events"]["gains"]["mode"] = "reset"
  with pytest.raises(envcfg.UnsupportedTermError, match="randomize_pd_gains, mode reset"):
    envcfg.from_manager_cfg(d, m, "go1")


def test_posture_and_electrical_power_cost_are_consumed():
  """envs/mdp/rewards.py:91-146.  `posture` (one std table) lands in the pose slot with three equal tables - the same
  arithmetic as `variable_posture` whatever the commanded speed; `electrical_power_cost` has a kernel slot of its own (pinned
  by the reference's class in tests/golden/mdp_velocity_*.npz)."""
  m = mjlab_b200.load_model("go1_flat")
  d = _load("Mjlab-Velocity-Flat-Unitree-Go1", False)
  all_j = dict(name="robot", joint_names=[".*"])
  ref = envcfg.from_manager_cfg(copy.deepcopy(d), m, "go1")
  d["rewards"]["power"] = dict(func="electrical_power_cost", weight=-2e-3, params=dict(asset_cfg=all_j))
  both = copy.deepcopy(d)
  both["rewards"]["posture2"] = dict(func="posture", weight=1.0, params=dict(std={".*": 0.3}, asset_cfg=all_j))
  with pytest.raises(envcfg.UnsupportedTermError, match="share one kernel slot"):
    envcfg.from_manager_cfg(both, m, "go1")
  pose_w = d["rewards"]["pose"]["weight"]
  d["rewards"].pop("pose")
  d["rewards"]["posture"] = dict(func="posture", weight=pose_w, params=dict(std={".*_hip_joint": 0.1, ".*_(thigh|calf)_joint": 0.3}, asset_cfg=all_j))
  c = envcfg.from_manager_cfg(d, m, "go1")
  k = envcfg.REWARD_NAMES.index
  assert c.reward_weights[k("electrical_power_cost")] == -2e-3 and ref.reward_weights[k("electrical_power_cost")] == 0
  assert c.reward_weights[k("pose")] == pose_w
  names = [n.removeprefix("robot/") for n in c.joint_names]
  want = np.array([0.1 if n.endswith("_hip_joint") else 0.3 for n in names])
  for key in ("standing", "walking", "running"):
    np.testing.assert_array_equal(getattr(c, f"std_{key}"), want)
  bad = copy.deepcopy(d)
  bad["rewards"]["power"]["params"]["asset_cfg"] = dict(name="robot", joint_names=[".*_hip_joint"])
  with pytest.raises(envcfg.UnsupportedTermError, match="all actuated joints"):
    envcfg.from_manager_cfg(bad, m, "go1")


def test_reward_weight_curriculum_is_consumed():
  """tasks/velocity/mdp/curriculums.py:95-107: stages of (step, weight) for a reward term named as in the cfg."""
  m = mjlab_b200.load_model("go1_flat")
  d = _load("Mjlab-Velocity-Flat-Unitree-Go1", False)
  d["rewards"]["torques"] = dict(func="joint_torques_l2", weight=0.0, params={})
  d["curriculum"]["torque_ramp"] = dict(func="reward_weight", params=dict(reward_name="torques", weight_stages=[dict(step=2, weight=-1e-4), dict(step=5, weight=-2e-4)]))
  d["curriculum"]["slip_ramp"] = dict(func="reward_weight", params=dict(reward_name="foot_slip", weight_stages=[dict(step=3, weight=-0.5)]))
  c = envcfg.from_manager_cfg(d, m, "go1")
  k = envcfg.REWARD_NAMES.index
  assert c.extra["reward_weight_stages"] == [(k("joint_torques_l2"), ((2, -1e-4), (5, -2e-4))), (k("foot_slip"), ((3, -0.5),))]
  d["curriculum"]["torque_ramp"]["params"]["reward_name"] = "nope"
  with pytest.raises(envcfg.UnsupportedTermError, match="no reward term named 'nope'"):
    envcfg.from_manager_cfg(d, m, "go1")


def test_any_builtin_sensor_is_an_observation_term():
  """builtin_sensor(sensor_name) (envs/mdp/observations.py:102-108) for sensors outside the kernels' term bank: a gather from
  sensordata; contact sensors and unknown names are refused by name."""
  m = mjlab_b200.load_model("go1_flat")
  d = _load("Mjlab-Velocity-Flat-Unitree-Go1", False)
  tmpl = d["observations"]["critic"]["terms"]["foot_height"]
  d["observations"]["critic"]["terms"]["acc"] = dict(tmpl, func="builtin_sensor", params={"sensor_name": "robot/imu_lin_acc"})
  d["observations"]["critic"]["terms"]["angmom"] = dict(tmpl, func="builtin_sensor", params={"sensor_name": "robot/root_angmom"})
  c = envcfg.from_manager_cfg(copy.deepcopy(d), m, "go1")
  sn = list(m.sensor_names)
  tail = c.extra["obs_layout"]["critic"][-2:]
  assert tail == [("acc", 3, "sensordata", int(m.sensor_adr[sn.index("robot/imu_lin_acc")])),
                  ("angmom", 3, "sensordata", int(m.sensor_adr[sn.index("robot/root_angmom")]))]
  assert c.extra["obs_custom"]
  for bad, match in (("robot/nope", "not in the scene"), ("feet_ground_contact_FR_foot_collision_found", "not in the scene")):
    x = copy.deepcopy(d)
    x["observations"]["critic"]["terms"]["acc"]["params"]["sensor_name"] = bad
    with pytest.raises(envcfg.UnsupportedTermError, match=match):
      envcfg.from_manager_cfg(x, m, "go1")


@pytest.mark.parametrize("play", [False, True])
def test_tracking_no_state_estimation_task_resolves_to_gathers(play):
  """The reference registers `Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation` (tasks/tracking/config/g1/__init__.py:15-22): the
  policy group without motion_anchor_pos_b and base_lin_vel (config/g1/env_cfgs.py:73-84).  Its dump resolves to the same kernel
  table as the full task plus a gather layout over the kernels' term bank; the full task stays on the kernel's buffers."""
  import dataclasses
  m = mjlab_b200.load_model("g1_flat")
  full = envcfg.tracking_from_manager_cfg(_load("Mjlab-Tracking-Flat-Unitree-G1", play), m, 500)
  assert not full.extra["obs_custom"]
  c = envcfg.tracking_from_manager_cfg(_load("Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation", play), m, 500)
  assert c.extra["obs_custom"]
  nj, nb = c.nj, int(c.nbody)
  lay = c.extra["obs_layout"]
  assert [t[0] for t in lay["policy"]] == ["command", "motion_anchor_ori_b", "base_ang_vel", "joint_pos", "joint_vel", "actions"]
  # a term with noise in a corrupted group reads the noisy copy (obs_policy), every other one the clean copy (obs_critic, where the
  # two body terms sit between motion_anchor_ori_b and base_lin_vel); joint_pos(biased=True) only exists in obs_policy
  dump = _load("Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation", play)["observations"]["policy"]
  off_p = {"command": 0, "motion_anchor_ori_b": 2 * nj + 3, "base_ang_vel": 2 * nj + 12, "joint_pos": 2 * nj + 15, "joint_vel": 3 * nj + 15, "actions": 4 * nj + 15}
  shift = 9 * nb
  nsrc = {"policy": 0, "critic": 0}
  for name, w, sg, o in lay["policy"]:
    noisy = dump["enable_corruption"] and dump["terms"][name]["noise"] is not None
    if name == "joint_pos" or noisy:
      assert (sg, o) == ("policy", off_p[name]), name
    else:
      assert (sg, o) == ("critic", off_p[name] + (shift if off_p[name] >= 2 * nj + 9 else 0)), name
    nsrc[sg] += 1
  assert nsrc["policy"] == (1 if play else 4) and nsrc["critic"] == (5 if play else 2)
  assert sum(t[1] for t in lay["policy"]) == c.n_obs_policy - 6 and sum(t[1] for t in lay["critic"]) == c.n_obs_critic
  assert all(t[2] == "critic" for t in lay["critic"])
  for f in dataclasses.fields(full):
    if f.name in ("extra", "noise_lo", "noise_hi"):
      continue
    x, y = getattr(full, f.name), getattr(c, f.name)
    if isinstance(x, (str, type(None))):
      assert x == y, f.name
    else:
      np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=f.name)
  # noise table: the two dropped terms' segments carry no noise any more, the others are the full task's
  keep = np.ones(c.n_obs_policy, bool)
  keep[2 * nj: 2 * nj + 3] = False
  keep[2 * nj + 9: 2 * nj + 12] = False
  np.testing.assert_array_equal(c.noise_lo[keep], full.noise_lo[keep])
  assert not c.noise_lo[~keep].any() and not c.noise_hi[~keep].any()


def test_tracking_3d_terminations_are_consumed():
  """bad_anchor_pos / bad_motion_body_pos (tasks/tracking/mdp/terminations.py:18-25,55-70) share the kernel slots of their z-only
  twins: the cfg chooses one of each pair."""
  m = mjlab_b200.load_model("g1_flat")
  d = _load("Mjlab-Tracking-Flat-Unitree-G1", False)
  z = envcfg.tracking_from_manager_cfg(copy.deepcopy(d), m, 500)
  assert z.term_full3d == 0
  d["terminations"]["anchor_pos"] = dict(d["terminations"]["anchor_pos"], func="bad_anchor_pos")
  d["terminations"]["anchor_pos"]["params"]["threshold"] = 0.3
  c = envcfg.tracking_from_manager_cfg(copy.deepcopy(d), m, 500)
  assert c.term_full3d == 1 and c.anchor_pos_threshold == 0.3 and c.ee_pos_threshold == z.ee_pos_threshold
  d["terminations"]["ee_body_pos"] = dict(d["terminations"]["ee_body_pos"], func="bad_motion_body_pos")
  c = envcfg.tracking_from_manager_cfg(copy.deepcopy(d), m, 500)
  assert c.term_full3d == 3 and list(c.ee_sel) == list(z.ee_sel)
  d["terminations"]["both"] = dict(d["terminations"]["anchor_pos"], func="bad_anchor_pos_z_only")
  with pytest.raises(envcfg.UnsupportedTermError, match="share one kernel slot"):
    envcfg.tracking_from_manager_cfg(d, m, 500)


def test_sim_options_of_the_cfg_are_consumed_or_refused():
  """`cfg.sim.mujoco` (src/mjlab/sim/sim.py:45-60): both friction cones, Newton and CG, both integrators are accepted; PGS / RK4 are
  refused by name.  The env built from such a cfg hands the options to `Simulation` (`ManagerBasedRlEnv.from_manager_cfg`)."""
  m = mjlab_b200.load_model("go1_flat")
  base = _load("Mjlab-Velocity-Flat-Unitree-Go1", False)
  for key, val in (("cone", "elliptic"), ("solver", "cg"), ("integrator", "euler")):
    d = copy.deepcopy(base)
    d["sim"]["mujoco"][key] = val
    envcfg.from_manager_cfg(d, m, "go1")
  for key, val in (("solver", "pgs"), ("integrator", "rk4"), ("cone", "spherical")):
    d = copy.deepcopy(base)
    d["sim"]["mujoco"][key] = val
    with pytest.raises(envcfg.UnsupportedTermError, match=val):
      envcfg.from_manager_cfg(d, m, "go1")
