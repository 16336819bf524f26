"""Mirror of the reference's tests/test_events.py:123-372 (randomize_pd_gains, randomize_effort_limits) plus the other
host-side event terms of envs/mdp/events.py (apply_external_force_torque, randomize_encoder_bias, sync_actuator_delays).
These terms are startup / reset-time torch indexing on the engine's model and data views - no kernel - so they run on CPU
with the same Mock environment the reference's tests build."""
from unittest.mock import Mock

import pytest
import torch

from mjlab_b200 import actuators as A
from mjlab_b200 import events
from mjlab_b200.managers import SceneEntityCfg

device = "cpu"


def _env_with_actuators():
  env = Mock()
  env.num_envs = 2
  env.device = device
  ent = Mock()
  builtin = A.BuiltinPositionActuator(torch.tensor([0, 1]), torch.tensor([0, 1]), ["a", "b"])
  xml = A.BuiltinPositionActuator(torch.tensor([2, 3]), torch.tensor([2, 3]), ["c", "d"])
  ideal = Mock(spec=["ctrl_ids", "stiffness", "damping", "default_stiffness", "default_damping", "force_limit", "set_gains", "set_effort_limit"])
  ideal.ctrl_ids = torch.tensor([4, 5])
  ideal.stiffness = torch.full((2, 2), 100.0)
  ideal.damping = torch.full((2, 2), 10.0)
  ideal.default_stiffness = torch.full((2, 2), 100.0)
  ideal.default_damping = torch.full((2, 2), 10.0)
  ideal.force_limit = torch.full((2, 2), 50.0)
  ideal.set_gains = A.IdealPdActuator.set_gains.__get__(ideal)
  ideal.set_effort_limit = A.IdealPdActuator.set_effort_limit.__get__(ideal)
  ent.actuators = [builtin, xml, ideal]
  env.scene = {"robot": ent}
  env.sim = Mock()
  env.sim.model = Mock()
  env.sim.model.actuator_gainprm = torch.ones((2, 6, 10)) * 50.0
  env.sim.model.actuator_biasprm = torch.zeros((2, 6, 10))
  env.sim.model.actuator_biasprm[:, :, 1] = -50.0
  env.sim.model.actuator_biasprm[:, :, 2] = -5.0
  env.sim.model.actuator_forcerange = torch.zeros((2, 6, 2))
  env.sim.model.actuator_forcerange[:, :, 0] = -100.0
  env.sim.model.actuator_forcerange[:, :, 1] = 100.0
  defaults = {"actuator_gainprm": torch.ones((6, 10)) * 50.0, "actuator_biasprm": torch.zeros((6, 10))}
  defaults["actuator_biasprm"][:, 1] = -50.0
  defaults["actuator_biasprm"][:, 2] = -5.0
  env.sim.get_default_field = lambda f: defaults[f]
  return env, ideal


def test_randomize_pd_gains():
  env, ideal = _env_with_actuators()
  M = env.sim.model
  events.randomize_pd_gains(env, torch.tensor([0]), kp_range=(1.5, 1.5), kd_range=(2.0, 2.0), asset_cfg=SceneEntityCfg("robot"),
                            distribution="uniform", operation="scale")
  for grp in ([0, 1], [2, 3]):
    assert torch.allclose(M.actuator_gainprm[0, grp, 0], torch.tensor([75.0, 75.0]))
    assert torch.allclose(M.actuator_biasprm[0, grp, 1], torch.tensor([-75.0, -75.0]))
    assert torch.allclose(M.actuator_biasprm[0, grp, 2], torch.tensor([-10.0, -10.0]))
    assert torch.allclose(M.actuator_gainprm[1, grp, 0], torch.tensor([50.0, 50.0]))     # env 1 untouched
  assert torch.allclose(ideal.stiffness[0], torch.tensor([150.0, 150.0]))
  assert torch.allclose(ideal.damping[0], torch.tensor([20.0, 20.0]))
  # "scale" starts from the defaults: a second call does not accumulate (events.py:535-537)
  events.randomize_pd_gains(env, torch.tensor([0]), kp_range=(1.5, 1.5), kd_range=(2.0, 2.0), asset_cfg=SceneEntityCfg("robot"))
  assert torch.allclose(M.actuator_gainprm[0, [0, 1], 0], torch.tensor([75.0, 75.0]))
  assert torch.allclose(ideal.stiffness[0], torch.tensor([150.0, 150.0]))
  events.randomize_pd_gains(env, torch.tensor([1]), kp_range=(200.0, 200.0), kd_range=(25.0, 25.0), asset_cfg=SceneEntityCfg("robot"),
                            distribution="uniform", operation="abs")
  for grp in ([0, 1], [2, 3]):
    assert torch.allclose(M.actuator_gainprm[1, grp, 0], torch.tensor([200.0, 200.0]))
    assert torch.allclose(M.actuator_biasprm[1, grp, 1], torch.tensor([-200.0, -200.0]))
    assert torch.allclose(M.actuator_biasprm[1, grp, 2], torch.tensor([-25.0, -25.0]))
  assert torch.allclose(ideal.stiffness[1], torch.tensor([200.0, 200.0]))
  assert torch.allclose(ideal.damping[1], torch.tensor([25.0, 25.0]))
  # a subset of the actuators by id, all envs
  events.randomize_pd_gains(env, None, kp_range=(2.0, 2.0), kd_range=(1.0, 1.0), asset_cfg=SceneEntityCfg("robot", actuator_ids=[1]))
  assert torch.allclose(M.actuator_gainprm[:, [2, 3], 0], torch.full((2, 2), 100.0))
  assert torch.allclose(M.actuator_gainprm[0, [0, 1], 0], torch.tensor([75.0, 75.0]))
  with pytest.raises(ValueError, match="operation"):
    events.randomize_pd_gains(env, None, (1, 1), (1, 1), operation="add")


def test_randomize_pd_gains_draws_differ_per_env_and_actuator():
  env, _ = _env_with_actuators()
  torch.manual_seed(0)
  events.randomize_pd_gains(env, None, kp_range=(0.5, 2.0), kd_range=(0.5, 2.0), distribution="log_uniform")
  kp = env.sim.model.actuator_gainprm[:, :4, 0] / 50.0
  assert kp.min() >= 0.5 and kp.max() <= 2.0 and kp.unique().numel() == 8
  assert torch.allclose(env.sim.model.actuator_biasprm[:, :4, 1], -env.sim.model.actuator_gainprm[:, :4, 0])


def test_randomize_pd_gains_requires_per_world_fields():
  env, _ = _env_with_actuators()
  env.sim.model.actuator_gainprm = torch.ones((1, 6, 10)).expand(2, 6, 10)     # the shared (stride-0) view of an unexpanded field
  with pytest.raises(ValueError, match="expand_model_fields"):
    events.randomize_pd_gains(env, None, (1.5, 1.5), (2.0, 2.0))


def test_randomize_effort_limits():
  env, ideal = _env_with_actuators()
  F = env.sim.model.actuator_forcerange
  events.randomize_effort_limits(env, torch.tensor([0]), effort_limit_range=(2.0, 2.0), asset_cfg=SceneEntityCfg("robot"),
                                 distribution="uniform", operation="scale")
  for grp in ([0, 1], [2, 3]):
    assert torch.allclose(F[0, grp, 0], torch.tensor([-200.0, -200.0]))
    assert torch.allclose(F[0, grp, 1], torch.tensor([200.0, 200.0]))
  assert torch.allclose(ideal.force_limit[0], torch.tensor([100.0, 100.0]))
  events.randomize_effort_limits(env, torch.tensor([1]), effort_limit_range=(150.0, 150.0), asset_cfg=SceneEntityCfg("robot"),
                                 distribution="uniform", operation="abs")
  for grp in ([0, 1], [2, 3]):
    assert torch.allclose(F[1, grp, 0], torch.tensor([-150.0, -150.0]))
    assert torch.allclose(F[1, grp, 1], torch.tensor([150.0, 150.0]))
  assert torch.allclose(ideal.force_limit[1], torch.tensor([150.0, 150.0]))


def test_apply_external_force_torque_and_encoder_bias():
  env = Mock()
  env.num_envs, env.device = 4, device
  ent = Mock()
  ent.num_bodies, ent.num_joints = 3, 5
  ent.data.encoder_bias = torch.zeros(4, 5)
  calls = {}
  ent.write_external_wrench_to_sim = lambda f, t, env_ids=None, body_ids=None: calls.update(f=f, t=t, env_ids=env_ids, body_ids=body_ids)
  env.scene = {"robot": ent}
  events.apply_external_force_torque(env, torch.tensor([1, 3]), force_range=(-5.0, 5.0), torque_range=(0.5, 0.5), asset_cfg=SceneEntityCfg("robot"))
  assert calls["f"].shape == (2, 3, 3) and calls["t"].shape == (2, 3, 3) and calls["body_ids"] == slice(None)
  assert calls["f"].abs().max() <= 5.0 and torch.all(calls["t"] == 0.5) and calls["env_ids"].tolist() == [1, 3]
  events.apply_external_force_torque(env, None, (1.0, 1.0), (0.0, 0.0), asset_cfg=SceneEntityCfg("robot", body_ids=[2]))
  assert calls["f"].shape == (4, 1, 3) and calls["body_ids"] == [2]
  events.randomize_encoder_bias(env, torch.tensor([0, 2]), bias_range=(-0.01, 0.01), asset_cfg=SceneEntityCfg("robot"))
  b = ent.data.encoder_bias
  assert b[[0, 2]].abs().max() <= 0.01 and b[[0, 2]].abs().min() > 0 and torch.all(b[[1, 3]] == 0)
  events.randomize_encoder_bias(env, None, bias_range=(0.02, 0.02), asset_cfg=SceneEntityCfg("robot", joint_ids=[1, 4]))
  assert torch.all(b[:, [1, 4]] == 0.02) and torch.all(b[[1, 3]][:, [0, 2, 3]] == 0)


def test_sync_actuator_delays_one_lag_per_env():
  env = Mock()
  env.num_envs, env.device = 6, device
  ent = Mock()
  got = []
  d1, d2, plain = Mock(spec=["set_lags"]), Mock(spec=["set_lags"]), Mock(spec=["ctrl_ids"])
  d1.set_lags = lambda lags, ids: got.append((lags.clone(), ids.clone()))
  d2.set_lags = lambda lags, ids: got.append((lags.clone(), ids.clone()))
  ent.actuators = [d1, plain, d2]
  env.scene = {"robot": ent}
  events.sync_actuator_delays(env, None, lag_range=(1, 4), asset_cfg=SceneEntityCfg("robot"))
  assert len(got) == 2 and torch.equal(got[0][0], got[1][0]) and got[0][1].tolist() == list(range(6))
  assert got[0][0].min() >= 1 and got[0][0].max() <= 4
  ent.actuators = [plain]
  got.clear()
  events.sync_actuator_delays(env, None, lag_range=(1, 4), asset_cfg=SceneEntityCfg("robot"))
  assert not got
