"""Actuator domain-randomisation events on the GPU env (envs/mdp/events.py:519-744): a consumed cfg's startup
`randomize_pd_gains` / `randomize_effort_limits` / `randomize_encoder_bias` give every world its own gains and limits, and the
physics kernels use them (actuator_force per world = clamp(kp (ctrl - q) - kd qd, -/+ limit) with THAT world's numbers)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_startup_pd_gain_and_effort_limit_events_reach_the_physics():
  from mjlab_b200 import events
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from tests.test_gpu_sim_api import _task_cfg
  d = _task_cfg("Mjlab-Velocity-Flat-Unitree-Go1")
  ac = dict(name="robot")   # all actuator objects (actuator_ids index `entity.actuators`, one object per actuator cfg)
  d["events"]["gains"] = dict(func="randomize_pd_gains", mode="startup", interval_range_s=None,
                              params=dict(kp_range=[0.5, 1.5], kd_range=[0.5, 2.0], operation="scale", distribution="log_uniform", asset_cfg=ac))
  d["events"]["effort"] = dict(func="randomize_effort_limits", mode="startup", interval_range_s=None,
                               params=dict(effort_limit_range=[0.2, 1.0], operation="scale", asset_cfg=ac))
  d["events"]["enc"] = dict(func="randomize_encoder_bias", mode="startup", interval_range_s=None, params=dict(bias_range=[-0.01, 0.01]))
  n = 64
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=n, seed=7), DEV)
  sim, mj = env.sim, env.sim.mj_model
  nu = mj.nu
  assert len(env.robot.actuators) == len(__import__("mjlab_b200.robots", fromlist=["go1_cfg"]).go1_cfg()[0])
  assert sorted(int(i) for a in env.robot.actuators for i in a.ctrl_ids) == list(range(nu))
  g, b, fr = sim.model.actuator_gainprm, sim.model.actuator_biasprm, sim.model.actuator_forcerange
  assert g.shape == (n, nu, 10) and g.stride(0) == nu * 10 and fr.stride(0) == nu * 2
  g0 = torch.as_tensor(mj.actuator_gainprm, dtype=torch.float32, device=DEV)
  b0 = torch.as_tensor(mj.actuator_biasprm, dtype=torch.float32, device=DEV)
  f0 = torch.as_tensor(mj.actuator_forcerange, dtype=torch.float32, device=DEV)
  kp, kd, el = g[:, :, 0] / g0[:, 0], b[:, :, 2] / b0[:, 2], fr[:, :, 1] / f0[:, 1]
  assert 0.5 <= float(kp.min()) and float(kp.max()) <= 1.5 and float(kp.std()) > 0.1
  assert 0.5 <= float(kd.min()) and float(kd.max()) <= 2.0 and float(kd.std()) > 0.1
  assert 0.2 <= float(el.min()) and float(el.max()) <= 1.0 and float(el.std()) > 0.1
  torch.testing.assert_close(b[:, :, 1], -g[:, :, 0])                       # bias[1] = -kp stays tied to the gain
  torch.testing.assert_close(fr[:, :, 0], -fr[:, :, 1])
  eb = env.robot.data.encoder_bias
  assert float(eb.abs().max()) <= 0.01 and float(eb.std()) > 1e-3
  env.reset()
  gen = torch.Generator(device=DEV).manual_seed(1)
  for _ in range(3):
    env.step(torch.rand(n, nu, device=DEV, generator=gen) * 4 - 2)          # large actions: many joints saturate at their limit
  sim.forward()
  torch.cuda.synchronize()
  D = sim.data
  jq = torch.as_tensor(mj.jnt_qposadr[mj.actuator_trnid[:, 0]], device=DEV).long()
  jv = torch.as_tensor(mj.jnt_dofadr[mj.actuator_trnid[:, 0]], device=DEV).long()
  want = g[:, :, 0] * D.ctrl + b[:, :, 1] * D.qpos[:, jq] + b[:, :, 2] * D.qvel[:, jv]
  want = torch.maximum(torch.minimum(want, fr[:, :, 1]), fr[:, :, 0])
  torch.testing.assert_close(D.actuator_force, want, rtol=1e-5, atol=1e-4)
  sat = (D.actuator_force.abs() == fr[:, :, 1])
  assert 0.005 < float(sat.float().mean()) < 0.95                           # both branches of the clamp occur
  # "scale" on the gains starts from the compiled values (no accumulation); on the limits it multiplies the current ones
  before = fr.clone()
  events.randomize_pd_gains(env, None, (2.0, 2.0), (1.0, 1.0), operation="scale")
  torch.testing.assert_close(sim.model.actuator_gainprm[:, :, 0], (2.0 * g0[:, 0]).expand(n, nu))
  events.randomize_effort_limits(env, torch.arange(8, device=DEV), (0.5, 0.5), operation="scale")
  torch.testing.assert_close(sim.model.actuator_forcerange[:8], 0.5 * before[:8])
  torch.testing.assert_close(sim.model.actuator_forcerange[8:], before[8:])


def test_external_force_event_pushes_the_selected_body():
  from mjlab_b200 import events
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from mjlab_b200.managers import SceneEntityCfg
  n = 16
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=n, seed=2), DEV)
  env.reset()
  ac = SceneEntityCfg("robot", body_names=["trunk"])
  ac.resolve(env.scene)
  ids = torch.arange(0, n, 2, device=DEV)
  env.sim.forward()
  acc0 = env.sim.data.qacc[:, :2].clone()
  events.apply_external_force_torque(env, ids, force_range=(40.0, 40.0), torque_range=(0.0, 0.0), asset_cfg=ac)
  trunk = int(env.robot.indexing.body_ids[ac.body_ids[0]])
  x = env.sim.data.xfrc_applied
  assert torch.all(x[ids, trunk, :3] == 40.0) and float(x[1::2].abs().max()) == 0.0 and float(x[ids][:, :, 3:].abs().max()) == 0.0
  env.sim.forward()
  torch.cuda.synchronize()
  dacc = env.sim.data.qacc[:, :2] - acc0
  # 40 N along +x and +y on the trunk of a 12.7 kg robot: the pushed worlds' base accelerates by m/s^2 more (same state, same
  # controls: a paired comparison), the others are bit-identical
  assert float(dacc[ids].min()) > 1.0 and float(dacc[1::2].abs().max()) == 0.0


def test_reward_weight_curriculum_changes_the_kernel_weights_at_its_steps():
  """`reward_weight` (tasks/velocity/mdp/curriculums.py:95-107) on the fused env: the weight of the named term switches when
  common_step_counter passes a stage's step; the per-term step reward of the kernel follows (weight x raw term)."""
  from mjlab_b200 import envcfg
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from tests.test_gpu_sim_api import _task_cfg
  d = _task_cfg("Mjlab-Velocity-Flat-Unitree-Go1")
  d["rewards"]["power"] = dict(func="electrical_power_cost", weight=0.0, params={})
  d["curriculum"]["power_ramp"] = dict(func="reward_weight", params=dict(reward_name="power", weight_stages=[dict(step=2, weight=-1e-3), dict(step=4, weight=-3e-3)]))
  n = 32
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=n, seed=1), DEV)
  k = envcfg.REWARD_NAMES.index("electrical_power_cost")
  assert "electrical_power_cost" in env.reward_manager.active_terms
  env.reset()
  gen = torch.Generator(device=DEV).manual_seed(0)
  seen = []
  for t in range(7):
    env.step(torch.rand(n, env.vcfg.nj, device=DEV, generator=gen) * 2 - 1)
    # step t+1 has been taken with the weight in force BEFORE it; the curriculum then updates for the next one
    seen.append((float(env.vcfg.reward_weights[k]), env._buf["step_reward"][:, k].clone()))
  ws = [w for w, _ in seen]
  assert ws == [0.0, 0.0, -1e-3, -1e-3, -3e-3, -3e-3, -3e-3]          # common_step_counter > 2 after the 3rd step, > 4 after the 5th
  assert float(seen[1][1].abs().max()) == 0.0 and float(seen[2][1].abs().max()) == 0.0      # steps 2 and 3 ran with weight 0
  assert float(seen[3][1].max()) <= 0.0 and float(seen[3][1].min()) < 0.0                   # step 4: -1e-3 x power
  # same raw power, three times the weight: ratio of the per-term rewards of consecutive steps is bounded by the physics only;
  # compare against the raw term recomputed from the state the kernel saw (actuator_force x joint velocity, positive part)
  D, mj = env.sim.data, env.sim.mj_model
  jv = torch.as_tensor(mj.jnt_dofadr[mj.actuator_trnid[:, 0]], device=DEV).long()
  alive = ~env.reset_buf
  raw = (D.actuator_force * D.qvel[:, jv]).clamp(min=0).sum(1)
  # (the forward pass after resets refreshed actuator_force for ALL envs from the post-step state, so only the sign / scale is checked)
  assert float(seen[6][1][alive].min()) < 0.0 and float(raw[alive].min()) >= 0.0


def test_builtin_sensor_observation_terms_outside_the_bank():
  """A cfg that adds the accelerometer and the root angular momentum to the critic: the extra columns are the physics kernels'
  sensordata slices of the state the observations were computed from."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from tests.test_gpu_sim_api import _task_cfg
  d = _task_cfg("Mjlab-Velocity-Flat-Unitree-Go1")
  tmpl = d["observations"]["critic"]["terms"]["foot_height"]
  d["observations"]["critic"]["terms"]["acc"] = dict(tmpl, func="builtin_sensor", params={"sensor_name": "robot/imu_lin_acc"})
  d["observations"]["critic"]["terms"]["angmom"] = dict(tmpl, func="builtin_sensor", params={"sensor_name": "robot/root_angmom"})
  n = 32
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=n, seed=3), DEV)
  mj = env.sim.mj_model
  sn = list(mj.sensor_names)
  a0, m0 = int(mj.sensor_adr[sn.index("robot/imu_lin_acc")]), int(mj.sensor_adr[sn.index("robot/root_angmom")])
  obs, _ = env.reset()
  base = env._raw_obs["critic"].shape[1]
  assert obs["critic"].shape == (n, base + 6) and env.observation_manager.group_obs_dim["critic"] == (base + 6,)
  gen = torch.Generator(device=DEV).manual_seed(0)
  for _ in range(4):
    obs, *_ = env.step(torch.rand(n, env.vcfg.nj, device=DEV, generator=gen) * 2 - 1)
    sd = env.sim.data.sensordata
    torch.testing.assert_close(obs["critic"][:, base:base + 3], sd[:, a0:a0 + 3])
    torch.testing.assert_close(obs["critic"][:, base + 3:], sd[:, m0:m0 + 3])
    torch.testing.assert_close(obs["critic"][:, :base], env._raw_obs["critic"])
  assert float(obs["critic"][:, base:base + 3].abs().max()) > 1.0       # an accelerometer at rest reads ~g
