"""Actuator ops and the EntityData views (a10/a11) on the GPU, against torch fp32
restatements of the reference formulas (actuator/pd_actuator.py:94-109, dc_actuator.py:136-162,
utils/buffers/delay_buffer.py, learned_actuator.py:144-207, entity/data.py) - mirroring tests/test_pd_actuator.py,
test_dc_actuator.py, test_delayed_actuator.py, test_learned_actuator.py, test_entity_data.py of the reference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cmd(n, nj, seed=0):
  from mjlab_b200.actuators import ActuatorCmd
  g = torch.Generator(device=DEV).manual_seed(seed)
  r = lambda s=1.0: (torch.rand(n, nj, device=DEV, generator=g) * 2 - 1) * s  # noqa: E731
  return ActuatorCmd(position_target=r(), velocity_target=r(2), effort_target=r(5), joint_pos=r(), joint_vel=r(30))


def test_ideal_pd_matches_formula():
  from mjlab_b200.actuators import IdealPdActuator
  n, nj = 257, 29
  act = IdealPdActuator(n, nj, stiffness=80.0, damping=2.5, effort_limit=20.0, device=DEV)
  ids = torch.arange(0, n, 2, device=DEV)
  act.set_gains(ids, kp=torch.full((len(ids), nj), 40.0, device=DEV), kd=torch.full((len(ids),), 1.0, device=DEV))   # (ids, nj) / one per env
  c = _cmd(n, nj)
  got = act.compute(c).clone()
  want = torch.clamp(act.stiffness * (c.position_target - c.joint_pos) + act.damping * (c.velocity_target - c.joint_vel) + c.effort_target,
                     -act.force_limit, act.force_limit)
  torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-5)
  assert (got.abs() <= 20.0).all() and (got.abs() == 20.0).any()


def test_dc_motor_matches_formula():
  from mjlab_b200.actuators import DcMotorActuator
  n, nj = 128, 12
  act = DcMotorActuator(n, nj, stiffness=50.0, damping=1.0, effort_limit=23.7, saturation_effort=35.0, velocity_limit=30.0, device=DEV)
  c = _cmd(n, nj, seed=1)
  got = act.compute(c).clone()
  tau = act.stiffness * (c.position_target - c.joint_pos) + act.damping * (c.velocity_target - c.joint_vel) + c.effort_target
  corner = act.velocity_limit_motor * (1 + act.force_limit / act.saturation_effort)
  v = torch.clamp(c.joint_vel, min=-corner, max=corner)
  top = act.saturation_effort * (1.0 - v / act.velocity_limit_motor)
  bot = act.saturation_effort * (-1.0 - v / act.velocity_limit_motor)
  want = torch.clamp(tau, min=torch.clamp(bot, min=-act.force_limit), max=torch.clamp(top, max=act.force_limit))
  torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-5)
  # stall torque / no-load speed corners (tests/test_dc_actuator.py): at v = velocity_limit the positive torque is zero
  from mjlab_b200.actuators import ActuatorCmd
  z = torch.zeros(n, nj, device=DEV)
  c2 = ActuatorCmd(z + 10, z, z, z, z + 30.0)
  assert torch.allclose(act.compute(c2), z, atol=1e-5)


def test_delay_buffer_semantics():
  from mjlab_b200.actuators import DelayBuffer
  n, w = 5, 3
  buf = DelayBuffer(min_lag=2, max_lag=2, batch_size=n, device=DEV)
  outs = []
  for t in range(6):
    x = torch.full((n, w), float(t + 1), device=DEV)
    buf.append(x)
    outs.append(buf.compute().clone())
  # first push back-fills, lag is clamped to what has been pushed: 1, 1, 1, 2, 3, 4
  assert [float(o[0, 0]) for o in outs] == [1.0, 1.0, 1.0, 2.0, 3.0, 4.0]
  buf.reset(torch.tensor([1, 3], device=DEV))
  buf.append(torch.full((n, w), 10.0, device=DEV))
  o = buf.compute()
  assert float(o[1, 0]) == 10.0 and float(o[3, 0]) == 10.0 and float(o[0, 0]) == 5.0
  # per-env lags
  buf2 = DelayBuffer(min_lag=0, max_lag=3, batch_size=4, device=DEV, update_period=1000)
  hist = []
  for t in range(5):
    x = torch.arange(4, device=DEV, dtype=torch.float32)[:, None] * 100 + t
    buf2.append(x)
    buf2.set_lags(torch.tensor([0, 1, 2, 3], device=DEV)) if t == 0 else None
    hist.append(buf2.compute().clone())
  last = hist[-1][:, 0].tolist()
  lags = buf2.current_lags.tolist()
  assert last == [w_ * 100 + 4 - l for w_, l in enumerate(lags)]


def test_delayed_actuator_delays_position_target():
  from mjlab_b200.actuators import ActuatorCmd, DelayedActuator, IdealPdActuator
  n, nj = 8, 4
  base = IdealPdActuator(n, nj, stiffness=10.0, damping=0.0, effort_limit=1e6, device=DEV)
  act = DelayedActuator(base, delay_target="position", delay_min_lag=1, delay_max_lag=1)
  z = torch.zeros(n, nj, device=DEV)
  outs = []
  for t in range(4):
    outs.append(act.compute(ActuatorCmd(z + float(t), z, z, z, z)).clone())
  assert [float(o[0, 0]) for o in outs] == [0.0, 0.0, 10.0, 20.0]


@pytest.mark.parametrize("order", ["pos_vel", "vel_pos"])
def test_learned_mlp_matches_torch(order):
  from mjlab_b200.actuators import LearnedMlpActuator
  torch.manual_seed(0)
  H = 3
  net = torch.nn.Sequential(torch.nn.Linear(2 * H, 32), torch.nn.Softsign(), torch.nn.Linear(32, 32), torch.nn.Softsign(), torch.nn.Linear(32, 1)).to(DEV)
  n, nj = 64, 12
  act = LearnedMlpActuator(n, nj, net, pos_scale=-1.0, vel_scale=1.0, torque_scale=1.0, effort_limit=23.7, saturation_effort=35.0,
                           velocity_limit=30.0, input_order=order, history_length=H, activation="softsign", device=DEV)
  ph = torch.zeros(n, nj, H, device=DEV)
  vh = torch.zeros(n, nj, H, device=DEV)
  for t in range(4):
    c = _cmd(n, nj, seed=10 + t)
    got = act.compute(c).clone()
    pe = c.position_target - c.joint_pos
    if t == 0:
      ph[:] = pe[..., None]; vh[:] = c.joint_vel[..., None]
    else:
      ph = torch.cat([pe[..., None], ph[..., :-1]], -1); vh = torch.cat([c.joint_vel[..., None], vh[..., :-1]], -1)
    p, v = (ph * -1.0).reshape(-1, H), (vh * 1.0).reshape(-1, H)
    x = torch.cat([p, v] if order == "pos_vel" else [v, p], 1)
    with torch.inference_mode():
      tau = net(x).reshape(n, nj) * 1.0
    corner = 30.0 * (1 + 23.7 / 35.0)
    vc = torch.clamp(c.joint_vel, -corner, corner)
    want = torch.clamp(tau, min=torch.clamp(35.0 * (-1 - vc / 30.0), min=-23.7), max=torch.clamp(35.0 * (1 - vc / 30.0), max=23.7))
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


# ---- the same classes against the REFERENCE'S OWN classes (tests/golden/actuators.npz, tools/make_actuator_golden.py) ----------
def _golden():
  from pathlib import Path
  return np.load(Path(__file__).resolve().parent / "golden" / "actuators.npz")


def _gcmd(g, prefix, t):
  from mjlab_b200.actuators import ActuatorCmd
  return ActuatorCmd(**{k: torch.as_tensor(g[f"{prefix}_{k}"][t], device=DEV) for k in
                        ("position_target", "velocity_target", "effort_target", "joint_pos", "joint_vel")})


class _Replay:
  """Feeds the draws recorded from the reference's DelayBuffer into mjlab_b200.actuators' `_randint` / `_rand` hooks."""

  def __init__(self, g, prefix, monkeypatch):
    from mjlab_b200 import actuators as ACT
    self.draws = [g[f"{prefix}_draw{i}"] for i in range(int(g[f"{prefix}_ndraws"]))]
    self.i = 0
    monkeypatch.setattr(ACT, "_randint", lambda lo, hi, shape, device, generator=None: self._next(shape, device, torch.long, lo, hi))
    monkeypatch.setattr(ACT, "_rand", lambda shape, device, generator=None: self._next(shape, device, torch.float32, 0, 1))

  def _next(self, shape, device, dtype, lo, hi):
    v = self.draws[self.i]
    self.i += 1
    assert tuple(v.shape) == tuple(shape), (self.i, v.shape, shape)     # same call sequence as the reference
    assert v.min() >= lo and v.max() < hi
    return torch.as_tensor(v, device=device).to(dtype)


def test_ideal_pd_and_dc_motor_match_reference_classes():
  from mjlab_b200.actuators import DcMotorActuator, IdealPdActuator
  g = _golden()
  T, B, NJ = g["pd_out"].shape
  act = IdealPdActuator(B, NJ, stiffness=80.0, damping=2.5, effort_limit=20.0, device=DEV)
  for t in range(T):
    if t == 4:
      act.set_gains(torch.tensor([0, 2, 4], device=DEV), kp=torch.full((3, NJ), 40.0, device=DEV), kd=torch.tensor([1.0, 1.5, 2.0], device=DEV))
    if t == 8:
      act.set_effort_limit(torch.tensor([1], device=DEV), torch.full((1, NJ), 5.0, device=DEV))
    np.testing.assert_allclose(act.compute(_gcmd(g, "pd", t)).cpu().numpy(), g["pd_out"][t], rtol=1e-6, atol=1e-5)
  dc = DcMotorActuator(B, NJ, stiffness=50.0, damping=1.0, effort_limit=23.7, saturation_effort=35.0, velocity_limit=30.0, device=DEV)
  for t in range(T):
    np.testing.assert_allclose(dc.compute(_gcmd(g, "dc", t)).cpu().numpy(), g["dc_out"][t], rtol=1e-6, atol=1e-5)
  assert (np.abs(g["dc_out"]) < 23.7 - 1e-3).any() and (np.abs(np.abs(g["dc_out"]) - 23.7) < 1e-5).any()   # both clips occur


def test_delay_buffer_matches_reference_class(monkeypatch):
  """Lag policy (period 3, per-env phase, hold probability 0.3, reset of two rows at t = 7) and the lagged reads, with the
  reference's own random draws replayed: current_lags and outputs equal the reference's at every step."""
  from mjlab_b200.actuators import DelayBuffer
  g = _golden()
  x = torch.as_tensor(g["delay_in"], device=DEV)
  T, B = x.shape[:2]
  rp = _Replay(g, "delay", monkeypatch)
  buf = DelayBuffer(min_lag=1, max_lag=4, batch_size=B, device=DEV, update_period=3, per_env_phase=True, hold_prob=0.3)
  for t in range(T):
    if t == 7:
      buf.reset(torch.tensor([1, 4], device=DEV))
    buf.append(x[t])
    y = buf.compute()
    np.testing.assert_array_equal(buf.current_lags.cpu().numpy(), g["delay_lags"][t])
    np.testing.assert_array_equal(y.cpu().numpy(), g["delay_out"][t])
  assert rp.i == len(rp.draws)
  rp = _Replay(g, "delay2", monkeypatch)
  buf = DelayBuffer(min_lag=0, max_lag=3, batch_size=B, device=DEV, per_env=False)
  for t in range(T):
    buf.append(x[t])
    y = buf.compute()
    np.testing.assert_array_equal(buf.current_lags.cpu().numpy(), g["delay2_lags"][t])
    np.testing.assert_array_equal(y.cpu().numpy(), g["delay2_out"][t])
  assert rp.i == len(rp.draws)


def test_delayed_actuator_matches_reference_class(monkeypatch):
  from mjlab_b200.actuators import DelayedActuator, IdealPdActuator
  g = _golden()
  T, B, NJ = g["dly_out"].shape
  rp = _Replay(g, "dly", monkeypatch)
  act = DelayedActuator(IdealPdActuator(B, NJ, stiffness=10.0, damping=0.5, effort_limit=50.0, device=DEV),
                        delay_target=("position", "effort"), delay_min_lag=0, delay_max_lag=3, delay_update_period=4,
                        delay_hold_prob=0.0, delay_per_env_phase=True)
  for t in range(T):
    if t == 5:
      act.set_lags(torch.tensor([3, 0], device=DEV), torch.tensor([0, 5], device=DEV))
    if t == 9:
      act.reset(torch.tensor([2], device=DEV))
    np.testing.assert_allclose(act.compute(_gcmd(g, "dly", t)).cpu().numpy(), g["dly_out"][t], rtol=1e-6, atol=1e-5)
  assert rp.i == len(rp.draws)


@pytest.mark.parametrize("order", ["pos_vel", "vel_pos"])
def test_learned_mlp_matches_reference_class(order):
  from mjlab_b200.actuators import LearnedMlpActuator
  g = _golden()
  T, B, NJ = g[f"mlp_out_{order}"].shape
  H = 3
  net = torch.nn.Sequential(torch.nn.Linear(2 * H, 16), torch.nn.Softsign(), torch.nn.Linear(16, 16), torch.nn.Softsign(), torch.nn.Linear(16, 1))
  with torch.no_grad():
    for i, lyr in enumerate(l for l in net if isinstance(l, torch.nn.Linear)):
      lyr.weight.copy_(torch.as_tensor(g[f"mlp_w{i}"])); lyr.bias.copy_(torch.as_tensor(g[f"mlp_b{i}"]))
  act = LearnedMlpActuator(B, NJ, net, pos_scale=-1.0, vel_scale=0.05, torque_scale=8.0, effort_limit=23.7, saturation_effort=35.0,
                           velocity_limit=30.0, input_order=order, history_length=H, activation="softsign", device=DEV)
  for t in range(T):
    if t == 6:
      act.reset(torch.tensor([1, 3], device=DEV))
    np.testing.assert_allclose(act.compute(_gcmd(g, "mlp", t)).cpu().numpy(), g[f"mlp_out_{order}"][t], rtol=2e-5, atol=2e-5)


def test_circular_buffer_on_device_matches_reference_class():
  from mjlab_b200.buffers import CircularBuffer
  g = _golden()
  x = torch.as_tensor(g["delay_in"], device=DEV)
  B = x.shape[1]
  cb = CircularBuffer(max_len=4, batch_size=B, device=DEV)
  for t in range(8):
    if t == 5:
      cb.reset([0, 3])
    cb.append(x[t])
    for lag in range(6):
      np.testing.assert_array_equal(cb[lag].cpu().numpy(), g["circ_out"][t, lag])
    np.testing.assert_array_equal(cb.buffer.cpu().numpy(), g["circ_buffer"][t])


def _sim(robot="g1", n=16):
  import mjlab_b200
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  m = mjlab_b200.load_model(f"{robot}_flat")
  return Simulation(n, SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20)), m, DEV)


def test_entity_data_velocity_roundtrips_and_staleness():
  """tests/test_entity_data.py:44-158 of the reference: world/body frame round trips; reads are stale until forward()."""
  from mjlab_b200.entity import Entity
  sim = _sim()
  robot = Entity("robot", sim)
  n = sim.num_envs
  g = torch.Generator(device=DEV).manual_seed(3)
  pose = robot.data.default_root_state[:, :7].clone()
  q = torch.randn(n, 4, device=DEV, generator=g)
  pose[:, 3:7] = q / q.norm(dim=1, keepdim=True)
  pose[:, 2] += 0.5
  robot.write_root_link_pose_to_sim(pose)
  vel = torch.randn(n, 6, device=DEV, generator=g)
  robot.write_root_link_velocity_to_sim(vel)
  before = robot.data.root_link_pose_w.clone()
  assert not torch.allclose(before, pose, atol=1e-4)          # stale until forward
  sim.forward()
  torch.testing.assert_close(robot.data.root_link_pose_w, pose, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(robot.data.root_link_vel_w, vel, atol=1e-4, rtol=1e-4)
  torch.testing.assert_close(robot.data.root_link_lin_vel_w, vel[:, :3], atol=1e-4, rtol=1e-4)
  from mjlab_b200.entity import quat_apply_inverse
  torch.testing.assert_close(robot.data.root_link_ang_vel_b, quat_apply_inverse(pose[:, 3:7], vel[:, 3:]), atol=1e-4, rtol=1e-4)
  torch.testing.assert_close(sim.data.qvel[:, 3:6], quat_apply_inverse(pose[:, 3:7], vel[:, 3:]), atol=1e-5, rtol=1e-5)
  # shapes of every property family
  nb, ng, ns, nj = len(robot.body_names), len(robot.geom_names), len(robot.site_names), robot.num_joints
  d = robot.data
  assert d.body_link_pose_w.shape == (n, nb, 7) and d.body_com_vel_w.shape == (n, nb, 6) and d.geom_pose_w.shape == (n, ng, 7)
  assert d.site_vel_w.shape == (n, ns, 6) and d.joint_pos.shape == (n, nj) and d.actuator_force.shape == (n, nj)
  assert d.projected_gravity_b.shape == (n, 3) and d.heading_w.shape == (n,) and d.body_link_quat_w.shape == (n, nb, 4)
  # joint writes with env / joint subsets
  ids = torch.tensor([1, 5], device=DEV)
  jids = torch.tensor([0, 3, 7], device=DEV)
  robot.write_joint_state_to_sim(torch.full((2, 3), 0.25, device=DEV), torch.full((2, 3), -1.0, device=DEV), joint_ids=jids, env_ids=ids)
  assert torch.allclose(d.joint_pos[ids][:, jids], torch.full((2, 3), 0.25, device=DEV))
  assert torch.allclose(d.joint_vel[ids][:, jids], torch.full((2, 3), -1.0, device=DEV))
  # site quaternion from rotation matrix is unit and consistent with the body orientation for identity-site frames
  sq = d.site_quat_w
  assert torch.allclose(sq.norm(dim=-1), torch.ones_like(sq[..., 0]), atol=1e-5)


def test_ideal_pd_equals_builtin_position_actuator():
  """tests/test_pd_actuator.py:53-131 of the reference: explicit PD torque == MuJoCo position-actuator force."""
  from mjlab_b200.actuators import ActuatorCmd, IdealPdActuator
  from mjlab_b200.entity import Entity
  sim = _sim("go1", 8)
  m = sim.mj_model
  robot = Entity("robot", sim)
  g = torch.Generator(device=DEV).manual_seed(5)
  n, nj = sim.num_envs, robot.num_joints
  sim.data.qpos[:, 2] += 0.3
  jp = robot.data.default_joint_pos + (torch.rand(n, nj, device=DEV, generator=g) - 0.5) * 0.2
  jv = (torch.rand(n, nj, device=DEV, generator=g) - 0.5) * 2
  robot.write_joint_state_to_sim(jp, jv)
  target = robot.data.default_joint_pos + (torch.rand(n, nj, device=DEV, generator=g) - 0.5) * 0.3
  robot.set_joint_position_target(target)
  robot.write_data_to_sim()
  sim.forward()
  ctrl = robot.indexing.ctrl_ids.cpu().numpy()
  kp = torch.as_tensor(np.asarray(m.actuator_gainprm[ctrl, 0], np.float32), device=DEV)
  kd = -torch.as_tensor(np.asarray(m.actuator_biasprm[ctrl, 2], np.float32), device=DEV)
  lim = torch.as_tensor(np.asarray(m.actuator_forcerange[ctrl, 1], np.float32), device=DEV)
  act = IdealPdActuator(n, nj, 0.0, 0.0, 0.0, device=DEV)
  act.set_gains(slice(None), kp=kp.expand(n, nj), kd=kd.expand(n, nj))
  act.set_effort_limit(slice(None), lim.expand(n, nj))
  z = torch.zeros(n, nj, device=DEV)
  tau = act.compute(ActuatorCmd(target, z, z, robot.data.joint_pos, robot.data.joint_vel))
  torch.testing.assert_close(tau, robot.data.actuator_force, rtol=1e-4, atol=1e-4)


def test_vecenv_wrapper_contract():
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from mjlab_b200.vecenv import RslRlVecEnvWrapper
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=32, seed=3), DEV)
  venv = RslRlVecEnvWrapper(env, clip_actions=0.5)
  assert venv.num_actions == 12 and venv.num_envs == 32 and venv.max_episode_length == 1000
  venv.episode_length_buf = torch.randint_like(venv.episode_length_buf, high=1000)
  ptr = env._buf["episode_length"].data_ptr()
  obs, rew, dones, extras = venv.step(torch.full((32, 12), 3.0, device=DEV))
  assert env._buf["episode_length"].data_ptr() == ptr
  assert obs["policy"].shape == (32, 48) and obs["critic"].shape == (32, 72) and tuple(obs.batch_size) == (32,)
  assert dones.dtype == torch.long and rew.shape == (32,) and "time_outs" in extras
  assert float(env.action_manager.action.abs().max()) <= 0.5 + 1e-6     # clipped before the env sees it
  o2 = venv.get_observations()
  assert torch.equal(o2["policy"], obs["policy"])


def test_observation_pipeline_matches_reference_manager(monkeypatch):
  """clip -> scale -> delay -> history -> concatenate of `obs_pipeline.ObservationPipeline` against the reference's own
  ObservationManager (tests/golden/obs_pipeline.npz, tools/make_obs_pipeline_golden.py) on the same term inputs and the same
  recorded lag draws: per-env stochastic delay with an update period and hold probability, per-term and group-level history
  (term-major, oldest first), partial reset at step 5 (back-fill on the next push), cached compute() without a second push."""
  from pathlib import Path
  from mjlab_b200.obs_pipeline import ObservationPipeline
  g = np.load(Path(__file__).resolve().parent / "golden" / "obs_pipeline.npz")
  seq = {k: torch.as_tensor(g[f"seq_{k}"], device=DEV) for k in ("a", "b", "c")}
  T, B = g["ref_policy"].shape[:2]
  post = {
    "policy": [
      dict(name="a", clip=(-1.0, 1.5), scale=None, delay_min_lag=0, delay_max_lag=0, history_length=0),
      dict(name="b", clip=None, scale=(0.5, 2.0, 1.0, -1.0), delay_min_lag=1, delay_max_lag=3, delay_update_period=2, delay_hold_prob=0.25, history_length=0),
      dict(name="c", clip=(-2.0, 2.0), scale=0.1, delay_min_lag=0, delay_max_lag=2, history_length=3),
    ],
    "critic": [dict(name="a", history_length=2, flatten_history_dim=True), dict(name="c", history_length=2, flatten_history_dim=True)],
  }
  # replay the reference's draws (same hooks as the DelayBuffer tests; the fixture stores them as draw0 .. drawN)
  from mjlab_b200 import actuators as ACT
  draws, pos = [g[f"draw{i}"] for i in range(int(g["ndraws"]))], [0]

  def nxt(shape, device, dtype):
    v = draws[pos[0]]
    pos[0] += 1
    assert tuple(v.shape) == tuple(shape), (pos[0], v.shape, shape)
    return torch.as_tensor(v, device=device).to(dtype)

  monkeypatch.setattr(ACT, "_randint", lambda lo, hi, shape, device, generator=None: nxt(shape, device, torch.long))
  monkeypatch.setattr(ACT, "_rand", lambda shape, device, generator=None: nxt(shape, device, torch.float32))
  pipe = ObservationPipeline({"policy": [("a", 3), ("b", 4), ("c", 2)], "critic": [("a", 3), ("c", 2)]}, post, B, DEV)
  assert [int(np.prod(d)) for d in pipe.term_dims("policy")] == g["term_dims_policy"].tolist()
  assert [int(np.prod(d)) for d in pipe.term_dims("critic")] == g["term_dims_critic"].tolist()
  for t in range(T):
    raw = {"policy": torch.cat([seq["a"][t], seq["b"][t], seq["c"][t]], 1), "critic": torch.cat([seq["a"][t], seq["c"][t]], 1)}
    if t == 5:
      pipe.reset(torch.tensor([1, 3], device=DEV))
    o = pipe.compute(raw, update_history=True)
    again = pipe.compute(raw)
    for grp in ("policy", "critic"):
      np.testing.assert_allclose(o[grp].cpu().numpy(), g[f"ref_{grp}"][t], rtol=1e-6, atol=1e-6, err_msg=f"{grp} t={t}")
      assert again[grp] is o[grp]
  assert pos[0] == len(draws)
