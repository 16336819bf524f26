"""GPU parity of the fused env step (include/b200env.h) through the C ABI:
(1) kernels vs golden vectors from the reference's own functions, (2) kernels vs the env oracle with resets,
(3) the full closed loop `ManagerBasedRlEnv.step` vs oracle physics + oracle env, (4) full-size properties."""
import ctypes as C

import numpy as np
import pytest
import torch

import mjlab_b200
from mjlab_b200 import envabi as A
from mjlab_b200 import envcfg
from oracle import envstep as OE
from oracle import mjstep as O
from tests.envutil import FIELDS, init_state_from_golden, load_golden, snapshot

pytestmark = pytest.mark.gpu
TOL = dict(rtol=3e-5, atol=3e-6)
DEV = "cuda:0"


def make_env(robot, num_envs, **kw):
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  kw.setdefault("foot_friction_range", None)
  return ManagerBasedRlEnv(VelocityEnvCfg(robot=robot, num_envs=num_envs, **kw), DEV)


class Stages:
  """Drives single stages of the env step through the C ABI on an existing env's buffers."""

  def __init__(self, env):
    self.env, self.lib = env, env._lib
    self.m, self.d, self.e = C.byref(env.sim._cm), C.byref(env.sim._cd), C.byref(env._ce)

  def run(self, k):
    st = torch.cuda.current_stream().cuda_stream
    lib = self.lib
    rc = [lambda: lib.b2_env_action(self.m, self.d, self.e, st), lambda: lib.b2_env_airtime(self.m, self.d, self.e, st),
          lambda: lib.b2_env_post(self.m, self.d, self.e, st),
          lambda: lib.b2_env_obs(self.m, self.d, self.e, 1, st) or lib.b2_env_log(self.d, self.e, st)][k]()
    assert rc == 0, lib.b2_last_error()
    torch.cuda.synchronize()

  def refresh(self):
    """by-value parameters changed on the host side"""
    A.fill_params(self.env._ce, self.env.vcfg)

  def put(self, name, arr):
    t = self.env._buf[name]
    t.copy_(torch.as_tensor(np.asarray(arr)).to(t.dtype).reshape(t.shape))

  def get(self, name):
    return self.env._buf[name].cpu().numpy()

  def upload(self, d):
    data = self.env.sim.data
    for f in FIELDS + ["ctrl"]:
      t = getattr(data, f)
      t.copy_(torch.as_tensor(np.asarray(getattr(d, f), np.float32)).reshape(t.shape))


def _init_from_golden(S, g):
  S.put("command", g["command0"]); S.put("heading_target", g["heading_target"]); S.put("is_heading", g["is_heading"])
  S.put("is_standing", g["is_standing"]); S.put("episode_length", g["episode_len0"])


GENERIC_WEIGHTS = [0.5, -2.0, -1e-4, -1e-3, -1e-6, -0.01, -1.0, -2e-3]   # is_alive, is_terminated, joint_torques_l2, joint_vel_l2, joint_acc_l2, action_acc_l2, flat_orientation_l2, electrical_power_cost


@pytest.mark.parametrize("fixture", ["g1", "go1_fallen"])
def test_generic_terminations_match_reference(fixture):
  """root_height_below_minimum / nan_detection (envs/mdp/terminations.py:35-47) in k_env_post against the reference's functions."""
  robot = fixture.split("_")[0]
  g = load_golden(fixture)
  T, B = g["actions"].shape[:2]
  env = make_env(robot, B)
  m = env.sim.mj_model
  env.vcfg.min_root_height, env.vcfg.nan_termination = float(g["min_root_height"]), True
  S = Stages(env)
  S.refresh()
  _init_from_golden(S, g)
  for t in range(T):
    S.upload(snapshot(g, t, m))
    if t == 3:
      env.sim.data.qacc[1, 2] = float("inf")
    S.run(2)
    td = S.get("term_dones").astype(bool)
    np.testing.assert_array_equal(td[:, 3], g["ref_root_low"][t])
    want_nan = g["ref_nan"][t].copy()
    want_nan[1] |= t == 3
    np.testing.assert_array_equal(td[:, 4], want_nan)
    np.testing.assert_array_equal(S.get("terminated").astype(bool), td[:, 1:].any(1))


@pytest.mark.parametrize("fixture", ["g1", "go1", "g1_fallen", "go1_fallen"])
def test_env_kernels_match_reference_golden(fixture):
  robot = fixture.split("_")[0]
  g = load_golden(fixture)
  T, B = g["actions"].shape[:2]
  env = make_env(robot, B)
  m = env.sim.mj_model
  env.vcfg.enable_corruption = False
  env.vcfg.push_enabled = False
  env.vcfg.reward_weights[14:] = GENERIC_WEIGHTS   # the generic term library is off (weight 0) in the shipped tasks: switch it on
  S = Stages(env)
  S.refresh()
  _init_from_golden(S, g)
  S.put("cmd_time_left", np.full(B, 1e6))
  for n in ("air_state", "air_last_time", "peak_heights", "action", "prev_action", "prev_prev_action", "episode_sums", "cmd_metrics"):
    env._buf[n].zero_()
  alive = np.ones(B, bool)
  checked = ntrue = 0
  sd, tm = env.sim.data.sensordata, env.sim.data.time
  w = env.vcfg.reward_weights.astype(np.float32)
  for t in range(T):
    d = snapshot(g, t, m)
    S.put("action_in", g["actions"][t])
    S.run(0)
    for k in range(4):
      sd.copy_(torch.as_tensor(g["sub_sensordata"][t, k], dtype=torch.float32))
      tm.copy_(torch.as_tensor(g["sub_time"][t, k], dtype=torch.float32).reshape(tm.shape))
      S.run(1)
    np.testing.assert_allclose(S.get("air_state").transpose(1, 0, 2)[:, alive], g["ref_air"][t][:, alive], **TOL)
    S.upload(d)
    S.run(2)
    td = S.get("term_dones").astype(bool)
    np.testing.assert_array_equal(td[alive, 0], g["ref_time_out"][t][alive])
    # bad_orientation / illegal_contact depend on the current state only: compared for EVERY env, reset or not, so the
    # `*_fallen` fixtures pin the True branch of both flags bit-exactly against the reference's own functions
    np.testing.assert_array_equal(td[:, 1], g["ref_fell_over"][t])
    np.testing.assert_array_equal(td[:, 2], g["ref_illegal"][t])
    ntrue += int(td[:, 1:].sum())
    np.testing.assert_allclose(S.get("step_reward")[alive], (g["ref_rew_terms"][t] * w)[alive], **TOL)
    np.testing.assert_allclose(S.get("reward")[alive], (g["ref_rew_terms"][t] * w * np.float32(env.vcfg.step_dt)).sum(1)[alive], rtol=1e-4, atol=1e-5)
    alive &= ~S.get("reset_buf").astype(bool)
    np.testing.assert_allclose(S.get("peak_heights")[alive], g["ref_peak"][t][alive], **TOL)
    S.upload(d)
    S.run(3)
    np.testing.assert_allclose(S.get("command")[alive], g["ref_cmd_after"][t][alive], rtol=3e-5, atol=5e-6)
    np.testing.assert_allclose(S.get("heading_error")[alive], g["ref_heading_error"][t][alive], rtol=3e-5, atol=5e-6)
    np.testing.assert_allclose(S.get("obs_policy")[alive], g["ref_obs_policy"][t][alive], **TOL)
    np.testing.assert_allclose(S.get("obs_critic")[alive], g["ref_obs_critic"][t][alive], **TOL)
    # base_lin_vel / base_ang_vel (envs/mdp/observations.py:24-35) = the reference EntityData's root_link_*_vel_b
    np.testing.assert_allclose(S.get("obs_extra")[alive], np.concatenate([g["ref_root_lin_vel_b"][t], g["ref_root_ang_vel_b"][t]], 1)[alive], **TOL)
    checked += int(alive.sum())
  if fixture.endswith("fallen"):
    assert ntrue >= 20 and checked >= 2 * T
  else:
    assert checked > B * T // 2


STATE = ["action", "prev_action", "prev_prev_action", "raw_action", "joint_pos_target", "command", "heading_target", "cmd_time_left",
         "cmd_metrics", "cmd_counter", "is_heading", "is_standing", "air_state", "air_last_time", "peak_heights", "push_time_left",
         "episode_length", "reward", "step_reward", "episode_sums", "term_dones", "terminated", "time_outs", "reset_buf", "obs_policy",
         "obs_critic", "obs_extra", "log_env"]


def _compare(S, s, names):
  for n in names:
    a, b = S.get(n), np.asarray(getattr(s, n))
    if a.dtype.kind == "f":
      np.testing.assert_allclose(a, b.reshape(a.shape), err_msg=n, **TOL)
    else:
      np.testing.assert_array_equal(a.astype(np.int64), b.astype(np.int64).reshape(a.shape), err_msg=n)


@pytest.mark.parametrize("robot", ["g1", "go1"])
def test_env_kernels_match_oracle_with_resets(robot):
  """In-kernel Philox draws, resets, resampling, pushes, noise: device state == oracle state after every stage."""
  g = load_golden(robot)
  T, B = g["actions"].shape[:2]
  env = make_env(robot, B, seed=0x1234_5678_9ABC)
  m, cfg = env.sim.mj_model, env.vcfg
  cfg.max_episode_length, cfg.cmd_resample, cfg.push_interval = 6, (0.03, 0.09), (0.03, 0.07)
  cfg.reset_joint_pos, cfg.reset_joint_vel = (-0.1, 0.1), (-0.2, 0.2)
  S = Stages(env)
  S.refresh()
  origins = env._buf["env_origins"].cpu().numpy()
  s = OE.make_state(cfg, B, origins)
  _init_from_golden(S, g)
  init_state_from_golden(s, g)
  ep0 = np.arange(B) % 5
  S.put("episode_length", ep0)
  s.episode_length[:] = ep0
  enc = np.random.default_rng(11).uniform(-0.02, 0.02, (B, cfg.nj)).astype(np.float32)
  S.put("encoder_bias", enc)
  s.encoder_bias[:] = enc
  S.put("push_time_left", np.zeros(B))
  nreset = 0
  data = env.sim.data
  for t in range(T):
    d = snapshot(g, t, m)
    u = OE.draws(cfg.seed, B, t)
    S.put("action_in", g["actions"][t])
    S.run(0)
    OE.apply_action(cfg, s, d, g["actions"][t])
    np.testing.assert_allclose(data.ctrl.cpu().numpy(), d.ctrl, **TOL)
    for k in range(4):
      data.sensordata.copy_(torch.as_tensor(g["sub_sensordata"][t, k], dtype=torch.float32))
      data.time.copy_(torch.as_tensor(g["sub_time"][t, k], dtype=torch.float32).reshape(data.time.shape))
      S.run(1)
      OE.update_air_time(cfg, s, g["sub_sensordata"][t, k].astype(np.float32), g["sub_time"][t, k])
    S.upload(d)
    S.run(2)
    OE.post(cfg, m, s, d, u=u)
    _compare(S, s, [n for n in STATE if n not in ("obs_policy", "obs_critic")])
    for f in ("qpos", "qvel", "ctrl", "time"):
      np.testing.assert_allclose(getattr(data, f).cpu().numpy().reshape(getattr(d, f).shape), getattr(d, f), err_msg=f, **TOL)
    assert int(S.get("any_reset")[0]) == int(s.reset_buf.sum())
    nreset += int(s.reset_buf.sum())
    S.run(3)
    OE.obs_stage(cfg, m, s, d, u=u)
    _compare(S, s, STATE)
    np.testing.assert_allclose(data.qvel.cpu().numpy(), d.qvel, **TOL)
    np.testing.assert_allclose(S.get("log"), s.log, rtol=1e-4, atol=1e-5)
    assert int(S.get("step_counter")[0]) == s.step_counter == t + 1
  assert nreset >= B


def _oracle_data(m, B):
  ds = [O.Data(m) for _ in range(B)]
  for x in ds:
    O.forward(m, x, 35, 300)
  return ds


def _batch(ds, m):
  from types import SimpleNamespace as NS
  fields = ["qpos", "qvel", "ctrl", "xpos", "xquat", "subtree_com", "cvel", "site_xpos", "sensordata", "qacc_warmstart", "qacc", "actuator_force"]
  d = NS(**{f: np.stack([np.asarray(getattr(x, f), np.float64) for x in ds]) for f in fields})
  d.time = np.array([x.time for x in ds])
  return d


def _scatter(ds, d):
  for w, x in enumerate(ds):
    x.qpos[:], x.qvel[:], x.ctrl[:], x.qacc_warmstart[:] = d.qpos[w], d.qvel[w], d.ctrl[w], d.qacc_warmstart[w]
    x.time = float(d.time[w])
    x.sensordata[:] = d.sensordata[w]


@pytest.mark.parametrize("robot,mode,generic", [("go1", "reference", False), ("g1", "masked", False), ("go1", "reference", True),
                                                ("go1", "elliptic", False)])
def test_env_step_closed_loop_matches_oracle(robot, mode, generic):
  """env.step() for a few env steps (graph replay, resets included) against oracle physics + oracle env.  `generic`: with the
  generic reward / termination library (envs/mdp/rewards.py, terminations.py) switched on next to the task's own terms.
  "elliptic": the reference forward mode on the elliptic friction cone (`MujocoCfg.cone`)."""
  B, T = 4, 5
  kw = {}
  if mode == "elliptic":
    from mjlab_b200.sim import MujocoCfg, SimulationCfg
    mode = "reference"
    kw["sim"] = SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20, cone="elliptic"))
  env = make_env(robot, B, seed=7, forward_mode=mode, **kw)
  m, cfg = env.sim.mj_model, env.vcfg
  assert m.opt_cone == (1 if kw else 0)
  cfg.max_episode_length = 3
  if generic:
    cfg.reward_weights[14:] = GENERIC_WEIGHTS
    cfg.min_root_height, cfg.nan_termination = 0.262, True    # the open-loop Go1 sags through this height within the roll-out
  A.fill_params(env._ce, cfg)
  env._capture()
  s = OE.make_state(cfg, B, env._buf["env_origins"].cpu().numpy())
  s.push_time_left[:] = env._buf["push_time_left"].cpu().numpy()
  ds = _oracle_data(m, B)
  # reset() on both sides
  obs, _ = env.reset()
  d = _batch(ds, m)
  s.reset_buf[:] = True
  s.log_env[:] = 0
  u = OE.draws(cfg.seed, B, 0)
  OE.reset_idx(cfg, m, s, d, np.arange(B), u)
  _scatter(ds, d)
  for x in ds:
    O.forward(m, x, 35, 300)
  d = _batch(ds, m)
  R = OE.root_kinematics(cfg, d)
  OE.observations(cfg, s, d, R, u)
  s.step_counter += 1       # reset() consumes one draw counter value: the first step() must not reuse its draws
  torch.cuda.synchronize()
  np.testing.assert_allclose(obs["critic"].cpu().numpy(), s.obs_critic, rtol=2e-3, atol=2e-3)
  np.testing.assert_allclose(env.sim.data.qpos.cpu().numpy(), d.qpos, rtol=1e-5, atol=1e-6)
  rng = np.random.default_rng(5)
  for t in range(T):
    a = rng.uniform(-1, 1, (B, cfg.nj)).astype(np.float32)
    obs, rew, term, trunc, extras = env.step(torch.as_tensor(a, device=DEV))
    u = OE.draws(cfg.seed, B, s.step_counter)
    d = _batch(ds, m)
    OE.apply_action(cfg, s, d, a)
    _scatter(ds, d)
    for k in range(4):
      for x in ds:
        O.step(m, x, 35, 300)
      d = _batch(ds, m)
      OE.update_air_time(cfg, s, d.sensordata.astype(np.float32), d.time)
    OE.post(cfg, m, s, d, u=u)
    _scatter(ds, d)
    if s.reset_buf.any():
      for w, x in enumerate(ds):
        if mode == "reference" or s.reset_buf[w]:
          O.forward(m, x, 35, 300)
      d = _batch(ds, m)
    OE.obs_stage(cfg, m, s, d, u=u)
    _scatter(ds, d)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(term.cpu().numpy(), s.terminated)
    np.testing.assert_array_equal(trunc.cpu().numpy(), s.time_outs)
    np.testing.assert_array_equal(env.episode_length_buf.cpu().numpy(), s.episode_length)
    # joint_acc_l2 / joint_torques_l2 square quantities that carry the float32-vs-float64 difference of the contact solve: looser in
    # the closed loop (the golden tests above pin every term on identical inputs)
    np.testing.assert_allclose(rew.cpu().numpy(), s.reward, rtol=4e-2 if generic else 5e-3, atol=1e-3 if generic else 2e-4)
    np.testing.assert_allclose(env.sim.data.qpos.cpu().numpy(), d.qpos, rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(obs["critic"].cpu().numpy(), s.obs_critic, rtol=2e-2, atol=1e-2)
    pol, spol = obs["policy"].cpu().numpy(), s.obs_policy
    np.testing.assert_allclose(pol, spol, rtol=2e-2, atol=1e-2)
    log = extras["log"]
    assert log["Episode/num_resets"] == float(s.reset_buf.sum())
  assert s.step_counter == T + 1 and env.common_step_counter == T
  assert int(env._buf["step_counter"][0]) == T + 1
  if generic:
    assert float(np.abs(s.episode_sums[:, 14:]).sum()) > 0 or bool(s.reset_buf.any())


@pytest.mark.parametrize("robot", ["g1"])
def test_env_full_size_properties(robot):
  """4096 envs (BASELINE config): finite outputs, episode bookkeeping, reset envs return to the init pose, graph replays."""
  B = 4096
  env = make_env(robot, B, seed=1, foot_friction_range=(0.3, 1.2))
  cfg = env.vcfg
  obs, _ = env.reset()
  gen = torch.Generator(device=DEV).manual_seed(0)
  env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (B,), device=DEV, generator=gen)
  total_resets = 0
  for t in range(60):
    a = torch.rand(B, cfg.nj, device=DEV, generator=gen) * 2 - 1
    before = env.episode_length_buf.clone()
    obs, rew, term, trunc, extras = env.step(a)
    done = term | trunc
    assert torch.isfinite(obs["policy"]).all() and torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    ep = env.episode_length_buf
    assert torch.equal(ep[~done], before[~done] + 1) and (ep[done] == 0).all()
    assert torch.equal(trunc, before + 1 >= cfg.max_episode_length)
    if done.any():
      q = env.sim.data.qpos[done]
      assert torch.allclose(q[:, 2], torch.full_like(q[:, 2], float(cfg.default_root_state[2])))
      org = env._buf["env_origins"][done]
      assert ((q[:, 0:2] - org[:, 0:2]).abs() <= 0.5 + 1e-5).all()
      assert torch.allclose(q[:, 3:7].norm(dim=1), torch.ones_like(q[:, 2]), atol=1e-5)
      assert (env.action_manager.action[done] == 0).all()
    total_resets += int(done.sum())
    assert extras["log"]["Episode/num_resets"] == float(done.sum())
  assert total_resets > 100
  fr = env.sim.model.geom_friction
  ids = env.robot.indexing.geom_ids[env.robot.find_geoms(env._foot_geom_names())[0]]
  f0 = fr[:, ids, 0]
  assert f0.min() >= 0.3 and f0.max() <= 1.2 and f0.std() > 0.1


def test_ppo_learns_on_the_env():
  """End-to-end: 30 PPO iterations on Go1 (1024 envs) raise the mean step reward and lengthen episodes."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from mjlab_b200.ppo import OnPolicyRunner, PpoCfg
  from mjlab_b200.vecenv import RslRlVecEnvWrapper
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=1024, seed=1), DEV)
  runner = OnPolicyRunner(RslRlVecEnvWrapper(env), PpoCfg())
  hist = runner.learn(30)
  first = np.mean([h["mean_step_reward"] for h in hist[:3]])
  last = np.mean([h["mean_step_reward"] for h in hist[-3:]])
  assert all(np.isfinite(h["mean_step_reward"]) and np.isfinite(h["kl"]) for h in hist)
  assert last > first, (first, last)


def test_play_zero_agent():
  """BASELINE config 1 (`play --agent zero`, 64 envs).  Go1: the position actuators hold the default pose and the robot
  stays upright (it sags 5 cm under the soft gains until the calves touch the ground, as in the f64 host build).  G1: the shipped gains (knee kp 99 N m/rad
  under a 35 kg robot) do not hold the pose open loop - the robot tips forward within 1.5 s in the oracle as well - so the
  fall is detected, the env resets and the loop keeps running with finite values."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  for robot in ("go1", "g1"):
    env = ManagerBasedRlEnv(VelocityEnvCfg(robot=robot, num_envs=64, play=True, seed=42), DEV)
    assert not env.vcfg.enable_corruption and not env.vcfg.push_enabled and env.max_episode_length > 10**9
    env.reset()
    z0 = env.sim.data.qpos[:, 2].clone()
    zero = torch.zeros(64, env.vcfg.nj, device=DEV)
    fell = torch.zeros(64, dtype=torch.bool, device=DEV)
    for _ in range(300):
      obs, rew, term, trunc, _ = env.step(zero)
      fell |= env.termination_manager.get_term("fell_over")
      assert torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    assert not bool(trunc.any())
    if robot == "go1":
      assert int(fell.sum()) == 0                                # never tips over ...
      assert float(obs["policy"][:, 8].mean()) < -0.95           # ... projected gravity z stays near -1
      z = env.sim.data.qpos[:, 2]
      assert float((z - z0).abs().max()) < 0.08                  # sags < 8 cm (episodes may end on calf-ground contact)
    else:
      assert bool(fell.all())                                    # open-loop G1 tips over ...
      assert int(env.episode_length_buf.max()) < 150             # ... every episode ends and restarts


def test_chunked_streams_do_not_change_results():
  """b2_env_substeps cuts the batch into chunks on side streams (graph edges): the env must produce bit-identical
  trajectories with 1, 2 and 4 chunks, resets and the gated forward pass included."""
  import os
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  outs = []
  old = {k: os.environ.get(k) for k in ("B200SIM_CHUNKS", "B200SIM_CHUNK_MIN")}
  try:
    for chunks in (1, 2, 4):
      os.environ["B200SIM_CHUNKS"], os.environ["B200SIM_CHUNK_MIN"] = str(chunks), "8"
      env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=100, seed=5), DEV)   # 100: a ragged last chunk
      env.reset()
      gen = torch.Generator(device=DEV).manual_seed(3)
      env.episode_length_buf = torch.randint(0, 1000, (100,), device=DEV, generator=gen)
      n0 = env._lib.b2_launch_count()
      for _ in range(25):
        obs, rew, term, trunc, _ = env.step(torch.rand(100, env.vcfg.nj, device=DEV, generator=gen) * 2 - 1)
      assert env.launches_per_step == 1 + chunks * 5 * 4 + 2 + chunks * 4 + 1 + 1
      outs.append((obs["critic"].clone(), rew.clone(), env.sim.data.qpos.clone(), env.sim.data.qvel.clone(), env.sim.data.ncon.clone(),
                   env.episode_length_buf.clone()))
  finally:
    for k, v in old.items():
      if v is None:
        os.environ.pop(k, None)
      else:
        os.environ[k] = v
  for o in outs[1:]:
    for a, b in zip(outs[0], o):
      assert torch.equal(a, b)


@pytest.mark.parametrize("task", ["velocity", "tracking", "rough_go1"])
def test_forward_pass_reuse_is_bit_identical(task):
  """The first sub-step of an env step reuses the position / velocity stage of the forward pass that ended the previous step
  (b2_substeps_after_forward): trajectories with the reuse on (default) and off (B200SIM_MEMO=0) are bit-identical - resets,
  interval pushes (which touch qvel AFTER the forward pass), domain randomisation at reset and a model edit in between
  included - and the reuse really is taken."""
  import os
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  old = os.environ.get("B200SIM_MEMO")
  outs = []
  try:
    for memo in ("0", "1"):
      os.environ["B200SIM_MEMO"] = memo
      if task == "velocity":
        env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=256, seed=5), DEV)
      elif task == "rough_go1":   # geom field, box-box narrowphase, terrain curriculum at reset
        env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", terrain="rough", num_envs=256, seed=5), DEV)
      else:
        env = TrackingRlEnv(TrackingEnvCfg(num_envs=256, seed=5), DEV)
      env.reset()
      gen = torch.Generator(device=DEV).manual_seed(3)
      env.episode_length_buf = torch.randint(0, env.max_episode_length, (256,), device=DEV, generator=gen)
      trace = []
      for k in range(40):
        if k == 20:   # an edit of model fields between two steps: the bridge access drops the memo
          assert not env.sim._memo_dirty
          env.sim.model.geom_friction[:, :, 0] *= 1.0
          assert env.sim._memo_dirty
        if k == 30:   # ... and so does handing out an option tensor through a cached `opt` object
          opt = env.sim.model.opt
          env.sim._memo_dirty = False
          _ = opt.gravity
          assert env.sim._memo_dirty
        if k % 2:     # every other step ends with a time-out, so with the forward pass (it only runs when some env was reset)
          env.episode_length_buf[k % 256] = env.max_episode_length
        obs, rew, term, trunc, _ = env.step(torch.rand(256, env.vcfg.nj, device=DEV, generator=gen) * 2 - 1)
        trace.append((obs["critic"].clone(), rew.clone(), term.clone(), trunc.clone(), env.sim.data.qpos.clone(), env.sim.data.qvel.clone(),
                      env.sim.data.qacc.clone(), env.sim.data.ncon.clone(), env.sim.data.sensordata.clone()))
      outs.append(trace)
      # the flag / snapshot left behind: after a step every world's stage outputs belong to the state in HBM (forward pass over all
      # worlds in the reference mode), unless an interval push changed qvel afterwards
      d = env.sim.data
      same = (d.memo_qpos == d.qpos).all(1) & (d.memo_qvel == d.qvel).all(1) & (d.smooth_memo == 1)
      assert float(same.float().mean()) > 0.9
  finally:
    if old is None:
      os.environ.pop("B200SIM_MEMO", None)
    else:
      os.environ["B200SIM_MEMO"] = old
  for k, (a, b) in enumerate(zip(*outs)):
    for x, y in zip(a, b):
      assert torch.equal(x, y), (k, float((x.float() - y.float()).abs().max()))


def test_fused_post_experiment_knob_still_gives_identical_results():
  """B200SIM_FUSE_POST=1 (sensors + integration inside the solver's tail, measured slower and off by default,
  profiles/r02_fused_post_experiment.txt) must keep producing the default path's trajectories bit for bit."""
  import os
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  old = os.environ.get("B200SIM_FUSE_POST")
  outs = []
  try:
    for knob in ("0", "1"):
      os.environ["B200SIM_FUSE_POST"] = knob
      env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=96, seed=5), DEV)
      env.reset()
      gen = torch.Generator(device=DEV).manual_seed(3)
      env.episode_length_buf = torch.randint(0, 1000, (96,), device=DEV, generator=gen)
      for _ in range(15):
        obs, rew, term, trunc, _ = env.step(torch.rand(96, env.vcfg.nj, device=DEV, generator=gen) * 2 - 1)
      outs.append((obs["critic"].clone(), rew.clone(), env.sim.data.qpos.clone(), env.sim.data.qvel.clone(), env.sim.data.sensordata.clone()))
  finally:
    if old is None:
      os.environ.pop("B200SIM_FUSE_POST", None)
    else:
      os.environ["B200SIM_FUSE_POST"] = old
  for a, b in zip(*outs):
    assert torch.equal(a, b)
