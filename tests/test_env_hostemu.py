"""The env-layer device code (mjlab_b200/csrc/b2_env.cuh, compiled single-lane for the host) against
(a) golden vectors from the reference's own functions and (b) the env oracle including resets, command
resampling, pushes and observation noise with the in-kernel Philox draws."""
import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import envabi as A
from mjlab_b200 import envcfg
from oracle import envstep as OE
from tests.envutil import FIELDS, init_state_from_golden, load_golden, snapshot
from tests.hostemu.harness import HostEnv, HostSim

TOL = dict(rtol=3e-5, atol=3e-6)
STATE = ["action", "prev_action", "prev_prev_action", "raw_action", "joint_pos_target", "command", "heading_target", "cmd_time_left",
         "cmd_metrics", "cmd_counter", "is_heading", "is_standing", "air_state", "air_last_time", "peak_heights", "push_time_left",
         "episode_length", "reward", "step_reward", "episode_sums", "term_dones", "terminated", "time_outs", "reset_buf", "obs_policy",
         "obs_critic", "log_env"]


def upload(sim, d):
  for f in FIELDS + ["ctrl"]:
    getattr(sim, f)[...] = getattr(d, f).reshape(getattr(sim, f).shape)


def compare_state(env, s, names=STATE):
  for n in names:
    a, b = getattr(env, n), getattr(s, n)
    if a.dtype.kind == "f":
      np.testing.assert_allclose(a, b.reshape(a.shape), err_msg=n, **TOL)
    else:
      np.testing.assert_array_equal(a.astype(np.int64), np.asarray(b).astype(np.int64).reshape(a.shape), err_msg=n)


GENERIC_WEIGHTS = [0.5, -2.0, -1e-4, -1e-3, -1e-6, -0.01, -1.0, -2e-3]   # is_alive, is_terminated, joint_torques_l2, joint_vel_l2, joint_acc_l2, action_acc_l2, flat_orientation_l2, electrical_power_cost


@pytest.mark.parametrize("fixture", ["g1", "go1_fallen"])
def test_generic_terminations_match_reference(fixture):
  """root_height_below_minimum / nan_detection (envs/mdp/terminations.py:35-47) against the reference's functions: both depend on
  the current state only, so every env is compared at every step."""
  robot = fixture.split("_")[0]
  m = mjlab_b200.load_model(f"{robot}_flat")
  cfg = envcfg.velocity_cfg(m, robot)
  g = load_golden(fixture)
  cfg.min_root_height, cfg.nan_termination = float(g["min_root_height"]), True
  T, B = g["actions"].shape[:2]
  sim = HostSim(m, B, nconmax=35, njmax=300)
  env = HostEnv(sim, cfg)
  init_state_from_golden(env, g)
  for t in range(T):
    upload(sim, snapshot(g, t, m))
    if t == 3:
      sim.qacc[1, 2] = np.inf     # a world going bad: nan_detection fires for it alone
    env.stage(2)
    np.testing.assert_array_equal(env.term_dones[:, 3], g["ref_root_low"][t])
    want_nan = g["ref_nan"][t].copy()
    want_nan[1] |= t == 3
    np.testing.assert_array_equal(env.term_dones[:, 4], want_nan)
    assert np.array_equal(env.terminated.astype(bool), env.term_dones[:, 1:].any(1))
  assert g["ref_root_low"].any() and not g["ref_root_low"].all()


@pytest.mark.parametrize("fixture", ["g1", "go1", "g1_fallen", "go1_fallen"])
def test_kernels_match_reference_golden(fixture):
  """Before an env's first reset the kernels must reproduce the reference's reward terms / observations / flags.  The
  `*_fallen` fixtures hold states for which the reference's bad_orientation / illegal_contact return True; those two flags
  depend on the current state only, so they are compared for every env at every step, reset or not."""
  robot = fixture.split("_")[0]
  m = mjlab_b200.load_model(f"{robot}_flat")
  cfg = envcfg.velocity_cfg(m, robot)
  cfg.enable_corruption = False
  cfg.push_enabled = False
  cfg.reward_weights[14:] = GENERIC_WEIGHTS   # the generic term library is off (weight 0) in the shipped tasks: switch it on
  g = load_golden(fixture)
  T, B = g["actions"].shape[:2]
  sim = HostSim(m, B, nconmax=35, njmax=300)
  env = HostEnv(sim, cfg)
  init_state_from_golden(env, g)
  env.cmd_time_left[:] = 1e6
  alive = np.ones(B, bool)
  checked = ntrue = 0
  for t in range(T):
    d = snapshot(g, t, m)
    env.action_in[:] = g["actions"][t]
    env.stage(0)
    for k in range(4):
      sim.sensordata[:] = g["sub_sensordata"][t, k]
      sim.time[:] = g["sub_time"][t, k].reshape(sim.time.shape)
      env.stage(1)
    np.testing.assert_allclose(env.air_state.transpose(1, 0, 2)[:, alive], g["ref_air"][t][:, alive], **TOL)
    upload(sim, d)
    env.stage(2)
    np.testing.assert_array_equal(env.term_dones[alive, 0], g["ref_time_out"][t][alive])
    np.testing.assert_array_equal(env.term_dones[:, 1], g["ref_fell_over"][t])
    np.testing.assert_array_equal(env.term_dones[:, 2], g["ref_illegal"][t])
    ntrue += int(env.term_dones[:, 1:].sum())
    w = cfg.reward_weights.astype(np.float32)
    np.testing.assert_allclose(env.step_reward[alive], (g["ref_rew_terms"][t] * w)[alive], **TOL)
    np.testing.assert_allclose(env.reward[alive], (g["ref_rew_terms"][t] * w * np.float32(cfg.step_dt)).sum(1)[alive], rtol=1e-4, atol=1e-5)
    alive &= ~env.reset_buf.astype(bool)
    np.testing.assert_allclose(env.peak_heights[alive], g["ref_peak"][t][alive], **TOL)
    upload(sim, d)  # the reset rewrote qpos/qvel of done envs; the golden observations come from the un-reset snapshot
    env.stage(3)
    np.testing.assert_allclose(env.command[alive], g["ref_cmd_after"][t][alive], rtol=3e-5, atol=5e-6)
    np.testing.assert_allclose(env.heading_error[alive], g["ref_heading_error"][t][alive], rtol=3e-5, atol=5e-6)
    np.testing.assert_allclose(env.obs_policy[alive], g["ref_obs_policy"][t][alive], **TOL)
    np.testing.assert_allclose(env.obs_critic[alive], g["ref_obs_critic"][t][alive], **TOL)
    # base_lin_vel / base_ang_vel (envs/mdp/observations.py:24-35) = the reference EntityData's root_link_*_vel_b
    np.testing.assert_allclose(env.obs_extra[alive], np.concatenate([g["ref_root_lin_vel_b"][t], g["ref_root_ang_vel_b"][t]], 1)[alive], **TOL)
    checked += int(alive.sum())
  if fixture.endswith("fallen"):
    assert ntrue >= 20 and checked >= 2 * T   # True flags seen; the two upright envs stay comparable throughout
  else:
    assert checked > B * T // 2


@pytest.mark.parametrize("robot,inject", [("g1", False), ("go1", False), ("g1", True)])
def test_kernels_match_oracle_with_resets(robot, inject):
  """Full stage sequence incl. resets / resampling / push / noise: kernel state == oracle state after every stage."""
  m = mjlab_b200.load_model(f"{robot}_flat")
  cfg = envcfg.velocity_cfg(m, robot)
  cfg.seed = 0x1234_5678_9ABC
  cfg.max_episode_length = 6       # several time-outs inside the fixture
  cfg.cmd_resample = (0.03, 0.09)  # timed resamples every 2-5 env steps
  cfg.push_interval = (0.03, 0.07)
  cfg.reset_joint_pos = (-0.1, 0.1)
  cfg.reset_joint_vel = (-0.2, 0.2)
  g = load_golden(robot)
  T, B = g["actions"].shape[:2]
  origins = np.random.default_rng(3).uniform(-4, 4, (B, 3)).astype(np.float32)
  sim = HostSim(m, B, nconmax=35, njmax=300)
  env = HostEnv(sim, cfg, inject=inject)
  s = OE.make_state(cfg, B, origins)
  env.env_origins[:] = origins
  init_state_from_golden(env, g)
  init_state_from_golden(s, g)
  s.episode_length[:] = env.episode_length[:] = np.arange(B) % 5
  rng = np.random.default_rng(11)
  enc = rng.uniform(-0.02, 0.02, (B, cfg.nj)).astype(np.float32)
  env.encoder_bias[:] = enc
  s.encoder_bias[:] = enc
  nreset = npush = 0
  for t in range(T):
    d = snapshot(g, t, m)
    u = rng.uniform(0, 1, (B, A.NDRAW)).astype(np.float32) if inject else OE.draws(cfg.seed, B, t)
    if inject:
      env.inject[:] = u
    env.action_in[:] = g["actions"][t]
    env.stage(0)
    OE.apply_action(cfg, s, d, g["actions"][t])
    np.testing.assert_allclose(sim.ctrl, d.ctrl, **TOL)
    for k in range(4):
      sim.sensordata[:] = g["sub_sensordata"][t, k]
      sim.time[:] = g["sub_time"][t, k].reshape(sim.time.shape)
      env.stage(1)
      OE.update_air_time(cfg, s, g["sub_sensordata"][t, k].astype(np.float32), g["sub_time"][t, k])
    upload(sim, d)
    env.stage(2)
    OE.post(cfg, m, s, d, u=u)
    compare_state(env, s, [n for n in STATE if n not in ("obs_policy", "obs_critic")])
    for f in ("qpos", "qvel", "ctrl", "time"):
      np.testing.assert_allclose(getattr(sim, f).reshape(getattr(d, f).shape), getattr(d, f), err_msg=f, **TOL)
    assert env.any_reset[0] == int(s.reset_buf.sum())
    nreset += int(s.reset_buf.sum())
    before = d.qvel.copy()
    env.stage(3)
    OE.obs_stage(cfg, m, s, d, u=u)
    npush += int(np.any(before != d.qvel, axis=1).sum())
    compare_state(env, s)
    np.testing.assert_allclose(sim.qvel, d.qvel, **TOL)
    np.testing.assert_allclose(env.log, s.log, rtol=1e-4, atol=1e-5)
    assert int(env.step_counter[0]) == s.step_counter == t + 1
  assert nreset >= B and npush > 0


@pytest.mark.parametrize("mode", [1, 3])
def test_terrain_curriculum_matches_oracle(mode):
  """Rough-terrain reset path: terrain_levels_vel + update_env_origins (mode bit 0) and the play-mode randomize_terrain
  event (bit 1) inside the reset kernel == the env oracle, level by level, including the random restart past the top row."""
  m = mjlab_b200.load_model("g1_flat")
  cfg = envcfg.velocity_cfg(m, "g1")
  cfg.seed = 0xBEEF
  cfg.max_episode_length = 4
  rng = np.random.default_rng(5)
  rows, cols = 4, 3
  cfg.terrain_mode, cfg.terrain_half_size = mode, 4.0
  cfg.terrain_origins = rng.uniform(-6.0, 6.0, (rows, cols, 3))      # roots sit near (0, 0): distances 0 .. 8.5 m
  cfg.terrain_origins[..., 2] = rng.uniform(-0.5, 0.5, (rows, cols))
  g = load_golden("g1")
  T, B = g["actions"].shape[:2]
  sim = HostSim(m, B, nconmax=35, njmax=300)
  env = HostEnv(sim, cfg)
  s = OE.make_state(cfg, B)
  lv, ty = rng.integers(0, rows, B).astype(np.int32), rng.integers(0, cols, B).astype(np.int32)
  for st in (env, s):
    init_state_from_golden(st, g)
    st.terrain_level[:], st.terrain_type[:] = lv, ty
    st.env_origins[:] = cfg.terrain_origins[lv, ty].astype(np.float32)
    st.episode_length[:] = np.arange(B) % 4
  moved = wrapped = 0
  for t in range(T):
    d = snapshot(g, t, m)
    u = OE.draws(cfg.seed, B, t)
    env.action_in[:] = g["actions"][t]
    env.stage(0)
    OE.apply_action(cfg, s, d, g["actions"][t])
    upload(sim, d)
    before = s.terrain_level.copy()
    env.stage(2)
    OE.post(cfg, m, s, d, u=u)
    compare_state(env, s, ["terrain_level", "terrain_type", "env_origins", "reset_buf", "log_env"])
    for f in ("qpos", "qvel"):
      np.testing.assert_allclose(getattr(sim, f).reshape(getattr(d, f).shape), getattr(d, f), err_msg=f, **TOL)
    moved += int((before != s.terrain_level).sum())
    wrapped += int(((before == rows - 1) & s.reset_buf & (s.terrain_level != rows - 1)).sum())
    env.stage(3)
    OE.obs_stage(cfg, m, s, d, u=u)
    np.testing.assert_allclose(env.log, s.log, rtol=1e-4, atol=1e-5)
    assert np.isclose(env.log[A.K["B2E_LOG_TERRAIN_LEVEL"]], s.terrain_level.sum())
  assert moved > B // 2 and (mode == 3 or wrapped > 0)
  # reset poses sit on the (possibly updated) origins
  assert np.all((s.terrain_level >= 0) & (s.terrain_level < rows)) and np.all((s.terrain_type >= 0) & (s.terrain_type < cols))


def test_reward_manager_zeroes_nan_and_inf_terms():
  """RewardManager.compute turns a NaN / +-Inf term VALUE into 0 for that env and term, the other terms still count
  (managers/reward_manager.py:77-91; the reference's tests/test_rewards.py:235-273)."""
  m = mjlab_b200.load_model("g1_flat")
  cfg = envcfg.velocity_cfg(m, "g1")
  cfg.push_enabled = False
  g = load_golden("g1")
  B = g["actions"].shape[1]
  k = envcfg.REWARD_NAMES.index("angular_momentum")
  rew, step = [], []
  for poison in (False, True):
    sim = HostSim(m, B, nconmax=35, njmax=300)
    env = HostEnv(sim, cfg)
    init_state_from_golden(env, g)
    env.action_in[:] = g["actions"][0]
    env.stage(0)
    upload(sim, snapshot(g, 0, m))
    if poison:
      sim.sensordata[1, cfg.angmom_adr] = np.nan
      sim.sensordata[3, cfg.angmom_adr] = np.inf
      sim.sensordata[4, cfg.angmom_adr] = -np.inf     # squared: +inf
    env.stage(2)
    rew.append(env.reward.copy()); step.append(env.step_reward.copy())
  clean_r, bad_r = rew
  clean_s, bad_s = step
  assert np.all(np.isfinite(bad_r)) and np.all(np.isfinite(bad_s))
  for w in (1, 3, 4):
    assert bad_s[w, k] == 0.0 and clean_s[w, k] != 0.0
    np.testing.assert_allclose(bad_r[w], clean_r[w] - clean_s[w, k] * np.float32(cfg.step_dt), rtol=1e-5, atol=1e-7)
  others = [w for w in range(B) if w not in (1, 3, 4)]
  np.testing.assert_array_equal(bad_r[others], clean_r[others])


def test_action_history_tracking_and_reset():
  """ActionManager: action / prev_action / prev_prev_action shift on every process_action and are zeroed for the envs that reset
  (managers/action_manager.py:100-120; the reference's tests/test_action_manager.py:51-118), here in k_env_action / the reset inside
  k_env_post."""
  m = mjlab_b200.load_model("go1_flat")
  cfg = envcfg.velocity_cfg(m, "go1")
  cfg.push_enabled = False
  B = 4
  sim = HostSim(m, B, nconmax=35, njmax=300)
  env = HostEnv(sim, cfg)
  sim.forward()
  assert not env.action.any() and not env.prev_action.any() and not env.prev_prev_action.any()
  acts = [np.full((B, cfg.nj), 0.1 * i, np.float32) for i in range(1, 5)]
  for t, a in enumerate(acts):
    env.action_in[:] = a
    env.stage(0)
    np.testing.assert_array_equal(env.action, a)
    np.testing.assert_array_equal(env.prev_action, acts[t - 1] if t >= 1 else 0 * a)
    np.testing.assert_array_equal(env.prev_prev_action, acts[t - 2] if t >= 2 else 0 * a)
  # envs 0 and 2 time out in this step's post stage -> their history is cleared, the others keep theirs
  env.episode_length[:] = 0
  env.episode_length[[0, 2]] = cfg.max_episode_length
  env.stage(2)
  assert env.reset_buf[[0, 2]].all() and not env.reset_buf[[1, 3]].any()
  for w in (0, 2):
    assert not env.action[w].any() and not env.prev_action[w].any() and not env.prev_prev_action[w].any()
  for w in (1, 3):
    np.testing.assert_array_equal(env.action[w], acts[3][w])
    np.testing.assert_array_equal(env.prev_action[w], acts[2][w])
    np.testing.assert_array_equal(env.prev_prev_action[w], acts[1][w])
