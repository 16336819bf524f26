"""The tracking-env device code (mjlab_b200/csrc/b2_trk.cuh, single-lane host build) against (a) golden vectors from the
reference's own MotionCommand / reward / observation / termination code and (b) the tracking oracle including resets,
adaptive / uniform start-frame sampling, pushes and observation noise with the in-kernel Philox draws."""
import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import envcfg
from mjlab_b200 import trkabi as TA
from oracle import trackstep as OT
from tests.envutil import FIELDS, snapshot_from
from tests.hostemu.harness import HostSim, HostTrk
from tests.test_trk_oracle import setup

TOL = dict(rtol=5e-5, atol=5e-6)
STATE = ["action", "prev_action", "raw_action", "joint_pos_target", "time_steps", "rel_pos", "rel_quat", "metrics", "push_time_left",
         "episode_length", "reward", "step_reward", "episode_sums", "term_dones", "terminated", "time_outs", "reset_buf", "obs_policy",
         "obs_critic", "bin_failed_count", "current_bin_failed"]


def upload(sim, d):
  for f in FIELDS + ["ctrl"]:
    if hasattr(d, f):
      getattr(sim, f)[...] = np.asarray(getattr(d, f)).reshape(getattr(sim, f).shape)


def compare(env, s, names):
  for n in names:
    a, b = getattr(env, n), np.asarray(getattr(s, n))
    if a.dtype.kind == "f":
      np.testing.assert_allclose(a, b.reshape(a.shape), err_msg=n, **TOL)
    else:
      np.testing.assert_array_equal(a.astype(np.int64), b.astype(np.int64).reshape(a.shape), err_msg=n)


def test_kernels_match_reference_golden():
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, cfg = setup(g)
  cfg.max_episode_length = 10**6
  T, B = g["actions"].shape[:2]
  sim = HostSim(m, B, nconmax=35, njmax=250)
  env = HostTrk(sim, cfg, mo)
  env.env_origins[:] = g["origins"]
  env.encoder_bias[:] = g["enc_bias"]
  env.time_steps[:] = g["time0"] + 1
  env.rel_pos[:], env.rel_quat[:] = g["rel0_pos"], g["rel0_quat"]
  alive = np.ones(B, bool)
  w = cfg.reward_weights.astype(np.float32)
  checked = 0
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    env.action_in[:] = g["actions"][t]
    env.stage(0)
    upload(sim, d)
    env.stage(2)
    td = env.term_dones.astype(bool)
    np.testing.assert_array_equal(td[alive, 1], g["ref_term_anchor_pos"][t][alive])
    np.testing.assert_array_equal(td[alive, 2], g["ref_term_anchor_ori"][t][alive])
    np.testing.assert_array_equal(td[alive, 3], g["ref_term_ee"][t][alive])
    np.testing.assert_allclose(env.step_reward[alive], (g["ref_rew_terms"][t] * w)[alive], **TOL)
    np.testing.assert_allclose(env.reward[alive], g["ref_reward"][t][alive], **TOL)
    alive &= ~env.reset_buf.astype(bool)
    upload(sim, d)   # the golden observations come from the un-reset snapshot
    env.stage(3)
    np.testing.assert_allclose(env.metrics[alive], g["ref_metrics"][t][alive], rtol=1e-4, atol=2e-5)
    np.testing.assert_array_equal(env.time_steps[alive], g["ref_time_steps"][t][alive])
    np.testing.assert_allclose(env.rel_pos[alive], g["ref_rel_pos"][t][alive], **TOL)
    np.testing.assert_allclose(env.rel_quat[alive], g["ref_rel_quat"][t][alive], **TOL)
    np.testing.assert_allclose(env.obs_policy[alive], g["ref_obs_policy"][t][alive], **TOL)
    np.testing.assert_allclose(env.obs_critic[alive], g["ref_obs_critic"][t][alive], **TOL)
    checked += int(alive.sum())
  assert checked >= 3 * T


@pytest.mark.parametrize("mode,inject", [("adaptive", False), ("uniform", False), ("adaptive", True), ("start", False)])
def test_kernels_match_oracle_with_resets(mode, inject):
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, cfg = setup(g)
  cfg.enable_corruption, cfg.push_enabled, cfg.sampling_mode = True, True, mode
  cfg.seed, cfg.max_episode_length, cfg.push_interval = 0xABCDEF12345, 5, (0.03, 0.07)
  cfg.adaptive_kernel_size, cfg.adaptive_alpha = 2, 0.3
  cfg.time_step_total = 30      # the clip ends inside the fixture for envs that start late
  cfg.bin_count = 4
  T, B = g["actions"].shape[:2]
  sim = HostSim(m, B, nconmax=35, njmax=250)
  env = HostTrk(sim, cfg, mo, inject=inject)
  s = OT.make_state(cfg, B, g["origins"])
  env.env_origins[:] = g["origins"]
  for o in (env, s):
    o.encoder_bias[:] = g["enc_bias"]
    o.time_steps[:] = np.minimum(g["time0"] + 1, 27)
    o.rel_pos[:], o.rel_quat[:] = g["rel0_pos"], g["rel0_quat"]
    o.episode_length[:] = np.arange(B) % 4
    o.bin_failed_count[:] = [0.5, 0.0, 2.0, 0.25]
  rng = np.random.default_rng(5)
  nreset = 0
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    u = rng.uniform(0, 1, (B, TA.NDRAW)).astype(np.float32) if inject else OT.draws_trk(cfg.seed, B, t)
    if inject:
      env.inject[:] = u
    env.action_in[:] = g["actions"][t]
    env.stage(0)
    OT.apply_action(cfg, s, d, g["actions"][t])
    np.testing.assert_allclose(sim.ctrl, d.ctrl, **TOL)
    upload(sim, d)
    env.stage(2)
    OT.post(cfg, m, mo, s, d, u=u)
    compare(env, s, [n for n in STATE if n not in ("obs_policy", "obs_critic", "metrics", "rel_pos", "rel_quat")])
    for f in ("qpos", "qvel", "ctrl"):
      np.testing.assert_allclose(getattr(sim, f).reshape(getattr(d, f).shape), getattr(d, f), err_msg=f, **TOL)
    assert env.any_reset[0] == int(s.reset_buf.sum())
    nreset += int(s.reset_buf.sum())
    env.stage(3)
    OT.obs_stage(cfg, m, mo, s, d, u=u)
    compare(env, s, STATE)
    np.testing.assert_allclose(sim.qvel, d.qvel, **TOL)
    np.testing.assert_allclose(sim.qpos, d.qpos, **TOL)
    assert int(env.step_counter[0]) == s.step_counter == t + 1
  assert nreset >= B


@pytest.mark.parametrize("full", [3, 0])
def test_3d_termination_variants_kernel_source_vs_reference(full):
  """bad_anchor_pos / bad_motion_body_pos (3-D distance) against the reference's functions; `full = 0`: the z-only variants at the
  same threshold, which give different flags on these states."""
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, cfg = setup(g)
  cfg.max_episode_length = 10**6
  cfg.anchor_pos_threshold = cfg.ee_pos_threshold = 0.05
  cfg.term_full3d = full
  T, B = g["actions"].shape[:2]
  sim = HostSim(m, B, nconmax=35, njmax=250)
  env = HostTrk(sim, cfg, mo)
  env.env_origins[:] = g["origins"]
  env.encoder_bias[:] = g["enc_bias"]
  env.time_steps[:] = g["time0"] + 1
  env.rel_pos[:], env.rel_quat[:] = g["rel0_pos"], g["rel0_quat"]
  alive = np.ones(B, bool)
  checked = 0
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    env.action_in[:] = g["actions"][t]
    env.stage(0)
    upload(sim, d)
    env.stage(2)
    td = env.term_dones.astype(bool)
    np.testing.assert_array_equal(td[alive, 1], g["ref_term_anchor_pos_3d" if full else "ref_term_anchor_pos_z_at_thr3d"][t][alive])
    np.testing.assert_array_equal(td[alive, 3], g["ref_term_ee_3d" if full else "ref_term_ee_z_at_thr3d"][t][alive])
    checked += int(alive.sum())
    alive &= ~env.reset_buf.astype(bool)
    upload(sim, d)
    env.stage(3)
  assert checked >= B
