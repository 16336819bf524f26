"""GPU parity of the tracking-task env step through the C ABI (include/b200trk.h): kernels vs golden vectors from the
reference's own MotionCommand / rewards / observations / terminations, kernels vs the oracle with resets, full-size run."""
import ctypes as C

import numpy as np
import pytest
import torch

from mjlab_b200 import trkabi as TA
from oracle import trackstep as OT
from tests.envutil import FIELDS, snapshot_from
from tests.test_trk_oracle import setup

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = dict(rtol=5e-5, atol=5e-6)


class Drv:
  def __init__(self, env):
    self.env, self.lib = env, env._lib
    self.m, self.d, self.e = C.byref(env.sim._cm), C.byref(env.sim._cd), C.byref(env._ce)

  def run(self, k):
    st = torch.cuda.current_stream().cuda_stream
    lib = self.lib
    rc = {0: lambda: lib.b2_trk_action(self.m, self.d, self.e, st), 2: lambda: lib.b2_trk_post(self.m, self.d, self.e, st),
          3: lambda: lib.b2_trk_obs(self.m, self.d, self.e, 1, st) or lib.b2_trk_log(self.d, self.e, st)}[k]()
    assert rc == 0, lib.b2_last_error()
    torch.cuda.synchronize()

  def put(self, name, arr):
    t = self.env._buf[name]
    t.copy_(torch.as_tensor(np.asarray(arr)).to(t.dtype).reshape(t.shape))

  def get(self, name):
    return self.env._buf[name].cpu().numpy()

  def upload(self, d):
    data = self.env.sim.data
    for f in FIELDS + ["ctrl"]:
      if hasattr(d, f) and f in data:
        t = getattr(data, f)
        t.copy_(torch.as_tensor(np.asarray(getattr(d, f), np.float32)).reshape(t.shape))


def make_env(g, mo, cfg_edit=None, **kw):
  import tempfile
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  f = tempfile.NamedTemporaryFile(suffix=".npz", delete=False)
  np.savez(f, fps=np.array([50.0]), **mo)
  f.close()
  return TrackingRlEnv(TrackingEnvCfg(num_envs=g["actions"].shape[1], motion_file=f.name, domain_randomization=False, **kw), DEV, vcfg_edit=cfg_edit)


def test_trk_kernels_match_reference_golden():
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, _ = setup(g)

  def edit(c):
    c.enable_corruption, c.push_enabled, c.sampling_mode, c.max_episode_length = False, False, "start", 10**6

  env = make_env(g, mo, edit)
  S = Drv(env)
  T, B = g["actions"].shape[:2]
  S.put("env_origins", g["origins"]); S.put("encoder_bias", g["enc_bias"]); S.put("time_steps", g["time0"] + 1)
  S.put("rel_pos", g["rel0_pos"]); S.put("rel_quat", g["rel0_quat"])
  alive = np.ones(B, bool)
  w = env.vcfg.reward_weights.astype(np.float32)
  checked = 0
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    S.put("action_in", g["actions"][t])
    S.run(0)
    S.upload(d)
    S.run(2)
    td = S.get("term_dones").astype(bool)
    np.testing.assert_array_equal(td[alive, 1], g["ref_term_anchor_pos"][t][alive])
    np.testing.assert_array_equal(td[alive, 2], g["ref_term_anchor_ori"][t][alive])
    np.testing.assert_array_equal(td[alive, 3], g["ref_term_ee"][t][alive])
    np.testing.assert_allclose(S.get("step_reward")[alive], (g["ref_rew_terms"][t] * w)[alive], **TOL)
    np.testing.assert_allclose(S.get("reward")[alive], g["ref_reward"][t][alive], **TOL)
    alive &= ~S.get("reset_buf").astype(bool)
    S.upload(d)
    S.run(3)
    np.testing.assert_allclose(S.get("metrics")[alive], g["ref_metrics"][t][alive], rtol=1e-4, atol=2e-5)
    np.testing.assert_array_equal(S.get("time_steps")[alive], g["ref_time_steps"][t][alive])
    np.testing.assert_allclose(S.get("rel_pos")[alive], g["ref_rel_pos"][t][alive], **TOL)
    np.testing.assert_allclose(S.get("rel_quat")[alive], g["ref_rel_quat"][t][alive], **TOL)
    np.testing.assert_allclose(S.get("obs_policy")[alive], g["ref_obs_policy"][t][alive], **TOL)
    np.testing.assert_allclose(S.get("obs_critic")[alive], g["ref_obs_critic"][t][alive], **TOL)
    checked += int(alive.sum())
  assert checked >= 3 * T


STATE = ["action", "prev_action", "raw_action", "joint_pos_target", "time_steps", "rel_pos", "rel_quat", "metrics", "push_time_left",
         "episode_length", "reward", "step_reward", "episode_sums", "term_dones", "terminated", "time_outs", "reset_buf", "obs_policy",
         "obs_critic", "bin_failed_count", "current_bin_failed"]


@pytest.mark.parametrize("mode", ["adaptive", "uniform"])
def test_trk_kernels_match_oracle_with_resets(mode):
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, _ = setup(g)

  def edit(c):
    c.enable_corruption, c.push_enabled, c.sampling_mode = True, True, mode
    c.seed, c.max_episode_length, c.push_interval = 0xABCDEF12345, 5, (0.03, 0.07)
    c.adaptive_kernel_size, c.adaptive_alpha, c.time_step_total, c.bin_count = 2, 0.3, 30, 4

  env = make_env(g, mo, edit)
  cfg = env.vcfg
  S = Drv(env)
  T, B = g["actions"].shape[:2]
  s = OT.make_state(cfg, B, g["origins"])
  init = dict(env_origins=g["origins"], encoder_bias=g["enc_bias"], time_steps=np.minimum(g["time0"] + 1, 27), rel_pos=g["rel0_pos"],
              rel_quat=g["rel0_quat"], episode_length=np.arange(B) % 4, bin_failed_count=np.array([0.5, 0.0, 2.0, 0.25]), push_time_left=np.zeros(B))
  for k, v in init.items():
    S.put(k, v)
    getattr(s, k)[:] = v
  data = env.sim.data
  nreset = 0
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    u = OT.draws_trk(cfg.seed, B, t)
    S.put("action_in", g["actions"][t])
    S.run(0)
    OT.apply_action(cfg, s, d, g["actions"][t])
    S.upload(d)
    S.run(2)
    OT.post(cfg, m, mo, s, d, u=u)
    for n in [n for n in STATE if n not in ("obs_policy", "obs_critic", "metrics", "rel_pos", "rel_quat")]:
      a, b = S.get(n), np.asarray(getattr(s, n))
      if a.dtype.kind == "f":
        np.testing.assert_allclose(a, b.reshape(a.shape), err_msg=n, **TOL)
      else:
        np.testing.assert_array_equal(a.astype(np.int64), b.astype(np.int64).reshape(a.shape), err_msg=n)
    for f in ("qpos", "qvel"):
      np.testing.assert_allclose(getattr(data, f).cpu().numpy(), getattr(d, f), err_msg=f, **TOL)
    nreset += int(s.reset_buf.sum())
    S.run(3)
    OT.obs_stage(cfg, m, mo, s, d, u=u)
    for n in STATE:
      a, b = S.get(n), np.asarray(getattr(s, n))
      if a.dtype.kind == "f":
        np.testing.assert_allclose(a, b.reshape(a.shape), err_msg=n, **TOL)
      else:
        np.testing.assert_array_equal(a.astype(np.int64), b.astype(np.int64).reshape(a.shape), err_msg=n)
  assert nreset >= B


def test_tracking_env_full_size():
  """BASELINE config 5: 4096 envs, synthetic clip, domain randomisation on; finite outputs, frames advance, resets resample."""
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  B = 4096
  env = TrackingRlEnv(TrackingEnvCfg(num_envs=B, seed=3), DEV)
  cfg = env.vcfg
  assert cfg.n_obs_policy == 160 and cfg.n_obs_critic == 286 and cfg.time_step_total == 500 and cfg.bin_count == 11
  obs, _ = env.reset()
  gen = torch.Generator(device=DEV).manual_seed(0)
  total = 0
  for t in range(40):
    before = env._buf["time_steps"].clone()
    obs, rew, term, trunc, extras = env.step((torch.rand(B, cfg.nj, device=DEV, generator=gen) * 2 - 1) * 0.2)
    done = term | trunc
    assert torch.isfinite(obs["policy"]).all() and torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    ts = env._buf["time_steps"]
    wrapped = before + 1 >= cfg.time_step_total       # end of the clip: the command resamples its start frame
    ok = ~done & ~wrapped
    assert torch.equal(ts[ok], before[ok] + 1) and (ts >= 0).all() and (ts < cfg.time_step_total).all()
    assert (env.episode_length_buf[done] == 0).all()
    total += int(done.sum())
    assert extras["log"]["Episode/num_resets"] == float(done.sum())
  assert total > B                     # the very first step terminates every env (zero-initialised relative targets, as in the reference)
  assert float(env._buf["bin_failed_count"].sum()) > 0
  # the command exposed to user code is the clip's joint state at the current frame
  cmd = env.command_manager.get_command("motion")
  assert cmd.shape == (B, 58)
  torch.testing.assert_close(obs["critic"][:, :58], cmd)
  ipos = env.sim.model.body_ipos
  tid = env.robot.indexing.body_ids[env.robot.find_bodies("torso_link")[0]]
  assert float(ipos[:, tid, 0].std()) > 0.005 and float(env._buf["encoder_bias"].abs().max()) <= 0.01 + 1e-6


def test_tracking_env_from_the_reference_manager_cfg_equals_the_builtin_task():
  """`TrackingEnvCfg.from_manager_cfg(load_env_cfg("Mjlab-Tracking-Flat-Unitree-G1"))` (the JSON dump of the reference's own cfg
  object, tests/golden/task_cfg) steps bit-identically to the built-in tracking task, startup events included (same torch seed)."""
  import json
  from pathlib import Path
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  d = json.loads((Path(__file__).resolve().parent / "golden" / "task_cfg" / "Mjlab-Tracking-Flat-Unitree-G1.json").read_text())
  B = 48
  outs = []
  for cfg in (TrackingEnvCfg(num_envs=B, seed=3), TrackingEnvCfg.from_manager_cfg(d, num_envs=B, seed=3)):
    torch.manual_seed(11)
    env = TrackingRlEnv(cfg, DEV)
    g = torch.Generator(device=DEV).manual_seed(2)
    env.reset()
    for _ in range(12):
      obs, rew, term, trunc, _ = env.step(torch.rand(B, env.vcfg.nj, device=DEV, generator=g) * 2 - 1)
    outs.append((obs["policy"].clone(), obs["critic"].clone(), rew.clone(), term.clone(), trunc.clone(), env.sim.data.qpos.clone(),
                 env.sim.model.body_ipos.clone(), env.sim.model.geom_friction.clone(), env._buf["encoder_bias"].clone()))
    assert env.max_episode_length == 500
  for x, y in zip(*outs):
    assert torch.equal(x, y)
  # an edited cfg: one termination removed, a reward re-weighted
  d["terminations"].pop("anchor_ori")
  d["rewards"]["motion_body_pos"]["weight"] = 2.0
  env = TrackingRlEnv(TrackingEnvCfg.from_manager_cfg(d, num_envs=8, seed=3), DEV)
  assert env.vcfg.anchor_ori_threshold > 1e8 and env.vcfg.reward_weights[2] == 2.0
  env.reset()
  env.step(torch.zeros(8, env.vcfg.nj, device=DEV))


def test_tracking_no_state_estimation_task():
  """The reference's registered `Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation` (its cfg dump): the same physics, rewards and
  resets as the full task (same seeds), a critic group equal to the full task's, and a policy group that is the full task's
  policy group without the motion_anchor_pos_b and base_lin_vel columns - the noisy columns bit for bit, the noise-free ones
  (command, last action) from the clean copy."""
  import json
  from pathlib import Path
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  cfgdir = Path(__file__).resolve().parent / "golden" / "task_cfg"
  full = json.loads((cfgdir / "Mjlab-Tracking-Flat-Unitree-G1.json").read_text())
  nse = json.loads((cfgdir / "Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation.json").read_text())
  B = 32
  outs = []
  for d in (full, nse):
    torch.manual_seed(11)
    env = TrackingRlEnv(TrackingEnvCfg.from_manager_cfg(d, num_envs=B, seed=3), DEV)
    g = torch.Generator(device=DEV).manual_seed(2)
    obs, _ = env.reset()
    hist = [(obs["policy"].clone(), obs["critic"].clone())]
    for _ in range(8):
      obs, rew, term, trunc, _ = env.step(torch.rand(B, env.vcfg.nj, device=DEV, generator=g) * 2 - 1)
      hist.append((obs["policy"].clone(), obs["critic"].clone(), rew.clone(), term.clone(), trunc.clone()))
    outs.append((env, hist))
  (e_full, h_full), (e_nse, h_nse) = outs
  nj = e_full.vcfg.nj
  assert e_nse.observation_manager.active_terms["policy"] == ["command", "motion_anchor_ori_b", "base_ang_vel", "joint_pos", "joint_vel", "actions"]
  assert e_nse.observation_manager.group_obs_dim["policy"] == (e_full.vcfg.n_obs_policy - 6,)
  keep = torch.ones(e_full.vcfg.n_obs_policy, dtype=torch.bool, device=DEV)
  keep[2 * nj: 2 * nj + 3] = False          # motion_anchor_pos_b
  keep[2 * nj + 9: 2 * nj + 12] = False     # base_lin_vel
  noisy = keep.clone()
  noisy[: 2 * nj] = False                   # command: no noise -> clean copy (identical values)
  noisy[4 * nj + 15:] = False               # last action: no noise
  for a, b in zip(h_full, h_nse):
    assert torch.equal(a[1], b[1])                                   # critic group: identical
    assert b[0].shape == (B, int(keep.sum()))
    sel = noisy[keep]
    # same physics and same Philox draws; the dropped terms' noise draws do not shift the others' (draws are keyed per column)
    torch.testing.assert_close(b[0][:, sel], a[0][:, keep][:, sel], rtol=0, atol=0)
    torch.testing.assert_close(b[0][:, ~sel], a[0][:, keep][:, ~sel], rtol=0, atol=0)
    for x, y in zip(a[2:], b[2:]):
      assert torch.equal(x, y)                                       # rewards, terminations, truncations


def test_builtin_no_state_estimation_equals_the_reference_cfg_variant():
  """`TrackingEnvCfg(has_state_estimation=False)` (built-in table) steps bit-identically to the env built from the dump of the
  reference's own `Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation` cfg."""
  import json
  from pathlib import Path
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  d = json.loads((Path(__file__).resolve().parent / "golden" / "task_cfg" / "Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation.json").read_text())
  B = 24
  outs = []
  for cfg in (TrackingEnvCfg(num_envs=B, seed=5, has_state_estimation=False), TrackingEnvCfg.from_manager_cfg(d, num_envs=B, seed=5)):
    torch.manual_seed(7)
    env = TrackingRlEnv(cfg, DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    obs, _ = env.reset()
    for _ in range(6):
      obs, rew, term, trunc, _ = env.step(torch.rand(B, env.vcfg.nj, device=DEV, generator=g) * 2 - 1)
    outs.append((obs["policy"].clone(), obs["critic"].clone(), rew.clone(), term.clone(), trunc.clone()))
    assert obs["policy"].shape == (B, env.vcfg.n_obs_policy - 6)
  for x, y in zip(*outs):
    assert torch.equal(x, y)


@pytest.mark.parametrize("full", [3, 0])
def test_trk_3d_termination_variants(full):
  """bad_anchor_pos / bad_motion_body_pos (3-D distance; tasks/tracking/mdp/terminations.py:18-25,55-70) on the GPU against the
  reference's own functions; `full = 0`: the height-only variants at the same threshold (different flags on these states)."""
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, _ = setup(g)

  def edit(c):
    c.enable_corruption, c.push_enabled, c.sampling_mode, c.max_episode_length = False, False, "start", 10**6
    c.anchor_pos_threshold = c.ee_pos_threshold = 0.05
    c.term_full3d = full

  env = make_env(g, mo, edit)
  S = Drv(env)
  T, B = g["actions"].shape[:2]
  S.put("env_origins", g["origins"]); S.put("encoder_bias", g["enc_bias"]); S.put("time_steps", g["time0"] + 1)
  S.put("rel_pos", g["rel0_pos"]); S.put("rel_quat", g["rel0_quat"])
  alive = np.ones(B, bool)
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    S.put("action_in", g["actions"][t])
    S.run(0)
    S.upload(d)
    S.run(2)
    td = S.get("term_dones").astype(bool)
    np.testing.assert_array_equal(td[alive, 1], g["ref_term_anchor_pos_3d" if full else "ref_term_anchor_pos_z_at_thr3d"][t][alive])
    np.testing.assert_array_equal(td[alive, 3], g["ref_term_ee_3d" if full else "ref_term_ee_z_at_thr3d"][t][alive])
    alive &= ~S.get("reset_buf").astype(bool)
    S.upload(d)
    S.run(3)
