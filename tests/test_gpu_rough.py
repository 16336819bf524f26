"""GPU: rough terrain (BASELINE config 4, Mjlab-Velocity-Rough-Unitree-G1) - the CUDA path on the 3564-box terrain grid
against the float64 oracle (contact lists bit-exact, state within the per-step tolerance), and the full-size env
(8192 envs, curriculum on) through size-independent properties."""
import numpy as np
import pytest
import torch

import mjlab_b200
from oracle import mjstep as O
from tests.test_gpu_parity import load_state, make_sim, relerr
from tests.test_rough_cpu import rough_states

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_rough_physics_matches_oracle():
  m = mjlab_b200.load_model("g1_rough")
  nw, nstep = 8, 3
  qpos, qvel, ctrl, xfrc = rough_states(m, nw, seed=21)
  sim = make_sim(m, nw, nconmax=64, export_efc_J=False)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  ds = []
  for w in range(nw):
    d = O.Data(m)
    d.qpos[:], d.qvel[:], d.ctrl[:] = qpos[w], qvel[w], ctrl[w]
    ds.append(d)
  field = set(m.field_geom.tolist())
  box_contacts = 0
  verr = []
  D = sim.data
  for s in range(nstep):
    sim.step()
    torch.cuda.synchronize()
    for w, d in enumerate(ds):
      O.step(m, d, 64, 300)
      n = len(d.contacts)
      assert int(D.ncon[w]) == n, (s, w)
      assert D.contact_geom[w, :n].cpu().tolist() == [list(c.geom) for c in d.contacts]     # bit-exact pair lists
      assert int(D.nefc[w]) == d.nefc
      if n:
        # the terrain spans +-100 m: one float32 ulp of a world coordinate there is 7.6e-6 m, which bounds every
        # distance computed from world positions (the flat-scene tests, near the origin, hold 2e-6)
        np.testing.assert_allclose(D.contact_dist[w, :n].cpu().numpy(), [c.dist for c in d.contacts], rtol=0, atol=3e-5)
        # where a capsule lies almost parallel to a box face the closest point slides along the segment under rounding
        # (its depth does not): positions agree to a fraction of a millimetre, depths to the ulp bound above
        np.testing.assert_allclose(D.contact_pos[w, :n].cpu().numpy(), [c.pos for c in d.contacts], rtol=0, atol=5e-4)
      box_contacts += sum(1 for c in d.contacts if c.geom[1] in field)
      assert relerr(D.qpos[w].cpu().numpy(), d.qpos) < 1e-4            # north_star: ~1e-4 relative per step
      verr.append(relerr(D.qvel[w].cpu().numpy(), d.qvel))
  # velocities: these states start with feet up to 5 cm inside step edges, so contact forces are hundreds of newtons and
  # the sub-millimetre slide of a contact point (above) moves joint torques visibly; typical worlds agree to 2e-3
  assert np.median(verr) < 2e-3 and max(verr) < 5e-2, (np.median(verr), max(verr))
  assert box_contacts >= 20
  # static terrain geoms hold the model constant; moving geoms follow the bodies
  np.testing.assert_allclose(D.geom_xpos[0, m.field_geom].cpu().numpy(), m.sgeom_xpos[m.field_geom], atol=1e-5)
  np.testing.assert_allclose(D.geom_xpos[0].cpu().numpy(), ds[0].geom_xpos, atol=5e-5)


def test_rough_env_closed_loop_small():
  """64 envs on the terrain: reset poses sit on sub-terrain origins, curriculum levels move, values stay finite."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  B = 64
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_task("Mjlab-Velocity-Rough-Unitree-G1", num_envs=B, seed=3), DEV)
  cfg = env.vcfg
  assert cfg.terrain_mode == 1 and cfg.terrain_origins.shape == (10, 20, 3)
  lv0 = env._buf["terrain_level"].clone()
  assert int(lv0.max()) <= 5                       # max_init_terrain_level
  obs, _ = env.reset()
  org = torch.as_tensor(cfg.terrain_origins, dtype=torch.float32, device=DEV)
  lv, ty = env._buf["terrain_level"].long(), env._buf["terrain_type"].long()
  assert torch.equal(env._buf["env_origins"], org[lv, ty])
  # the first reset measures the distance from the construction pose (world origin) to the env origin (reference
  # behaviour): envs farther than half a tile move up one level
  far = (org[lv0.long(), ty][:, :2].norm(dim=1) > 4.0)
  assert torch.equal(lv[far], lv0[far].long() + 1)
  q = env.sim.data.qpos
  assert ((q[:, :2] - env._buf["env_origins"][:, :2]).abs() <= 0.5 + 1e-5).all()
  assert torch.allclose(q[:, 2], env._buf["env_origins"][:, 2] + float(cfg.default_root_state[2]), atol=1e-5)
  gen = torch.Generator(device=DEV).manual_seed(0)
  nres = 0
  for t in range(120):
    a = torch.rand(B, cfg.nj, device=DEV, generator=gen) * 2 - 1
    obs, rew, term, trunc, extras = env.step(a)
    assert torch.isfinite(obs["policy"]).all() and torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    nres += int((term | trunc).sum())
    lv, ty = env._buf["terrain_level"].long(), env._buf["terrain_type"].long()
    assert torch.equal(env._buf["env_origins"], org[lv, ty]) and int(lv.min()) >= 0 and int(lv.max()) < 10
  assert nres > B // 2
  assert abs(extras["log"]["Curriculum/terrain_levels"] - float(env._buf["terrain_level"].float().mean())) < 1e-4
  # random flailing does not walk anywhere: the curriculum can only have moved envs down (or kept them)
  assert int((env._buf["terrain_level"].long() <= lv0.long() + 1).all())
  assert int(env.sim.data.ncon.max()) > 0


def test_rough_env_full_size_properties():
  """8192 envs (BASELINE config 4): finite outputs, contacts against terrain boxes, bookkeeping, overflow is counted."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  B = 8192
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_task("Mjlab-Velocity-Rough-Unitree-G1", num_envs=B, seed=1), DEV)
  m = env.sim._mj_model
  cfg = env.vcfg
  env.reset()
  gen = torch.Generator(device=DEV).manual_seed(0)
  env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (B,), device=DEV, generator=gen)
  total = 0
  for t in range(40):
    a = torch.rand(B, cfg.nj, device=DEV, generator=gen) * 2 - 1
    before = env.episode_length_buf.clone()
    obs, rew, term, trunc, extras = env.step(a)
    done = term | trunc
    assert torch.isfinite(obs["policy"]).all() and torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    ep = env.episode_length_buf
    assert torch.equal(ep[~done], before[~done] + 1) and (ep[done] == 0).all()
    total += int(done.sum())
  assert total > 100
  D = env.sim.data
  ncon = D.ncon
  assert int(ncon.max()) <= 35 and float(ncon.float().mean()) > 2.0
  w = int(torch.argmax(ncon))
  g2 = D.contact_geom[w, : int(ncon[w]), 1].cpu().numpy()
  assert np.isin(g2, m.field_geom).any()                       # feet rest on terrain boxes
  g1 = D.contact_geom[w, : int(ncon[w]), 0].cpu().numpy()
  assert np.isin(g1, m.dyn_geom).all()
  assert torch.isfinite(D.qpos).all() and torch.isfinite(D.qvel).all()


def test_rough_play_mode_randomizes_terrain():
  """`play` on the rough task: 5 x 5 random grid, no pushes / noise, and the `randomize_terrain` reset event moves every
  reset env to a uniformly drawn sub-terrain (taking effect at its next reset, as in the reference's event order)."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  B = 64
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_task("Mjlab-Velocity-Rough-Unitree-G1", play=True, num_envs=B, seed=1), DEV)
  cfg = env.vcfg
  assert cfg.terrain_mode == 3 and cfg.terrain_origins.shape == (5, 5, 3) and not cfg.push_enabled and not cfg.enable_corruption
  env.reset()
  org = torch.as_tensor(cfg.terrain_origins, dtype=torch.float32, device=DEV)
  ty0 = env._buf["terrain_type"].clone()
  zero = torch.zeros(B, cfg.nj, device=DEV)
  resets = 0
  for _ in range(200):
    obs, rew, term, trunc, extras = env.step(zero)
    resets += int((term | trunc).sum())
    lv, ty = env._buf["terrain_level"].long(), env._buf["terrain_type"].long()
    assert torch.equal(env._buf["env_origins"], org[lv, ty])
    assert torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
  assert resets > 0                                   # the open-loop G1 tips over (see test_play_zero_agent)
  assert int((env._buf["terrain_type"] != ty0).sum()) > 0
  assert 0.0 <= extras["log"]["Curriculum/terrain_levels"] <= 4.0


def test_rough_env_step_closed_loop_matches_oracles():
  """env.step() on the terrain (graph replay, resets with the terrain-level curriculum, forward after resets) against
  oracle physics (numpy, float64, incl. the geom field and box contacts) + oracle env, step by step."""
  from mjlab_b200 import envabi as A
  from oracle import envstep as OE
  from tests.test_gpu_env import _batch, _scatter, make_env
  B, T = 4, 4
  env = make_env("g1", B, terrain="rough", seed=11)
  m, cfg = env.sim.mj_model, env.vcfg
  cfg.max_episode_length = 3
  A.fill_params(env._ce, cfg)
  env._capture()
  s = OE.make_state(cfg, B, env._buf["env_origins"].cpu().numpy())
  s.terrain_level[:] = env._buf["terrain_level"].cpu().numpy()
  s.terrain_type[:] = env._buf["terrain_type"].cpu().numpy()
  s.push_time_left[:] = env._buf["push_time_left"].cpu().numpy()
  ds = [O.Data(m) for _ in range(B)]
  for x in ds:
    O.forward(m, x, 35, 300)
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
  OE.observations(cfg, s, d, OE.root_kinematics(cfg, d), u)
  s.step_counter += 1       # reset() consumes one draw counter value
  torch.cuda.synchronize()
  # the first reset ran the curriculum against the construction pose (reference behaviour): levels and origins agree
  np.testing.assert_array_equal(env._buf["terrain_level"].cpu().numpy(), s.terrain_level)
  np.testing.assert_allclose(env._buf["env_origins"].cpu().numpy(), s.env_origins, rtol=0, atol=1e-6)
  np.testing.assert_allclose(env.sim.data.qpos.cpu().numpy(), d.qpos, rtol=1e-5, atol=2e-5)
  np.testing.assert_allclose(obs["critic"].cpu().numpy(), s.obs_critic, rtol=2e-3, atol=2e-3)
  rng = np.random.default_rng(5)
  resets = 0
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
      for x in ds:
        O.forward(m, x, 35, 300)
      d = _batch(ds, m)
    OE.obs_stage(cfg, m, s, d, u=u)
    _scatter(ds, d)
    torch.cuda.synchronize()
    resets += int(s.reset_buf.sum())
    np.testing.assert_array_equal(term.cpu().numpy(), s.terminated)
    np.testing.assert_array_equal(trunc.cpu().numpy(), s.time_outs)
    np.testing.assert_array_equal(env._buf["terrain_level"].cpu().numpy(), s.terrain_level)
    np.testing.assert_allclose(env._buf["env_origins"].cpu().numpy(), s.env_origins, rtol=0, atol=1e-6)
    np.testing.assert_allclose(rew.cpu().numpy(), s.reward, rtol=5e-3, atol=3e-4)
    np.testing.assert_allclose(env.sim.data.qpos.cpu().numpy(), d.qpos, rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(obs["critic"].cpu().numpy(), s.obs_critic, rtol=2e-2, atol=1e-2)
    assert extras["log"]["Episode/num_resets"] == float(s.reset_buf.sum())
    assert abs(extras["log"]["Curriculum/terrain_levels"] - float(s.terrain_level.mean())) < 1e-5
  assert resets >= B           # max_episode_length 3: every env timed out once and went through the curriculum
