"""Host-side logic of the PPO learner (mjlab_b200/ppo.py) on CPU: running normalisation, GAE + time-out bootstrapping
against a scalar restatement, clipped losses decrease on a toy env, gradient averaging over 2 gloo ranks."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mjlab_b200 import ppo


def test_empirical_normalization_matches_numpy():
  torch.manual_seed(0)
  norm = ppo.EmpiricalNormalization(5)
  xs = [torch.randn(64, 5) * 3 + 2, torch.randn(17, 5) - 4, torch.randn(128, 5) * 0.1]
  for x in xs:
    norm.update(x)
  allx = torch.cat(xs).numpy()
  np.testing.assert_allclose(norm.mean.numpy()[0], allx.mean(0), rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(norm.var.numpy()[0], allx.var(0), rtol=1e-4, atol=1e-6)
  y = norm(torch.as_tensor(allx))
  assert abs(float(y.mean())) < 0.05 and abs(float(y.std()) - 1) < 0.05


class ToyEnv:
  """Reward = -|action - obs[:, :2]|^2, fixed horizon with time-outs; the vec-env surface OnPolicyRunner needs."""

  def __init__(self, n=64, horizon=10, device="cpu"):
    self.num_envs, self.num_actions, self.device, self.max_episode_length = n, 2, device, horizon
    self.episode_length_buf = torch.zeros(n, dtype=torch.long)
    self.g = torch.Generator().manual_seed(1)
    self._obs = torch.randn(n, 3, generator=self.g)

  def get_observations(self):
    return {"policy": self._obs, "critic": self._obs}

  def step(self, a):
    rew = -((a - self._obs[:, :2]) ** 2).sum(-1)
    self.episode_length_buf += 1
    to = self.episode_length_buf >= self.max_episode_length
    self.episode_length_buf[to] = 0
    self._obs = torch.randn(self.num_envs, 3, generator=self.g)
    return self.get_observations(), rew, to.long(), {"time_outs": to}


def test_gae_matches_scalar_restatement_and_training_improves():
  env = ToyEnv()
  cfg = ppo.PpoCfg(num_steps_per_env=8, actor_hidden_dims=(32,), critic_hidden_dims=(32,), num_learning_epochs=2, num_mini_batches=2)
  r = ppo.OnPolicyRunner(env, cfg, device="cpu")
  ro = r.collect()
  st = r.st
  T, B = st["rewards"].shape
  with torch.no_grad():
    last = r.policy.evaluate(env.get_observations()["critic"]).squeeze(-1)
  # note: `last` uses normaliser statistics updated during the roll-out, like the runner did at its end
  ret = torch.zeros(T, B)
  adv = torch.zeros(B)
  for t in reversed(range(T)):
    nv = last if t == T - 1 else st["values"][t + 1]
    nt = 1 - st["dones"][t]
    delta = st["rewards"][t] + nt * cfg.gamma * nv - st["values"][t]
    adv = delta + nt * cfg.gamma * cfg.lam * adv
    ret[t] = adv + st["values"][t]
  torch.testing.assert_close(ro["returns"], ret, rtol=1e-5, atol=1e-5)
  a = ro["advantages"]
  assert abs(float(a.mean())) < 1e-5 and abs(float(a.std()) - 1) < 1e-3
  first = float(ro["mean_step_reward"])
  hist = r.learn(60, init_at_random_ep_len=False)
  assert hist[-1]["mean_step_reward"] > first + 0.5, (first, hist[-1])
  assert all(np.isfinite(h["kl"]) for h in hist)


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(0)
  net = torch.nn.Linear(3, 2)
  x = torch.full((4, 3), float(rank + 1))
  net(x).sum().backward()
  local = [p.grad.clone() for p in net.parameters()]
  ppo.reduce_gradients(list(net.parameters()), world)
  q.put((rank, [g.numpy() for g in local], [p.grad.numpy().copy() for p in net.parameters()]))
  dist.barrier()
  dist.destroy_process_group()


def test_gradient_all_reduce_averages_over_ranks():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29650 + os.getpid() % 200
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  out = sorted(q.get(timeout=120) for _ in range(2))
  for p in procs:
    p.join(timeout=60)
  (_, l0, a0), (_, l1, a1) = out
  for g0, g1, r0, r1 in zip(l0, l1, a0, a1):
    np.testing.assert_allclose(r0, (g0 + g1) / 2, rtol=1e-6)
    np.testing.assert_allclose(r1, r0, rtol=0, atol=0)


def _seed_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  env = ToyEnv(n=16, horizon=10)
  cfg = ppo.PpoCfg(num_steps_per_env=4, actor_hidden_dims=(16,), critic_hidden_dims=(16,), num_learning_epochs=1, num_mini_batches=2, seed=7)
  r = ppo.OnPolicyRunner(env, cfg, device="cpu")
  params = torch.cat([p.detach().view(-1) for p in r.policy.parameters()]).numpy().copy()
  ro = r.collect()
  actions = r.st["actions"].numpy().copy()
  # the distributed update: gradients are views of one flat buffer, [gradients | KL] all-reduced once per mini-batch
  stats = r.update(ro)
  after = torch.cat([p.detach().view(-1) for p in r.policy.parameters()]).numpy().copy()
  assert r._flat is not None and all(p.grad.data_ptr() >= r._flat.data_ptr() for p in r._params)
  # the toy env's observations are identical on both ranks (same generator seed), so any difference in the sampled
  # actions comes from the exploration noise alone
  q.put((rank, params, actions, after, float(stats["kl"]), float(r._lr)))
  dist.barrier()
  dist.destroy_process_group()


def test_ranks_share_parameters_but_not_exploration_noise():
  """every rank seeded torch with cfg.seed, so all replicas drew the same action noise.  The reference seeds the
  agent with seed + local_rank (scripts/train.py:58-63): parameters equal (rank-0 broadcast), noise different."""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29850 + os.getpid() % 100
  procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  out = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
  for p in procs:
    p.join(timeout=60)
  (_, p0, a0, f0, kl0, lr0), (_, p1, a1, f1, kl1, lr1) = out
  np.testing.assert_array_equal(p0, p1)
  assert np.abs(a0 - a1).max() > 1e-3, "both ranks sampled identical actions: exploration noise is not rank-specific"
  # after the update the replicas still hold identical parameters (averaged gradients, the same KL -> the same learning rate)
  np.testing.assert_array_equal(f0, f1)
  assert np.abs(f0 - p0).max() > 1e-5 and kl0 == kl1 and lr0 == lr1


def test_policy_export_roundtrip(tmp_path):
  """TorchScript export of actor(normalizer(obs)) reproduces the live policy; the metadata side-car carries the reference's
  keys (rl/exporter_utils.py:23-62).  The env is a stand-in with the attributes the exporter reads."""
  import json
  from types import SimpleNamespace as NS
  import mjlab_b200
  from mjlab_b200 import envcfg, exporter
  m = mjlab_b200.load_model("go1_flat")
  cfg = envcfg.velocity_cfg(m, "go1")
  env = NS(scene={"robot": NS(joint_names=[n.removeprefix("robot/") for n in cfg.joint_names])}, sim=NS(mj_model=m), vcfg=cfg,
           command_manager=NS(active_terms=["twist"]), observation_manager=NS(active_terms={"policy": ["base_lin_vel", "actions"]}))
  torch.manual_seed(0)
  ac = ppo.ActorCritic(cfg.n_obs_policy, cfg.n_obs_critic, cfg.nj, ppo.PpoCfg())
  ac.update_normalization(torch.randn(256, cfg.n_obs_policy) * 2 + 1, torch.randn(256, cfg.n_obs_critic))
  out = exporter.export_policy_as_jit(ac, env, str(tmp_path), run_path="local/test")
  loaded = torch.jit.load(out)
  x = torch.randn(5, cfg.n_obs_policy)
  ac.eval()
  with torch.no_grad():
    np.testing.assert_allclose(loaded(x).numpy(), ac.act_inference(x).numpy(), rtol=1e-5, atol=1e-6)
  meta = json.load(open(out.replace(".pt", ".json")))
  assert set(meta) == {"run_path", "joint_names", "joint_stiffness", "joint_damping", "default_joint_pos", "command_names",
                       "observation_names", "action_scale"}
  assert len(meta["joint_names"].split(",")) == cfg.nj == len(meta["joint_stiffness"].split(","))
  kp = np.asarray(m.actuator_gainprm)[cfg.ctrl_ids, 0]
  np.testing.assert_allclose([float(v) for v in meta["joint_stiffness"].split(",")], kp, atol=1e-3)
  with pytest.raises(RuntimeError):
    exporter.export_policy_as_onnx(ac, env, str(tmp_path))
