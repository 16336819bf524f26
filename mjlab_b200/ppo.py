"""PPO learner for the B200 env (the step after the hot path in `train`).

The reference trains with rsl-rl-lib 3.1.0's `OnPolicyRunner` (src/mjlab/scripts/train.py:141-169, configured by
src/mjlab/rl/config.py:7-112 and src/mjlab/tasks/velocity/config/g1/rl_cfg.py:10-39); that package is not in the
reference tree and not installable here, so this module restates the algorithm it runs for the velocity tasks:
actor / critic MLPs (ELU, 512-256-128) with empirical observation normalisation and a state-independent scalar std,
24 steps per env per iteration, GAE(0.99, 0.95) with time-out bootstrapping, 5 epochs x 4 mini-batches of the clipped
surrogate + clipped value loss + entropy bonus, KL-adaptive learning rate, gradient-norm clipping, and for multi-GPU runs
(`--gpu-ids`, docs/api/distributed_training.md) one process per GPU with independent env batches whose gradients are
averaged by an NCCL all-reduce before every optimiser step (the only collective of the training path).

This is host-side plumbing around the env hot path: plain PyTorch modules (their GEMMs are cuBLAS library calls), no
custom kernels.  It exists so that BASELINE config 3 ("full PPO train step") can be measured end to end on this engine
and so that the env can be validated by actually learning on it.
"""

from __future__ import annotations

import math

from dataclasses import dataclass, field

import torch
import torch.distributed as dist
from torch import nn


@dataclass
class PpoCfg:
  """rl/config.py:7-112 + config/g1/rl_cfg.py:10-39."""

  num_steps_per_env: int = 24
  actor_hidden_dims: tuple = (512, 256, 128)
  critic_hidden_dims: tuple = (512, 256, 128)
  init_noise_std: float = 1.0
  actor_obs_normalization: bool = True
  critic_obs_normalization: bool = True
  num_learning_epochs: int = 5
  num_mini_batches: int = 4
  learning_rate: float = 1.0e-3
  schedule: str = "adaptive"
  gamma: float = 0.99
  lam: float = 0.95
  entropy_coef: float = 0.01
  desired_kl: float = 0.01
  max_grad_norm: float = 1.0
  value_loss_coef: float = 1.0
  use_clipped_value_loss: bool = True
  clip_param: float = 0.2
  clip_actions: float | None = None
  # "highest" = rsl_rl's plain fp32 (the default: numbers quoted against the reference use it); "high": TF32 tensor-core GEMMs
  # for the actor / critic (mini-batches of 24 576 x 512), reported separately
  matmul_precision: str = "highest"
  # capture the whole 24-step roll-out (policy inference + the env's kernels + GAE) in ONE CUDA graph: removes ~2000 eager
  # launches per iteration; needs the env's raw enqueue API, falls back to eager otherwise
  rollout_cuda_graph: bool = True
  # capture the 20 mini-batch updates of an iteration (forward, backward, clip, Adam, KL-adaptive learning rate on the
  # device) in ONE CUDA graph; single-process only (the distributed path all-reduces gradients eagerly)
  update_cuda_graph: bool = True
  seed: int = 42
  obs_groups: dict = field(default_factory=lambda: {"policy": ("policy",), "critic": ("critic",)})


class EmpiricalNormalization(nn.Module):
  """Running mean / variance normalisation, updated while the module is in training mode."""

  def __init__(self, dim: int, eps: float = 1e-2):
    super().__init__()
    self.eps = eps
    self.register_buffer("mean", torch.zeros(1, dim))
    self.register_buffer("var", torch.ones(1, dim))
    self.register_buffer("std", torch.ones(1, dim))
    self.register_buffer("count", torch.zeros((), dtype=torch.long))

  @torch.no_grad()
  def update(self, x: torch.Tensor) -> None:
    n = x.shape[0]
    self.count += n
    rate = n / self.count.to(x.dtype)
    mean_x = x.mean(0, keepdim=True)
    var_x = x.var(0, unbiased=False, keepdim=True)
    delta = mean_x - self.mean
    self.mean += rate * delta
    self.var += rate * (var_x - self.var + delta * (mean_x - self.mean))
    self.std.copy_(self.var.sqrt())

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    return (x - self.mean) / (self.std + self.eps)


def _mlp(n_in: int, hidden: tuple, n_out: int) -> nn.Sequential:
  layers: list[nn.Module] = []
  d = n_in
  for h in hidden:
    layers += [nn.Linear(d, h), nn.ELU()]
    d = h
  layers.append(nn.Linear(d, n_out))
  return nn.Sequential(*layers)


def _pad4(n: int) -> int:
  return (n + 3) & ~3


class ActorCritic(nn.Module):
  """rsl_rl's ActorCritic (MLP actor + MLP critic, state-independent action std, empirical observation normalisation).

  The first and last layers are built on feature counts rounded up to a multiple of 4 (99 -> 100 observations, 29 -> 32
  actions): cuBLAS only takes its TMA / vectorised kernels when the leading dimensions are 16-byte aligned, and the
  `align1` fallbacks were 8 of the 31 ms of a PPO update.  The extra inputs are always zero and the extra outputs are never
  read, so the function computed (and every gradient that matters) is that of the unpadded network."""

  def __init__(self, n_actor_obs: int, n_critic_obs: int, n_actions: int, cfg: PpoCfg):
    super().__init__()
    self.n_actor_obs, self.n_critic_obs, self.n_actions = n_actor_obs, n_critic_obs, n_actions
    self.w_actor_obs, self.w_critic_obs, self.w_actions = _pad4(n_actor_obs), _pad4(n_critic_obs), _pad4(n_actions)
    self.actor = _mlp(self.w_actor_obs, cfg.actor_hidden_dims, self.w_actions)
    self.critic = _mlp(self.w_critic_obs, cfg.critic_hidden_dims, 4)
    self.actor_norm = EmpiricalNormalization(self.w_actor_obs) if cfg.actor_obs_normalization else nn.Identity()
    self.critic_norm = EmpiricalNormalization(self.w_critic_obs) if cfg.critic_obs_normalization else nn.Identity()
    self.std = nn.Parameter(cfg.init_noise_std * torch.ones(n_actions))

  @staticmethod
  def _fit(x: torch.Tensor, width: int) -> torch.Tensor:
    return x if x.shape[-1] == width else nn.functional.pad(x, (0, width - x.shape[-1]))

  def actor_mean(self, obs: torch.Tensor) -> torch.Tensor:
    """obs: [.., n_actor_obs] or already padded to w_actor_obs."""
    return self.actor(self.actor_norm(self._fit(obs, self.w_actor_obs)))[..., : self.n_actions]

  def distribution(self, obs: torch.Tensor) -> torch.distributions.Normal:
    mean = self.actor_mean(obs)
    return torch.distributions.Normal(mean, self.std.expand_as(mean))

  def act_inference(self, obs: torch.Tensor) -> torch.Tensor:
    return self.actor_mean(obs)

  def evaluate(self, critic_obs: torch.Tensor) -> torch.Tensor:
    return self.critic(self.critic_norm(self._fit(critic_obs, self.w_critic_obs)))[..., :1]

  @torch.no_grad()
  def update_normalization(self, obs: torch.Tensor, critic_obs: torch.Tensor) -> None:
    if isinstance(self.actor_norm, EmpiricalNormalization):
      self.actor_norm.update(self._fit(obs, self.w_actor_obs))
    if isinstance(self.critic_norm, EmpiricalNormalization):
      self.critic_norm.update(self._fit(critic_obs, self.w_critic_obs))


def reduce_gradients(params, world_size: int) -> None:
  """Average the gradients over ranks with ONE all-reduce of the flattened vector (1.76 MB for the G1 nets)."""
  grads = [p.grad.view(-1) for p in params if p.grad is not None]
  flat = torch.cat(grads)
  dist.all_reduce(flat, op=dist.ReduceOp.SUM)
  flat /= world_size
  o = 0
  for p in params:
    if p.grad is not None:
      n = p.numel()
      p.grad.copy_(flat[o:o + n].view_as(p.grad))
      o += n


class OnPolicyRunner:
  """Roll-out + PPO update loop over a `ManagerBasedRlEnv` (through `RslRlVecEnvWrapper`)."""

  def __init__(self, venv, cfg: PpoCfg | None = None, device: str | None = None):
    self.env, self.cfg = venv, cfg or PpoCfg()
    self.device = torch.device(device or venv.device)
    torch.set_float32_matmul_precision(self.cfg.matmul_precision)
    self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    self.world_size = dist.get_world_size() if self.distributed else 1
    self.rank = dist.get_rank() if self.distributed else 0
    obs = venv.get_observations()
    self.n_obs, self.n_critic = obs["policy"].shape[1], obs["critic"].shape[1]
    self.n_act = venv.num_actions
    torch.manual_seed(self.cfg.seed)  # identical initial parameters on every rank; env batches differ by seed + rank
    self.policy = ActorCritic(self.n_obs, self.n_critic, self.n_act, self.cfg).to(self.device)
    if self.distributed:
      for t in list(self.policy.parameters()) + list(self.policy.buffers()):
        dist.broadcast(t.data, src=0)
    # Exploration noise and mini-batch permutations must differ per process (the reference seeds the whole agent with
    # seed + local_rank, scripts/train.py:58-63); the parameters above stay identical on every rank.
    torch.manual_seed(self.cfg.seed + self.rank)
    if torch.cuda.is_available():
      torch.cuda.manual_seed(self.cfg.seed + self.rank)
    # the learning rate lives on the device (KL-adaptive rule without a host round trip per mini-batch)
    cuda = self.device.type == "cuda"
    self._lr = torch.tensor(float(self.cfg.learning_rate), device=self.device)
    # fused: one multi-tensor kernel per step (capturable in a CUDA graph, learning rate read from the device tensor)
    self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=self._lr, **(dict(capturable=True, fused=True) if cuda else dict(foreach=False)))
    self._update_graph, self._update_warm = None, 0
    T, B = self.cfg.num_steps_per_env, venv.num_envs
    z = lambda *s: torch.zeros(T, B, *s, device=self.device)  # noqa: E731
    self.st = dict(obs=z(self.n_obs), critic=z(self.n_critic), actions=z(self.n_act), rewards=z(), dones=z(), values=z(),
                   logp=z(), mu=z(self.n_act), sigma=z(self.n_act))
    self.ro = dict(returns=z(), advantages=z(), mean_step_reward=torch.zeros((), device=self.device),
                   done_ret=torch.zeros((), device=self.device), done_len=torch.zeros((), device=self.device), done_n=torch.zeros((), device=self.device))
    self._rollout_graph, self._graph_env_epoch = None, -1
    torch.distributions.Distribution.set_default_validate_args(False)   # its checks synchronise with the host
    self.iteration = 0
    self.tot_steps = 0
    self.history: list[dict] = []
    self._ep_ret = torch.zeros(B, device=self.device)
    self._ep_len = torch.zeros(B, device=self.device)

  # ---- roll-out -------------------------------------------------------------------------------------
  def _raw_env(self):
    e = getattr(self.env, "env", None)
    if e is not None and getattr(getattr(getattr(e, "sim", None), "nan_guard", None), "enabled", False):
      return None   # NanGuard watches env.step(); the captured roll-out graph bypasses it, so fall back to stepping the env
    return e if (e is not None and hasattr(e, "_enqueue_step") and str(self.device).startswith("cuda")) else None

  def _step_env(self, a: torch.Tensor, raw):
    """(obs, reward, dones, time_outs) of one env step; `raw`: enqueue the env's kernels directly (graph capture)."""
    if raw is None:
      obs, rew, dones, extras = self.env.step(a)
      return obs, rew, dones, extras.get("time_outs")
    if self.env.clip_actions is not None:
      a = torch.clamp(a, -self.env.clip_actions, self.env.clip_actions)
    raw._buf["action_in"].copy_(a)
    raw._enqueue_step()
    return raw.obs_buf, raw.reward_buf, (raw.reset_terminated | raw.reset_time_outs).to(torch.long), raw.reset_time_outs

  @torch.no_grad()
  def _rollout_body(self, raw) -> None:
    cfg, st, ro = self.cfg, self.st, self.ro
    obs = self.env.get_observations() if raw is None else raw.obs_buf
    for k in ("done_ret", "done_len", "done_n"):
      ro[k].zero_()
    self.policy.train()
    for t in range(cfg.num_steps_per_env):
      st["obs"][t].copy_(obs["policy"])          # the env's observation buffers are overwritten by the next step
      st["critic"][t].copy_(obs["critic"])
      dist_ = self.policy.distribution(st["obs"][t])
      a = dist_.mean + dist_.stddev * torch.randn_like(dist_.mean)   # == dist_.sample() without its host-side std check
      st["actions"][t] = a
      st["values"][t] = self.policy.evaluate(st["critic"][t]).squeeze(-1)
      st["logp"][t] = dist_.log_prob(a).sum(-1)
      st["mu"][t], st["sigma"][t] = dist_.mean, dist_.stddev
      obs, rew, dones, time_outs = self._step_env(a, raw)
      self.policy.update_normalization(obs["policy"], obs["critic"])
      r = rew.clone()
      if time_outs is not None:  # bootstrap the value of truncated episodes
        r += cfg.gamma * st["values"][t] * time_outs.to(r.dtype)
      st["rewards"][t] = r
      st["dones"][t] = dones.to(r.dtype)
      self._ep_ret += rew
      self._ep_len += 1
      d = dones.bool()
      ro["done_n"] += d.sum()
      ro["done_ret"] += (self._ep_ret * d).sum()
      ro["done_len"] += (self._ep_len * d).sum()
      self._ep_ret.masked_fill_(d, 0.0)
      self._ep_len.masked_fill_(d, 0.0)
    last_values = self.policy.evaluate(obs["critic"]).squeeze(-1)
    adv = torch.zeros_like(last_values)   # GAE
    for t in reversed(range(cfg.num_steps_per_env)):
      nv = last_values if t == cfg.num_steps_per_env - 1 else st["values"][t + 1]
      nonterminal = 1.0 - st["dones"][t]
      delta = st["rewards"][t] + nonterminal * cfg.gamma * nv - st["values"][t]
      adv = delta + nonterminal * cfg.gamma * cfg.lam * adv
      ro["returns"][t] = adv + st["values"][t]
    advantages = ro["returns"] - st["values"]
    ro["advantages"].copy_((advantages - advantages.mean()) / (advantages.std() + 1e-8))
    ro["mean_step_reward"].copy_(st["rewards"].mean())

  def _capture_rollout(self, raw) -> None:
    torch.cuda.synchronize(self.device)
    side = torch.cuda.Stream(self.device)
    with torch.cuda.stream(side):
      self._rollout_body(raw)            # warm-up outside capture (library workspaces, lazy inits); a real roll-out
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g, stream=side):
        self._rollout_body(raw)
    torch.cuda.current_stream(self.device).wait_stream(side)
    self._rollout_graph, self._graph_env_epoch = g, getattr(raw, "_capture_epoch", 0)

  def collect(self) -> dict:
    cfg = self.cfg
    raw = self._raw_env() if cfg.rollout_cuda_graph else None
    T = cfg.num_steps_per_env
    if raw is not None:
      if self._rollout_graph is None or self._graph_env_epoch != getattr(raw, "_capture_epoch", 0):
        self._capture_rollout(raw)       # its warm-up pass WAS a real roll-out (capture itself executes nothing)
        raw.note_steps(T)
      else:
        self._rollout_graph.replay()
        raw.note_steps(T)
    else:
      self._rollout_body(None)
    self.tot_steps += T * self.env.num_envs * self.world_size
    return self.ro

  # ---- update ------------------------------------------------------------------------------------------
  @property
  def learning_rate(self) -> float:
    return float(self._lr)

  def _pack_rollout(self) -> None:
    """One packed copy of the roll-out, [N, obs | critic | actions | mu | sigma | logp | value | return | advantage]: a mini-batch
    is then ONE row gather.  Column blocks start on multiples of 4 floats and the observation blocks have the networks' padded
    widths (their padding columns stay zero): every view handed to a GEMM is 16-byte aligned."""
    cfg, st, ro = self.cfg, self.st, self.ro
    N = cfg.num_steps_per_env * self.env.num_envs
    no, nc, na = self.n_obs, self.n_critic, self.n_act
    pol = self.policy
    cols = [("obs", no, pol.w_actor_obs), ("critic", nc, pol.w_critic_obs), ("actions", na, pol.w_actions), ("mu", na, pol.w_actions),
            ("sigma", na, pol.w_actions), ("logp", 1, 1), ("values", 1, 1)]
    width = sum(w for _, _, w in cols) + 2
    if getattr(self, "_pack", None) is None or self._pack.shape != (N, width):
      self._pack = torch.zeros(N, width, device=self.device)
    off, o = {}, 0
    for k, n_real, w in cols:
      self._pack[:, o:o + n_real] = st[k].reshape(N, n_real)
      off[k] = (o, o + n_real)
      off[k + "_padded"] = (o, o + w)
      o += w
    self._pack[:, o] = ro["returns"].reshape(N)
    self._pack[:, o + 1] = ro["advantages"].reshape(N)
    off["ret"], off["adv"] = o, o + 1
    self._off = off

  def _mb_backward(self, idx: torch.Tensor) -> None:
    """Forward, losses and backward of one mini-batch (rows `idx` of the packed roll-out).  Leaves the gradients in the
    parameters' .grad (views of `self._flat` in distributed runs), this rank's KL in `self._kl` and the loss terms in
    `self._mb_stats`.  The Gaussian terms are written out (log-prob, entropy and KL of a diagonal Normal with a state-independent
    std) instead of going through torch.distributions - same formulas as rsl_rl's PPO.update."""
    cfg, off, pol, na = self.cfg, self._off, self.policy, self.n_act
    half_log_2pi = 0.5 * math.log(2.0 * math.pi)
    b = self._pack[idx]
    col = lambda k: b[:, off[k][0]:off[k][1]]  # noqa: E731
    mu = pol.actor_mean(col("obs_padded"))
    value = pol.evaluate(col("critic_padded")).squeeze(-1)
    std = pol.std
    log_std = std.log()
    logp = (-0.5 * ((col("actions") - mu) / std).square() - log_std).sum(-1) - na * half_log_2pi
    entropy = (0.5 + half_log_2pi + log_std).sum()            # per-row entropy of the diagonal Normal; identical rows
    with torch.no_grad():
      old_mu, old_sigma = col("mu"), col("sigma")
      kl = torch.sum(torch.log(std / old_sigma + 1e-5) + (old_sigma.square() + (old_mu - mu).square()) / (2.0 * std.square()) - 0.5, -1).mean()
    ratio = torch.exp(logp - b[:, off["logp"][0]])
    a, ret = b[:, off["adv"]], b[:, off["ret"]]
    surrogate = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - cfg.clip_param, 1.0 + cfg.clip_param)).mean()
    if cfg.use_clipped_value_loss:
      old_v = b[:, off["values"][0]]
      v_clip = old_v + (value - old_v).clamp(-cfg.clip_param, cfg.clip_param)
      v_loss = torch.max((value - ret).square(), (v_clip - ret).square()).mean()
    else:
      v_loss = (ret - value).square().mean()
    loss = surrogate + cfg.value_loss_coef * v_loss - cfg.entropy_coef * entropy
    if self._flat is not None:
      self._flat.zero_()
    else:
      self.optimizer.zero_grad(set_to_none=False)
    loss.backward()
    with torch.no_grad():
      self._kl.copy_(kl)
      self._mb_stats.copy_(torch.stack((v_loss.detach(), surrogate.detach(), entropy.detach())))

  def _mb_step(self) -> None:
    """After the gradients (and the KL) have been averaged over the ranks: KL-adaptive learning rate (on the device), gradient
    clipping, Adam.  The learning rate only enters the optimiser step, so adapting it here equals rsl_rl's order."""
    cfg = self.cfg
    with torch.no_grad():
      if self._flat is not None:
        self._flat /= self.world_size          # [gradients | KL], summed over the ranks by ONE all-reduce
      kl = self._kl
      if cfg.schedule == "adaptive" and cfg.desired_kl is not None:
        lr = self._lr
        down = torch.clamp(lr / 1.5, min=1e-5)
        up = torch.clamp(lr * 1.5, max=1e-2)
        lr.copy_(torch.where(kl > cfg.desired_kl * 2.0, down, torch.where((kl > 0.0) & (kl < cfg.desired_kl / 2.0), up, lr)))
    nn.utils.clip_grad_norm_(self._params, cfg.max_grad_norm, foreach=self.device.type == "cuda")
    self.optimizer.step()
    with torch.no_grad():
      self._stats[:3] += self._mb_stats
      self._stats[3] += kl

  def _setup_update(self) -> None:
    cfg = self.cfg
    N = cfg.num_steps_per_env * self.env.num_envs
    self._perm = torch.zeros(cfg.num_learning_epochs, N, dtype=torch.long, device=self.device)
    self._stats = torch.zeros(4, device=self.device)
    self._mb_stats = torch.zeros(3, device=self.device)
    self._params = [p for p in self.policy.parameters() if p.requires_grad]
    self._cur_idx = torch.zeros(N // cfg.num_mini_batches, dtype=torch.long, device=self.device)
    self._flat = None
    if self.distributed:
      # every gradient is a view of ONE flat buffer whose last element carries the mini-batch KL: the ranks exchange both
      # with a single all-reduce per mini-batch (1.76 MB for the G1 networks), no flatten / unflatten copies
      n = sum(p.numel() for p in self._params)
      self._flat = torch.zeros(n + 1, device=self.device)
      o = 0
      for p in self._params:
        p.grad = self._flat[o:o + p.numel()].view_as(p)
        o += p.numel()
      self._kl = self._flat[n:n + 1].view(())
    else:
      self._kl = torch.zeros((), device=self.device)
    self._graph_a = self._graph_b = None

  def _update_body(self) -> None:
    """All epochs x mini-batches of one PPO update; reads `self._perm`, writes `self._stats` (device tensors)."""
    cfg = self.cfg
    self._pack_rollout()
    N = cfg.num_steps_per_env * self.env.num_envs
    mb = N // cfg.num_mini_batches
    self._stats.zero_()
    self.policy.train()
    for e in range(cfg.num_learning_epochs):
      for i in range(cfg.num_mini_batches):
        self._mb_backward(self._perm[e, i * mb:(i + 1) * mb])
        if self.distributed:
          dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        self._mb_step()
    self._stats /= cfg.num_learning_epochs * cfg.num_mini_batches

  def _update_distributed_graphs(self) -> None:
    """Distributed update without eager kernel launches: per mini-batch ONE graph replay for forward + backward, ONE NCCL
    all-reduce of the flat [gradients | KL] buffer, ONE graph replay for clip + Adam.  (Capturing the NCCL call itself inside a
    CUDA graph hung with this torch / NCCL pair in round 1; keeping the collective outside the two graphs needs no such support
    and costs 3 launches per mini-batch instead of ~75.)"""
    cfg = self.cfg
    N = cfg.num_steps_per_env * self.env.num_envs
    mb = N // cfg.num_mini_batches
    self._pack_rollout()
    self._stats.zero_()
    self.policy.train()
    if self._graph_a is None:
      torch.cuda.synchronize(self.device)
      side = torch.cuda.Stream(self.device)
      ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
      with torch.cuda.stream(side):
        with torch.cuda.graph(ga, stream=side):
          self._mb_backward(self._cur_idx)
        with torch.cuda.graph(gb, stream=side):
          self._mb_step()
      torch.cuda.current_stream(self.device).wait_stream(side)
      self._graph_a, self._graph_b = ga, gb
    for e in range(cfg.num_learning_epochs):
      for i in range(cfg.num_mini_batches):
        self._cur_idx.copy_(self._perm[e, i * mb:(i + 1) * mb])
        self._graph_a.replay()
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        self._graph_b.replay()
    self._stats /= cfg.num_learning_epochs * cfg.num_mini_batches

  def update(self, ro: dict) -> dict:
    cfg = self.cfg
    N = cfg.num_steps_per_env * self.env.num_envs
    if not hasattr(self, "_perm"):
      self._setup_update()
    for e in range(cfg.num_learning_epochs):
      self._perm[e].copy_(torch.randperm(N, device=self.device))
    use_graph = cfg.update_cuda_graph and self.device.type == "cuda"
    if not use_graph or self._update_warm < 2:
      self._update_body()               # eager: also the warm-up that creates Adam's state and the .grad tensors
      self._update_warm += 1
    elif self.distributed:
      self._update_distributed_graphs()
    else:
      if self._update_graph is None:
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
          self._update_body()           # capture executes nothing; the replay below is this iteration's update
        self._update_graph = g
      self._update_graph.replay()
    return dict(zip(("value", "surrogate", "entropy", "kl"), self._stats.clone().unbind()))

  def learn(self, num_iterations: int, init_at_random_ep_len: bool = True, log_every: int = 0) -> list[dict]:
    if init_at_random_ep_len and self.iteration == 0:  # scripts/train.py:167-169
      self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
    for _ in range(num_iterations):
      ro = self.collect()
      up = self.update(ro)
      self.iteration += 1
      n = max(float(ro["done_n"]), 1.0)
      rec = dict(iteration=self.iteration, mean_step_reward=float(ro["mean_step_reward"]), episodes=float(ro["done_n"]),
                 mean_episode_return=float(ro["done_ret"]) / n, mean_episode_length=float(ro["done_len"]) / n,
                 lr=self.learning_rate, noise_std=float(self.policy.std.detach().mean()), **{k: float(v) for k, v in up.items()})
      self.history.append(rec)
      if log_every and self.rank == 0 and self.iteration % log_every == 0:
        print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in rec.items()}, flush=True)
    return self.history
