"""Per-term post-processing of the observation groups: clip -> scale -> delay -> history, then concatenation
(src/mjlab/managers/observation_manager.py:154-208; docs/api/observation_history_delay.md).

The fused `k_env_obs` kernel does the first two stages of the reference's pipeline for every term (compute, noise) and writes
the groups contiguously.  The remaining stages only exist when a term asks for them (none of the shipped tasks does), so they
run here, outside the step graph, on slices of the kernel's output: `DelayBuffer` (CUDA ring, `b2_delay_step`) quantised to
CONTROL steps, `CircularBuffer` history flattened term-major `[A_t0 .. A_tH-1, B_t0 ..]`, oldest first; `compute()` without
`update_history` returns the cached result (no double push, observation_manager.py:157-161); `reset(env_ids)` clears the
buffers of the reset envs so that their next push back-fills.
"""

from __future__ import annotations

import torch

from .buffers import CircularBuffer


class ObservationPipeline:
  def __init__(self, layout: dict[str, list[tuple[str, int]]], post: dict[str, list[dict]], num_envs: int, device: str):
    """layout: group -> [(term name, width)] in the kernel's order, or [(term name, width, source group, offset)]; post: group -> per-term settings (`envcfg.from_manager_cfg`
    `extra["obs_post"]`: clip, scale, delay_*, history_length, flatten_history_dim)."""
    from .actuators import DelayBuffer
    self.num_envs, self.device = num_envs, device
    self.groups: dict[str, list[dict]] = {}
    for g, terms in layout.items():
      o, out = 0, []
      assert len(terms) == len(post[g]), (g, len(terms), len(post[g]))
      for entry, p in zip(terms, post[g]):
        # (name, width): the next `width` columns of the kernel's group g; (name, width, source group, offset): a gather from the
        # kernel's term bank (envcfg.from_manager_cfg `extra["obs_layout"]`: any subset / order / group name)
        name, dim = entry[0], entry[1]
        src, off = (entry[2], entry[3]) if len(entry) == 4 else (g, o)
        t = dict(name=name, src=src, sl=slice(off, off + dim), dim=dim, clip=tuple(p["clip"]) if p.get("clip") else None, scale=None, delay=None,
                 history=None, flatten=bool(p.get("flatten_history_dim", True)))
        if p.get("scale") is not None:
          t["scale"] = torch.as_tensor(p["scale"], dtype=torch.float32, device=device)
        if p.get("delay_max_lag", 0) > 0:
          t["delay"] = DelayBuffer(min_lag=p.get("delay_min_lag", 0), max_lag=p["delay_max_lag"], batch_size=num_envs, device=device,
                                   per_env=p.get("delay_per_env", True), hold_prob=p.get("delay_hold_prob", 0.0),
                                   update_period=p.get("delay_update_period", 0), per_env_phase=p.get("delay_per_env_phase", True))
        if p.get("history_length", 0) > 0:
          t["history"] = CircularBuffer(max_len=p["history_length"], batch_size=num_envs, device=device)
        out.append(t)
        o += dim
      self.groups[g] = out
    self._cache: dict[str, torch.Tensor] | None = None

  def term_dims(self, group: str) -> list[tuple[int, ...]]:
    """Shape of every term after history (observation_manager.py `group_obs_term_dim`)."""
    out = []
    for t in self.groups[group]:
      if t["history"] is None:
        out.append((t["dim"],))
      else:
        H = t["history"].max_length
        out.append((H * t["dim"],) if t["flatten"] else (H, t["dim"]))
    return out

  def reset(self, env_ids=None) -> None:
    for terms in self.groups.values():
      for t in terms:
        if t["delay"] is not None:
          t["delay"].reset(env_ids)
        if t["history"] is not None:
          t["history"].reset(env_ids)

  def compute(self, raw: dict[str, torch.Tensor], update_history: bool = False) -> dict[str, torch.Tensor]:
    if not update_history and self._cache is not None:
      return self._cache
    out = {}
    for g, terms in self.groups.items():
      parts = []
      for t in terms:
        obs = raw[t["src"]][:, t["sl"]].clone()
        if t["clip"] is not None:
          obs = obs.clip_(min=t["clip"][0], max=t["clip"][1])
        if t["scale"] is not None:
          obs = obs.mul_(t["scale"])
        if t["delay"] is not None:
          t["delay"].append(obs)
          obs = t["delay"].compute().clone()
        if t["history"] is not None:
          h = t["history"]
          if update_history or not h.is_initialized:
            h.append(obs)
          obs = h.buffer.reshape(self.num_envs, -1) if t["flatten"] else h.buffer
        parts.append(obs)
      out[g] = torch.cat(parts, dim=-1)
    self._cache = out
    return out
