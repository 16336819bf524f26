"""Kernel-level profile of one PPO update (torch.profiler, eager launches) on the GPU box:  python tools/profile_ppo_update.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
from mjlab_b200.ppo import OnPolicyRunner, PpoCfg
from mjlab_b200.vecenv import RslRlVecEnvWrapper
env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=4096, seed=1), "cuda:0")
r = OnPolicyRunner(RslRlVecEnvWrapper(env), PpoCfg(update_cuda_graph=False))
r.learn(3)
ro = r.collect()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  r.update(ro)
  torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.device_time_total for k in ka)
print("total device time ms", tot / 1e3, "kernels", sum(k.count for k in ka))
for k in sorted(ka, key=lambda k: -k.device_time_total)[:22]:
  print(f"{k.device_time_total/1e3:8.2f} ms {k.count:5d}x  {k.key[:100]}")
