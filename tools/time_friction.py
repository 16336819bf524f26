"""Cost of the friction-loss rows on the headline workload (G1 flat, 4096 envs, full env step): the same env timed with
dof_frictionloss = 0 (friction-free solver kernels) and with a friction loss on every actuated joint (k_solve_t<*, true>,
29 extra rows per world).  CUDA events around graph replays, random actions; prints one JSON line.
    python tools/time_friction.py [frictionloss] [num_envs]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg  # noqa: E402


def run(fl: float, n: int, steps: int = 200, warm: int = 300) -> dict:
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=n, seed=42), "cuda:0")
  if fl > 0:
    env.sim.expand_model_fields(("dof_frictionloss",))
    env.sim.model.dof_frictionloss[:, 6:] = fl
  env.reset()
  g = torch.Generator(device="cuda:0").manual_seed(0)
  acts = [torch.rand(n, env.vcfg.nj, device="cuda:0", generator=g) * 2 - 1 for _ in range(16)]
  for k in range(warm):
    env.step(acts[k % 16])
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for k in range(steps):
    env.step(acts[k % 16])
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / steps
  D = env.sim.data
  return dict(frictionloss=fl, num_envs=n, ms_per_env_step=ms, env_steps_per_s=n / ms * 1e3, mean_nf=float(D.nf.float().mean()),
              mean_nefc=float(D.nefc.float().mean()), mean_niter=float(D.solver_niter.float().mean()))


if __name__ == "__main__":
  fl = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
  n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
  print(json.dumps(dict(workload="g1_flat velocity env step, random actions, no L2 flush (back-to-back replays)", off=run(0.0, n), on=run(fl, n))))
