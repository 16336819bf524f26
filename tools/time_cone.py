"""Cost of the elliptic friction cone on the headline workload (G1 flat, 4096 envs, full env step): the same env timed on the
pyramidal cone (k_solve, the shipped configuration) and on the elliptic cone (k_solve_t<false, false, true>: 3 rows per
frictional contact instead of 4, per-contact cone passes, a factorisation every Newton iteration).  CUDA events around graph
replays, random actions; prints one JSON line.
    python tools/time_cone.py [num_envs]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg  # noqa: E402
from mjlab_b200.sim import MujocoCfg, SimulationCfg  # noqa: E402


def run(cone: str, n: int, steps: int = 200, warm: int = 300) -> dict:
  sim = SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20, cone=cone))
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=n, seed=42, sim=sim), "cuda:0")
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
  return dict(cone=cone, num_envs=n, ms_per_env_step=ms, env_steps_per_s=n / ms * 1e3, mean_nefc=float(D.nefc.float().mean()),
              mean_ncon=float(D.ncon.float().mean()), mean_niter=float(D.solver_niter.float().mean()))


if __name__ == "__main__":
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
  if len(sys.argv) > 2:   # one cone only, few steps: the ncu capture (`ncu -k regex:k_solve -s 100 -c 1 python tools/time_cone.py 4096 elliptic`)
    print(json.dumps(run(sys.argv[2], n, steps=5, warm=30)))
    sys.exit(0)
  print(json.dumps(dict(workload="g1_flat velocity env step, random actions, no L2 flush (back-to-back replays)",
                        pyramidal=run("pyramidal", n), elliptic=run("elliptic", n))))
