"""Per-kernel timing of the physics step on realistic env states (GPU box).

  B200SIM_LIB=build/variants/lib_x.so B200SIM_JPOOL=640 python tools/time_solve.py [robot] [num_envs]

Runs the velocity env with random actions for a while (so the batch holds a realistic mix of standing / falling /
reset worlds), then brackets k_smooth and k_solve of 40 physics steps with CUDA events (256 MiB L2 flush between steps).
"""
import ctypes as C
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200 import engine as E  # noqa: E402
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg  # noqa: E402

robot = sys.argv[1] if len(sys.argv) > 1 else "g1"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = ManagerBasedRlEnv(VelocityEnvCfg(robot=robot, num_envs=nw, seed=42), "cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(0)
env.reset()
env.episode_length_buf = torch.randint(0, 1000, (nw,), device="cuda:0", generator=gen)
for _ in range(40):
  env.step(torch.rand(nw, env.vcfg.nj, device="cuda:0", generator=gen) * 2 - 1)
lib, pm, pd = env._lib, C.byref(env.sim._cm), C.byref(env.sim._cd)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda:0")
ts, tv, tp = [], [], []
for k in range(40):
  if k % 4 == 0:
    env._buf["action_in"].copy_(torch.rand(nw, env.vcfg.nj, device="cuda:0", generator=gen) * 2 - 1)
    E.check_call(lib, lib.b2_env_action(pm, pd, C.byref(env._ce), torch.cuda.current_stream().cuda_stream))
  flush.zero_()
  a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
  st = torch.cuda.current_stream().cuda_stream
  a.record()
  E.check_call(lib, lib.b2_stage(pm, pd, 0, 1, st))
  b.record()
  E.check_call(lib, lib.b2_stage(pm, pd, 1, 1, st))
  c.record()
  E.check_call(lib, lib.b2_stage(pm, pd, 2, 1, st))
  c2 = torch.cuda.Event(enable_timing=True)
  c2.record()
  torch.cuda.synchronize()
  ts.append(a.elapsed_time(b))
  tv.append(b.elapsed_time(c))
  tp.append(c.elapsed_time(c2))
d = env.sim.data
print(f"{os.environ.get('B200SIM_LIB', 'default'):40s} jpool={os.environ.get('B200SIM_JPOOL', '-'):>5s} k_smooth {statistics.mean(ts):.4f} ms  "
      f"k_solve {statistics.mean(tv):.4f} ms  (min {min(tv):.4f}) k_post {statistics.mean(tp):.4f}  nefc {d.nefc.float().mean().item():.1f} niter {d.solver_niter.float().mean().item():.2f} "
      f"dropped {int(d.nefc_dropped.sum().item())}")
ne = d.nefc.float()
q = torch.tensor([0.5, 0.75, 0.9, 0.95, 0.99, 1.0], device=ne.device)
print("nefc quantiles 50/75/90/95/99/100:", torch.quantile(ne, q).tolist(), " frac>64:", (ne > 64).float().mean().item(), " frac>96:", (ne > 96).float().mean().item(),
      " ncon q:", torch.quantile(d.ncon.float(), q).tolist(), " niter q:", torch.quantile(d.solver_niter.float(), q).tolist())
