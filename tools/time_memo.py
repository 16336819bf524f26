"""Timing of k_smooth's two paths on realistic env states (GPU box): the full position / velocity stage (b2_stage 4) against
the pass that reuses it (b2_stage 6, what the first sub-step after a forward pass runs), plus the share of worlds whose
snapshot matches after real env steps.

  python tools/time_memo.py [robot] [num_envs]
"""
import ctypes as C
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
match = []
d = env.sim.data
for _ in range(40):
  env.step(torch.rand(nw, env.vcfg.nj, device="cuda:0", generator=gen) * 2 - 1)
  same = (d.memo_qpos == d.qpos).all(1) & (d.memo_qvel == d.qvel).all(1) & (d.smooth_memo == 1)
  match.append(float(same.float().mean()))
print("share of worlds whose snapshot matches after an env step (min / mean):", min(match), statistics.mean(match))
lib, pm, pd = env._lib, C.byref(env.sim._cm), C.byref(env.sim._cd)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda:0")
res = {4: [], 6: []}
for k in range(60):
  stage = 4 if k % 2 == 0 else 6     # 6 right after 4 on the same state: every world hits
  flush.zero_()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  st = torch.cuda.current_stream().cuda_stream
  a.record()
  E.check_call(lib, lib.b2_stage(pm, pd, stage, 1, st))
  b.record()
  torch.cuda.synchronize()
  res[stage].append(a.elapsed_time(b))
print(f"k_smooth full {statistics.mean(res[4][3:]):.4f} ms   reuse {statistics.mean(res[6][3:]):.4f} ms  (min {min(res[6]):.4f})")
