"""How much of a heavy world's solver latency is contention?  (diagnostics build, GPU box)

  nvcc ... -DB2_TIMING -o build/variants/lib_timing.so mjlab_b200/csrc/b2_api.cu
  B200SIM_LIB=build/variants/lib_timing.so python tools/solve_isolation.py

Solves the SAME heaviest worlds of a realistic 4096-world batch (a) inside the full launch and (b) alone, k per SM, through
the masked forward pass; prints their per-world k_solve latency in each setting."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200 import engine as E  # noqa: E402
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg  # noqa: E402

nw = 4096
env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=nw, seed=42), "cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(0)
env.reset()
env.episode_length_buf = torch.randint(0, 1000, (nw,), device="cuda:0", generator=gen)
for _ in range(40):
  env.step(torch.rand(nw, env.vcfg.nj, device="cuda:0", generator=gen) * 2 - 1)
lib, pm, pd = env._lib, C.byref(env.sim._cm), C.byref(env.sim._cd)
lib.b2_debug_timing.argtypes = [C.c_void_p, C.c_int]
st = torch.cuda.current_stream().cuda_stream
one = torch.ones(1, dtype=torch.int32, device="cuda:0")


def latencies(mask):
  ptr = None if mask is None else mask.data_ptr()
  E.check_call(lib, lib.b2_forward_masked(pm, pd, ptr, one.data_ptr(), st))
  torch.cuda.synchronize()
  out = np.zeros(3 * nw, np.uint64)
  assert lib.b2_debug_timing(out.ctypes.data, nw) == 0
  t = out.reshape(nw, 3).astype(np.int64)
  return (t[:, 1] - t[:, 0]) / 1e3, (t[:, 1].max() - t[:, 0].min()) / 1e3


full, span = latencies(None)
full, span = latencies(None)
nit = env.sim.data.solver_niter.cpu().numpy()
order = np.argsort(-full)
print(f"full launch: span {span:.0f} us, mean latency {full.mean():.0f} us")
for k in (1, 2, 4, 8, 16):
  sel = order[: 148 * k]
  mask = torch.zeros(nw, dtype=torch.uint8, device="cuda:0")
  mask[torch.as_tensor(sel, device="cuda:0")] = 1
  alone, sp = latencies(mask)
  alone, sp = latencies(mask)
  top = order[:148]
  print(f"{148 * k:5d} heaviest worlds alone ({k:2d} per SM): span {sp:.0f} us; the 148 heaviest: {alone[top].mean():.0f} us alone vs {full[top].mean():.0f} us in the full launch "
        f"(niter {nit[top].mean():.1f}); all selected: {alone[sel].mean():.0f} vs {full[sel].mean():.0f}")
