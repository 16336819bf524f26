"""Timing of b2_env_substeps (4 physics sub-steps) under different chunk counts, plus the solver's per-world timeline of the
LAST sub-step (diagnostics build):  B200SIM_LIB=build/variants/lib_timing.so python tools/chunk_timeline.py"""
import ctypes as C
import os
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
lib, pm, pd, pe = env._lib, C.byref(env.sim._cm), C.byref(env.sim._cd), C.byref(env._ce)
has_tim = hasattr(lib, "b2_debug_timing")
if has_tim:
  lib.b2_debug_timing.argtypes = [C.c_void_p, C.c_int]
st = torch.cuda.current_stream().cuda_stream
snap = {k: getattr(env.sim.data, k).clone() for k in ("qpos", "qvel", "qacc_warmstart", "ctrl", "time")}
for chunks in (1, 2, 4, 8, 1, 4):
  os.environ["B200SIM_CHUNKS"] = str(chunks)
  ts = []
  for rep in range(6):
    for k, v in snap.items():
      getattr(env.sim.data, k).copy_(v)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    E.check_call(lib, lib.b2_env_substeps(pm, pd, pe, 4, st))
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  msg = f"chunks {chunks}: 4 sub-steps {np.median(ts):.3f} ms (min {min(ts):.3f})"
  if has_tim:
    out = np.zeros(3 * nw, np.uint64)
    lib.b2_debug_timing(out.ctypes.data, nw)
    t = out.reshape(nw, 3).astype(np.int64)
    dur = (t[:, 1] - t[:, 0]) / 1e3
    span = (t[:, 1].max() - t[:, 0].min()) / 1e3
    msg += f"; last solve: span over all chunks {span:.0f} us, per-world latency mean {dur.mean():.0f} p50 {np.median(dur):.0f} p99 {np.quantile(dur, .99):.0f} max {dur.max():.0f}"
    per = nw // chunks
    msg += " | chunk spans " + " ".join(f"{(t[c*per:(c+1)*per,1].max()-t[c*per:(c+1)*per,0].min())/1e3:.0f}@{(t[c*per:(c+1)*per,0].min()-t[:,0].min())/1e3:.0f}" for c in range(chunks))
  print(msg)
