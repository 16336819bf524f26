"""Per-world timeline of k_solve (diagnostics build, GPU box):

  nvcc ... -DB2_TIMING -o build/variants/lib_timing.so mjlab_b200/csrc/b2_api.cu
  B200SIM_LIB=build/variants/lib_timing.so python tools/solve_timeline.py [num_envs]

Prints the distribution of per-world solver latency, the kernel span, and how many warps are resident over time: is the
kernel bounded by throughput or by the latency of its heaviest worlds?"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200 import engine as E  # noqa: E402
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
FUSED = len(sys.argv) > 2 and sys.argv[2] == "fused"   # time k_solve with the sensor / integration stage in its tail
env = ManagerBasedRlEnv(VelocityEnvCfg(robot="g1", num_envs=nw, seed=42), "cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(0)
env.reset()
env.episode_length_buf = torch.randint(0, 1000, (nw,), device="cuda:0", generator=gen)
for _ in range(40):
  env.step(torch.rand(nw, env.vcfg.nj, device="cuda:0", generator=gen) * 2 - 1)
lib, pm, pd = env._lib, C.byref(env.sim._cm), C.byref(env.sim._cd)
lib.b2_debug_timing.argtypes = [C.c_void_p, C.c_int]
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
  e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
  e0.record()
  E.check_call(lib, lib.b2_stage(pm, pd, 0, 1, st))
  e1.record()
  torch.cuda.synchronize()
  E.check_call(lib, lib.b2_stage(pm, pd, 3 if FUSED else 1, 1, st))
  e2.record()
  torch.cuda.synchronize()
  print(f"rep {rep}: CUDA events: k_smooth+k_order {e0.elapsed_time(e1) * 1e3:.0f} us, k_solve{' (+fused sensor/integration stage)' if FUSED else ''} {e1.elapsed_time(e2) * 1e3:.0f} us")
  if hasattr(lib, "b2_debug_smooth_phases"):
    ph = np.zeros(8 * nw, np.uint64)
    lib.b2_debug_smooth_phases.argtypes = [C.c_void_p, C.c_int]
    lib.b2_debug_smooth_phases(ph.ctypes.data, nw)
    ph = ph.reshape(nw, 8).astype(np.float64)
    names = ["setup+kinematics", "subtree_com+cinert+cdof", "crb+qM", "collision", "com_vel+actuation", "rne", "qfrc_smooth", "factor+solve"]
    print("   k_smooth phase shares: " + ", ".join(f"{n} {100 * ph[:, k].sum() / ph.sum():.0f}%" for k, n in enumerate(names)), f"| cycles/world {ph.sum(1).mean():.0f}")
  out = np.zeros(3 * nw, np.uint64)
  assert lib.b2_debug_timing(out.ctypes.data, nw) == 0
  t = out.reshape(nw, 3)
  t0, t1, sm = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64), t[:, 2].astype(np.int64)
  base = t0.min()
  dur = (t1 - t0) / 1e3
  span = (t1.max() - base) / 1e3
  nefc = env.sim.data.nefc.cpu().numpy()
  nit = env.sim.data.solver_niter.cpu().numpy()
  print(f"rep {rep}: kernel span {span:.0f} us; per-world latency us: mean {dur.mean():.0f} p50 {np.median(dur):.0f} p90 {np.quantile(dur, .9):.0f} "
        f"p99 {np.quantile(dur, .99):.0f} max {dur.max():.0f}; sum/span = {dur.sum() / span:.0f} warps busy on average "
        f"(of {148 * 16} slots)")
  # resident warps over time
  grid = np.linspace(0, span, 21)
  s0, s1 = (t0 - base) / 1e3, (t1 - base) / 1e3
  print("   resident worlds at 5% marks:", [int(((s0 <= g) & (s1 > g)).sum()) for g in grid[:-1]])
  heavy = np.argsort(-dur)[:5]
  print("   slowest worlds: ", [(int(w), round(float(dur[w])), int(nefc[w]), int(nit[w]), round(float(s0[w]))) for w in heavy], "(world, us, nefc, niter, start us)")
  last = np.argsort(-s1)[:5]
  print("   last to finish: ", [(int(w), round(float(dur[w])), int(nefc[w]), int(nit[w]), round(float(s0[w]))) for w in last])
  c = np.corrcoef(dur, nefc * np.maximum(nit, 1))[0, 1]
  print(f"   corr(latency, nefc*niter) = {c:.2f}; latency by niter:", {int(k): round(float(dur[nit == k].mean())) for k in np.unique(nit)})
  if hasattr(lib, "b2_debug_phases"):
    ph = np.zeros(8 * nw, np.uint64)
    lib.b2_debug_phases.argtypes = [C.c_void_p, C.c_int]
    lib.b2_debug_phases(ph.ctypes.data, nw)
    ph = ph.reshape(nw, 8).astype(np.float64)
    names = ["rows+setup", "solve_init", "update(cost,grad,JTf)", "hessian JTDJ", "factor+solve", "ls_setup(Mv,Jv)", "linesearch", "tail"]
    for label, sel in (("all worlds", np.arange(nw)), ("slowest 2 %", np.argsort(-dur)[: nw // 50])):
      tot = ph[sel].sum()
      print(f"   phase shares, {label}: " + ", ".join(f"{n} {100 * ph[sel, k].sum() / tot:.0f}%" for k, n in enumerate(names)),
            f"| cycles/world {ph[sel].sum(1).mean():.0f}")
  if hasattr(lib, "b2_debug_lsstat") and rep == 2:
    ls = np.zeros(8, np.uint64)
    lib.b2_debug_lsstat.argtypes = [C.c_void_p]
    lib.b2_debug_lsstat(ls.ctypes.data)
    for label, o in (("worlds with > 60 rows", 0), ("other worlds", 4)):
      n = max(int(ls[o]), 1)
      print(f"   line searches ({label}, all launches so far): {int(ls[o])}; per search: {ls[o + 1] / n:.2f} single evaluations, "
            f"{ls[o + 2] / n:.2f} three-point rounds; {100 * ls[o + 3] / n:.0f} % reach the refinement phase")
  if not FUSED:
    E.check_call(lib, lib.b2_stage(pm, pd, 2, 1, st))
