"""CPU experiment: how does the rounding floor of the Newton convergence test (B2_FLOOR_MULT x eps x |cost|, float32 builds
only) trade iterations against accuracy?  Realistic heavy states (fallen G1 robots under random controls, produced by the
float64 host build), one forward pass per variant, error of qacc against the float64 solution."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import mjlab_b200  # noqa: E402
from oracle.cport import CpuSim  # noqa: E402
from tests.hostemu import harness as H  # noqa: E402

m = mjlab_b200.load_model("g1_flat")
nw = 256
rng = np.random.default_rng(0)
sim = CpuSim(m, nw, 35, 300)
sim.qpos[:] = m.key_qpos[0]
sim.qpos[:, 2] += 0.02
for k in range(30):
  sim.ctrl[:] = m.key_ctrl[0] + rng.uniform(-1, 1, (nw, m.nu)) * 1.0
  sim.step(4, 8)
state = {f: getattr(sim, f).copy() for f in ("qpos", "qvel", "ctrl", "qacc_warmstart")}
print("states: ncon mean %.1f max %d, nefc mean %.1f max %d" % (sim.ncon.mean(), sim.ncon.max(), sim.nefc.mean(), sim.nefc.max()))


def run(real, mult, ls_mult=16, nstep=0, grad_mult=0):
  out = ROOT / "build" / f"hostemu_{real}_{mult}_{ls_mult}_{grad_mult}.so"
  out.parent.mkdir(exist_ok=True)
  subprocess.run(["g++", "-O2", "-std=c++17", "-DB2_HOSTEMU", f"-DB2_REAL={real}", f"-DB2_FLOOR_MULT={mult}", f"-DB2_LS_FLOOR_MULT={ls_mult}", f"-DB2_GRAD_FLOOR_MULT={grad_mult}", "-shared", "-fPIC", "-ffp-contract=off",
                  "-Wno-unused-variable", "-o", str(out), str(ROOT / "tests" / "hostemu" / "hostemu.cpp")], check=True)
  orig = H.build
  H.build = lambda r="float": out
  try:
    hs = H.HostSim(m, nw, 35, 300, real=real)
  finally:
    H.build = orig
  for f, v in state.items():
    getattr(hs, f)[:] = v
  hs.lib.b2h_ls_pass_count.restype = C.c_longlong
  n0 = hs.lib.b2h_ls_pass_count()
  hs.forward()
  run.passes = hs.lib.b2h_ls_pass_count() - n0
  qacc = hs.qacc.astype(np.float64).copy()
  for _ in range(nstep):
    hs.step()
  run.qpos = hs.qpos.astype(np.float64).copy()
  return qacc, hs.solver_niter.copy(), hs.nefc.copy()


ref, nit64, nefc = run("double", 8)
print("float64: niter mean %.2f max %d" % (nit64.mean(), nit64.max()))
heavy = nefc > 60
for mult in (8, 32, 128, 512, 2048):
  q, nit, _ = run("float", mult)
  err = np.abs(q - ref).max(1) / np.maximum(np.abs(ref).max(1), 1.0)
  print(f"float32 floor x{mult:5d}: niter mean {nit.mean():.2f} (heavy worlds {nit[heavy].mean():.2f}, n={heavy.sum()}), max {nit.max()}; "
        f"qacc rel err median {np.median(err):.2e} p90 {np.quantile(err, .9):.2e} max {err.max():.2e}")

print("--- line-search derivative floor (float32), Newton floor x8")
ref, nit64, nefc = run("double", 8, 0, nstep=20)
qpos64 = run.qpos
print("float64: line-search row passes %d" % run.passes)
for ls_mult in (0, 4, 16, 64, 256):
  q, nit, _ = run("float", 8, ls_mult, nstep=20)
  err = np.abs(q - ref).max(1) / np.maximum(np.abs(ref).max(1), 1.0)
  eq = np.abs(run.qpos - qpos64).max(1)
  print(f"float32 ls floor x{ls_mult:4d}: row passes {run.passes} (forward pass), niter mean {nit.mean():.2f}; qacc rel err median {np.median(err):.2e} "
        f"p90 {np.quantile(err, .9):.2e} max {err.max():.2e}; qpos err after 20 steps median {np.median(eq):.2e} p90 {np.quantile(eq, .9):.2e}")

print("--- gradient-noise floor of the Newton convergence test (float32), line-search floor x16")
for gm in (0, 2, 8, 32, 128):
  q, nit, _ = run("float", 8, 16, nstep=20, grad_mult=gm)
  err = np.abs(q - ref).max(1) / np.maximum(np.abs(ref).max(1), 1.0)
  eq = np.abs(run.qpos - qpos64).max(1)
  print(f"float32 grad floor x{gm:4d}: niter mean {nit.mean():.2f} (heavy {nit[heavy].mean():.2f}); qacc rel err median {np.median(err):.2e} "
        f"p90 {np.quantile(err, .9):.2e} max {err.max():.2e}; qpos err after 20 steps median {np.median(eq):.2e} p90 {np.quantile(eq, .9):.2e}")
