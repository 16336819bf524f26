"""bench.py - env-steps/s of the G1 flat-velocity hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # own arm (CUDA engine)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference algorithm's CPU restatement

One "step" is one env step of the reference's hot loop
(src/mjlab/envs/manager_based_rl_env.py:272-313): process a fresh action batch,
then `decimation` = 4 physics steps, each preceded by the ctrl write of the
builtin position actuators (src/mjlab/actuator/builtin_group.py:68-81).
`config.workload` says exactly which parts of the env step are inside the timed
region.

Timing: CUDA events on the launching stream around every env step, an L2 flush
(256 MiB memset) between timed steps, >= 3 warm-up steps, barrier + synchronize
on both sides, max over ranks.  Roofline: a second pass replays the same steps
with the two stages launched separately (b2_stage) and CUDA events around each
k_solve launch; see DESIGN.md for the algorithmic-bytes definition.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DECIMATION = 4
NUM_ENVS = 4096
EPISODE_STEPS = 250  # staggered synthetic resets keep a mix of standing / falling worlds


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=100)
  ap.add_argument("--warmup", type=int, default=20)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--model", default="g1_flat")
  ap.add_argument("--num-envs", type=int, default=NUM_ENVS)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-roofline-pass", action="store_true")
  return ap.parse_args()


def action_scale(m) -> np.ndarray:
  """0.25 * effort / stiffness per actuator (g1_constants.py:287-295)."""
  kp = m.actuator_gainprm[:, 0]
  eff = m.actuator_forcerange[:, 1]
  return 0.25 * eff / kp


def algorithmic_bytes(m):
  """bytes that must cross HBM per world per physics step, split by stage."""
  nq, nv, nu, nb, ng, ns, nsd = m.nq, m.nv, m.nu, m.nbody, m.ngeom, m.nsite, m.nsensordata
  smooth_r = nq + nv + nu + 6 * nb
  smooth_w = nb * (3 + 4 + 3 + 3 + 6) + ng * 12 + ns * 12 + nu + nv
  solve_r = nv + 1
  solve_w = nq + nv + nv + nv + 1 + nsd
  return 4 * (smooth_r + smooth_w), 4 * (solve_r + solve_w)


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class Clocks:
  Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index: int):
    self.samples: list[list[str]] = []
    self.proc = None
    self.index = index

  def start(self):
    try:
      self.proc = subprocess.Popen(
        ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    assert self.proc is not None and self.proc.stdout is not None
    for line in self.proc.stdout:
      parts = [p.strip() for p in line.split(",")]
      if len(parts) >= 6:
        self.samples.append(parts)

  def stop(self) -> dict:
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for p in self.samples:
      try:
        sm.append(float(p[0])); mx.append(float(p[1]))
      except ValueError:
        continue
      for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
        if v.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port stepped on the host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
  name, seed, nworld, nsteps = args
  import mjlab_b200
  from oracle import mjstep as O
  m = mjlab_b200.load_model(name)
  rng = np.random.default_rng(seed)
  scale = action_scale(m)
  t0 = time.perf_counter()
  for _ in range(nworld):
    d = O.Data(m)
    d.qpos[:] = m.key_qpos[0]
    for _ in range(nsteps):
      act = rng.uniform(-1, 1, m.nu)
      d.ctrl[:] = m.key_ctrl[0] + scale * act
      for _ in range(DECIMATION):
        O.step(m, d, nconmax=35, njmax=300)
  return time.perf_counter() - t0


def cpu_env_steps_per_s(name: str, cores: int, worlds_per_core: int, nsteps: int) -> tuple[float, float]:
  """env-steps/s of the oracle port with `cores` processes; returns (value, seconds)."""
  import multiprocessing as mp
  jobs = [(name, 1000 + i, worlds_per_core, nsteps) for i in range(cores)]
  t0 = time.perf_counter()
  if cores == 1:
    _cpu_worker(jobs[0])
  else:
    with mp.get_context("fork").Pool(cores) as pool:
      pool.map(_cpu_worker, jobs)
  dt = time.perf_counter() - t0
  return cores * worlds_per_core * nsteps / dt, dt


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import mjlab_b200
  m = mjlab_b200.load_model(args.model)
  cores = os.cpu_count() or 1
  # bounded sample: each "step" of this arm is one env step of `cores` worlds (one per core)
  k, w = args.steps, args.warmup
  per_step_worlds = cores
  nsteps_sample = max(1, min(k, 3))
  value, secs = cpu_env_steps_per_s(args.model, cores, 1, nsteps_sample)
  ms = 1e3 * per_step_worlds / value
  line = {
    "impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
    "steps": k, "warmup": w, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f64", "data": "synthetic",
    "config": {"workload": workload_name(args.model, args.num_envs), "num_envs_per_gpu": args.num_envs,
               "decimation": DECIMATION, "timestep": 0.005},
    "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                     "sample": f"{cores} worlds x {nsteps_sample} env steps (x{DECIMATION} physics steps), one world per "
                               f"process, numpy float64 oracle/mjstep.py; {secs:.1f} s"},
    "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "note": "mujoco / mujoco_warp are not installable on this image: the CPU arm is the oracle port of the same algorithm",
  }
  print(json.dumps(line))


def workload_name(model: str, num_envs: int) -> str:
  return (f"{model} velocity env step, {num_envs} envs/GPU: action->ctrl (builtin position actuators) + "
          f"{DECIMATION} physics steps (FK, CRB, collision, Newton solve, sensors, implicitfast) + staggered resets; "
          "observation/reward/termination managers NOT yet inside the step (round 1)")


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
  import torch
  import torch.distributed as dist

  import mjlab_b200
  from mjlab_b200 import engine as E
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  import ctypes as C

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = f"cuda:{local}"
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))

  m = mjlab_b200.load_model(args.model)
  nw = args.num_envs
  cfg = SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20))
  sim = Simulation(nw, cfg, m, dev)
  torch.manual_seed(42 + rank)  # seed + rank (src/mjlab/scripts/train.py:59)

  f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
  key_qpos, key_ctrl, scale = f32(m.key_qpos[0]), f32(m.key_ctrl[0]), f32(action_scale(m))
  phase = torch.arange(nw, device=dev) % EPISODE_STEPS
  counter = torch.zeros((), dtype=torch.long, device=dev)
  host_actions = torch.empty((nw, m.nu), dtype=torch.float32).pin_memory()
  dev_actions = torch.empty((nw, m.nu), dtype=torch.float32, device=dev)
  host_result = torch.empty((nw, m.nq + m.nv), dtype=torch.float32).pin_memory()
  dev_result = torch.empty((nw, m.nq + m.nv), dtype=torch.float32, device=dev)

  def reset_all():
    sim.reset()
    sim.data.qpos[:] = key_qpos
    sim.data.qpos[:, 0:2] += torch.rand((nw, 2), device=dev) - 0.5
    sim.data.ctrl[:] = key_ctrl
    counter.zero_()

  def staggered_reset():
    mask = (phase + counter) % EPISODE_STEPS == 0
    sim._reset_mask.copy_(mask)
    sim.reset_graph.replay()
    sim.data.qpos[:] = torch.where(mask[:, None], key_qpos, sim.data.qpos)
    sim.data.ctrl[:] = torch.where(mask[:, None], key_ctrl, sim.data.ctrl)
    counter.add_(1)

  def env_step(actions, split=None):
    """actions: device tensor [nw, nu] in [-1, 1]."""
    target = key_ctrl + scale * actions          # JointAction.process_actions (joint_actions.py:84-86)
    for _ in range(DECIMATION):
      sim.data.ctrl[:] = target                   # apply_action + write_data_to_sim
      if split is None:
        sim.step()
      else:
        split()
    staggered_reset()

  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  reset_all()
  for _ in range(max(args.warmup, 3)):
    env_step(torch.rand((nw, m.nu), device=dev) * 2 - 1)
  barrier()

  # ---- timed region: device-resident inputs ----------------------------------------------------
  clocks = Clocks(local)
  clocks.start()
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  acts = torch.rand((args.steps, nw, m.nu), device=dev) * 2 - 1
  barrier()
  t_wall0 = time.perf_counter()
  for k in range(args.steps):
    flush.zero_()
    ev[k][0].record()
    env_step(acts[k])
    ev[k][1].record()
  barrier()
  t_wall = time.perf_counter() - t_wall0
  elapsed_ms = sum(a.elapsed_time(b) for a, b in ev)

  # ---- e2e: host buffers in, host result out, copies inside the timed region --------------------
  ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  host_all = (torch.rand((args.steps, nw, m.nu)) * 2 - 1)
  barrier()
  for k in range(args.steps):
    host_actions.copy_(host_all[k])
    flush.zero_()
    ev2[k][0].record()
    dev_actions.copy_(host_actions, non_blocking=True)
    env_step(dev_actions)
    dev_result[:, : m.nq] = sim.data.qpos
    dev_result[:, m.nq:] = sim.data.qvel
    host_result.copy_(dev_result, non_blocking=True)
    ev2[k][1].record()
    ev2[k][1].synchronize()  # the caller consumes the result before issuing the next action
  barrier()
  e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
  clk = clocks.stop()

  # ---- roofline pass: stage-split launches, CUDA events around each k_solve ----------------------
  roof = None
  solve_ms = smooth_ms = None
  if not args.no_roofline_pass:
    lib, cm, cd = sim._lib, sim._cm, sim._cd
    pairs: list = []

    def split():
      st = torch.cuda.current_stream().cuda_stream
      a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
      a.record()
      E.check_call(lib, lib.b2_stage(C.byref(cm), C.byref(cd), 0, 1, st))
      b.record()
      E.check_call(lib, lib.b2_stage(C.byref(cm), C.byref(cd), 1, 1, st))
      c.record()
      pairs.append((a, b, c))

    nroof = min(args.steps, 25)
    barrier()
    for k in range(nroof):
      flush.zero_()
      env_step(acts[k], split=split)
    barrier()
    smooth_ms = statistics.mean(a.elapsed_time(b) for a, b, c in pairs)
    solve_ms = statistics.mean(b.elapsed_time(c) for a, b, c in pairs)
    b_smooth, b_solve = algorithmic_bytes(m)
    peaks = {}
    try:
      peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
      pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_bytes, dom_ms, dom = (b_solve, solve_ms, "k_solve") if solve_ms >= smooth_ms else (b_smooth, smooth_ms, "k_smooth")
    achieved = dom_bytes * nw / (dom_ms * 1e-3) / 1e9
    roof = {
      "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
      "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
      "algorithmic_bytes_per_world": {"k_smooth": b_smooth, "k_solve": b_solve},
      "avg_launch_ms": {"k_smooth": smooth_ms, "k_solve": solve_ms},
      "whole_step": {"bytes_per_world": b_smooth + b_solve,
                     "achieved": (b_smooth + b_solve) * nw / ((smooth_ms + solve_ms) * 1e-3) / 1e9},
    }

  # solver statistics of the final state (explains the latency-bound kernel)
  niter = sim.data.solver_niter.float().mean().item()
  nefc = sim.data.nefc.float().mean().item()
  ncon = sim.data.ncon.float().mean().item()

  # ---- max over ranks ------------------------------------------------------------------------------
  t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  elapsed_ms, e2e_ms = t.tolist()
  total_env_steps = args.steps * nw * world
  value = total_env_steps / (elapsed_ms * 1e-3)
  e2e_value = total_env_steps / (e2e_ms * 1e-3)

  if rank == 0:
    line = {
      "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
      "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(args.model, nw), "num_envs_per_gpu": nw, "decimation": DECIMATION,
                 "timestep": 0.005, "nconmax": 35, "njmax": 300, "parallelism": f"replicas x{world} (no physics collective)",
                 "l2": "256 MiB flush between timed steps", "actions": "U(-1,1) redrawn every env step",
                 "mean_solver_niter": niter, "mean_nefc": nefc, "mean_ncon": ncon},
      "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": nw * m.nu * 4,
              "d2h_bytes_per_step": nw * (m.nq + m.nv) * 4},
      "gpu_launches": args.steps * (DECIMATION * 2 + 1),
      "clocks": clk,
      "wall_s": t_wall,
    }
    if roof is not None:
      line["roofline"] = roof
    if world == 1 and not args.no_cpu_baseline:
      v, secs = cpu_env_steps_per_s(args.model, 1, 4, 2)
      line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                              "sample": f"4 worlds x 2 env steps (x{DECIMATION} physics steps) of the same workload, "
                                        f"numpy float64 oracle/mjstep.py, {secs:.1f} s"}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  a = parse()
  if a.impl == "reference":
    run_reference(a)
  else:
    run_b200(a)
