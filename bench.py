"""bench.py - env-steps/s of the G1 flat-velocity hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # own arm (CUDA engine)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference algorithm's CPU restatement

One "step" is one `ManagerBasedRlEnv.step` of the reference's hot loop
(src/mjlab/envs/manager_based_rl_env.py:272-313) over all envs of the rank: action processing,
`decimation` = 4 physics steps with the contact air-time update, terminations, rewards, resets,
the forward pass after resets, commands, push events and observations.  `config.workload`
spells out what is inside the timed region.

Timing: CUDA events on the launching stream around every env step, an L2 flush
(256 MiB memset) between timed steps, >= 3 warm-up steps, barrier + synchronize
on both sides, max over ranks.  Roofline: a second pass enqueues the same
step eagerly with CUDA events around every kernel launch; see DESIGN.md for the algorithmic-bytes
definition.
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
# DRAM bytes per launch from the committed ncu captures of this workload (profiles/r01z_k_solve_final.txt,
# r01e_k_smooth.txt, r01e_k_post.txt)
NCU_TRAFFIC_BYTES = {"k_solve": 38_369_280 + 515_584, "k_smooth": 12_324_096 + 32_529_664, "k_post": 52_356_608 + 761_600}


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=100)
  ap.add_argument("--warmup", type=int, default=20)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--model", default="g1_flat")
  ap.add_argument("--task", default="velocity", choices=["velocity", "tracking", "rough"],
                  help="velocity (BASELINE configs 2/3, the headline) or tracking (config 5: synthetic clip, domain randomisation on)")
  ap.add_argument("--num-envs", type=int, default=None, help="envs per GPU (default 4096; 8192 for --task rough, BASELINE config 4)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-roofline-pass", action="store_true")
  ap.add_argument("--forward-mode", default="reference", choices=["reference", "masked"],
                  help="forward pass after resets: all envs when any env reset (the reference's behaviour) or only the reset envs")
  args = ap.parse_args()
  if args.task == "rough":
    args.model = "g1_rough"
  if args.num_envs is None:
    args.num_envs = 8192 if args.task == "rough" else NUM_ENVS
  return args


def action_scale(m) -> np.ndarray:
  """0.25 * effort / stiffness per actuator (g1_constants.py:287-295)."""
  kp = m.actuator_gainprm[:, 0]
  eff = m.actuator_forcerange[:, 1]
  return 0.25 * eff / kp


def algorithmic_bytes(m):
  """bytes that must cross HBM per world per physics step, split by stage."""
  nq, nv, nu, nb, ns, nsd = m.nq, m.nv, m.nu, m.nbody, m.nsite, m.nsensordata
  ng = m.ngeom if m.nfield == 0 else m.ndyn   # terrain-field geoms are static: their pose is never rewritten
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
  """The same env step on the host: oracle physics (one world at a time) + oracle env stage (batched over the worlds)."""
  name, seed, nworld, nsteps = args
  from types import SimpleNamespace as NS

  import mjlab_b200
  from mjlab_b200 import envcfg
  from oracle import envstep as OE
  from oracle import mjstep as O
  m = mjlab_b200.load_model(name)
  cfg = envcfg.velocity_cfg(m, name.split("_")[0])
  cfg.seed = seed
  rng = np.random.default_rng(seed)
  fields = ["qpos", "qvel", "ctrl", "xpos", "xquat", "subtree_com", "cvel", "site_xpos", "sensordata", "qacc_warmstart", "qacc"]
  ds = [O.Data(m) for _ in range(nworld)]
  for d in ds:
    d.qpos[:] = m.key_qpos[0]
    O.forward(m, d, 35, 300)

  def batch():
    b = NS(**{f: np.stack([np.asarray(getattr(x, f), np.float64) for x in ds]) for f in fields})
    b.time = np.array([x.time for x in ds])
    return b

  def scatter(b):
    for w, x in enumerate(ds):
      x.qpos[:], x.qvel[:], x.ctrl[:], x.qacc_warmstart[:] = b.qpos[w], b.qvel[w], b.ctrl[w], b.qacc_warmstart[w]
      x.time = float(b.time[w])

  s = OE.make_state(cfg, nworld)
  t0 = time.perf_counter()
  for _ in range(nsteps):
    b = batch()
    OE.apply_action(cfg, s, b, rng.uniform(-1, 1, (nworld, cfg.nj)))
    scatter(b)
    for _ in range(DECIMATION):
      for d in ds:
        O.step(m, d, nconmax=35, njmax=300)
      b = batch()
      OE.update_air_time(cfg, s, b.sensordata.astype(np.float32), b.time)
    OE.post(cfg, m, s, b)
    scatter(b)
    if s.reset_buf.any():
      for d in ds:
        O.forward(m, d, 35, 300)
      b = batch()
    OE.obs_stage(cfg, m, s, b)
    scatter(b)
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


def cpu_compiled_env_steps_per_s(name: str, cores: int, target_s: float = 12.0):
  """env-steps/s of the compiled CPU build (oracle/cport: the same algorithm, float64, one thread per core) on a bounded sample:
  per env step one action -> ctrl write, `decimation` physics steps and the numpy env stage (oracle/envstep.py) on the batch."""
  from types import SimpleNamespace as NS

  import mjlab_b200
  from mjlab_b200 import envcfg
  from oracle import envstep as OE
  from oracle.cport import CpuSim
  m = mjlab_b200.load_model(name)
  m.opt_timestep, m.opt_iterations, m.opt_ls_iterations = 0.005, 10, 20
  cfg = envcfg.velocity_cfg(m, name.split("_")[0])
  nworld = 64 * cores
  sim = CpuSim(m, nworld, 35, 300)
  sim.qpos[:] = m.key_qpos[0]
  rng = np.random.default_rng(7)
  s = OE.make_state(cfg, nworld)
  fields = ("qpos", "qvel", "ctrl", "xpos", "xquat", "subtree_com", "cvel", "site_xpos", "sensordata", "qacc_warmstart", "qacc")

  def view():
    d = NS(**{f: getattr(sim, f) for f in fields})
    d.time = sim.time.reshape(-1)
    return d

  def env_step():
    d = view()
    OE.apply_action(cfg, s, d, rng.uniform(-1, 1, (nworld, cfg.nj)))
    for _ in range(DECIMATION):
      sim.step(1, cores)
      OE.update_air_time(cfg, s, sim.sensordata.astype(np.float32), d.time)
    OE.post(cfg, m, s, d)
    OE.obs_stage(cfg, m, s, d)

  env_step()  # page in the buffers
  t0 = time.perf_counter()
  env_step()
  one = time.perf_counter() - t0
  nsteps = int(max(3, min(200, target_s / max(one, 1e-3))))
  t0 = time.perf_counter()
  for _ in range(nsteps):
    env_step()
  secs = time.perf_counter() - t0
  desc = (f"{nworld} worlds x {nsteps} env steps (x{DECIMATION} physics steps + numpy env stage; the forward pass after resets is not "
          f"included), compiled CPU build of the same algorithm (oracle/cport, g++ -O3, float64), {cores} threads, {secs:.1f} s")
  return nworld * nsteps / secs, secs, desc


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  k, w = args.steps, args.warmup
  value, secs, desc = cpu_compiled_env_steps_per_s(args.model, cores, target_s=20.0)
  ms = 1e3 * 64 * cores / value   # one "step" of this arm = one env step of its 64 x cores worlds
  line = {
    "impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
    "steps": k, "warmup": w, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f64", "data": "synthetic",
    "config": {"workload": workload_name(args.model, args.num_envs, args.forward_mode), "num_envs_per_gpu": args.num_envs,
               "decimation": DECIMATION, "timestep": 0.005},
    "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": desc},
    "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "note": "mujoco / mujoco_warp are not installable on this image: the CPU arm is a compiled all-cores CPU build of the same algorithm",
  }
  print(json.dumps(line))


def workload_name(model: str, num_envs: int, forward_mode: str = "reference", task: str = "velocity") -> str:
  if task == "tracking":
    return (f"{model} motion-tracking task (synthetic 500-frame clip, domain randomisation on), env step, {num_envs} envs/GPU: action -> ctrl, "
            f"{DECIMATION} physics steps, 4 terminations, 9 reward terms, resets with adaptive start-frame sampling, forward after resets "
            f"({forward_mode}), frame advance + anchor-relative targets, push events, observations (policy 160 + critic 286, noise) - one CUDA graph")
  if task == "rough":
    model += " (3564-box ROUGH_TERRAINS_CFG grid, terrain seed 0, terrain-level curriculum)"
  return (f"{model} velocity task, ManagerBasedRlEnv.step, {num_envs} envs/GPU: action -> ctrl, {DECIMATION} x (physics step + "
          "contact air-time update), terminations, 14 reward terms, resets (sim + events + managers), "
          f"forward after resets ({forward_mode}), commands, push events, observations (policy+critic, noise) - one CUDA graph")


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
  import ctypes as C

  import torch
  import torch.distributed as dist

  from mjlab_b200 import engine as E
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = f"cuda:{local}"
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))

  nw = args.num_envs
  robot = args.model.split("_")[0]
  # one independent world batch per rank, seed + rank (src/mjlab/scripts/train.py:59)
  def make_env(forward_mode):
    if args.task == "tracking":
      from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
      return TrackingRlEnv(TrackingEnvCfg(num_envs=nw, seed=42 + rank, forward_mode=forward_mode), dev)
    return ManagerBasedRlEnv(VelocityEnvCfg(robot=robot, terrain="rough" if args.task == "rough" else "flat", num_envs=nw,
                                            seed=42 + rank, forward_mode=forward_mode), dev)

  env = make_env(args.forward_mode)
  m, cfg = env.sim.mj_model, env.vcfg
  sim = env.sim
  nj = cfg.nj
  gen = torch.Generator(device=dev).manual_seed(1234 + rank)
  env.reset()
  # rsl_rl's init_at_random_ep_len (src/mjlab/scripts/train.py:167-169): staggered time-outs from the first step on
  env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (nw,), device=dev, generator=gen)

  nout = cfg.n_obs_policy + cfg.n_obs_critic + 3
  host_actions = torch.empty((nw, nj), dtype=torch.float32).pin_memory()
  host_result = torch.empty((nw, nout), dtype=torch.float32).pin_memory()
  dev_result = torch.empty((nw, nout), dtype=torch.float32, device=dev)
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

  def rand_actions(k):
    return torch.rand((k, nw, nj), device=dev, generator=gen) * 2 - 1

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  warm = max(args.warmup, 3)
  wa = rand_actions(warm)
  for k in range(warm):
    env.step(wa[k])
  barrier()

  # ---- timed region: device-resident inputs ----------------------------------------------------
  clocks = Clocks(local)
  clocks.start()
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  acts = rand_actions(args.steps)
  resets = torch.zeros((), dtype=torch.float32, device=dev)
  barrier()
  t_wall0 = time.perf_counter()
  for k in range(args.steps):
    flush.zero_()
    ev[k][0].record()
    env.step(acts[k])
    ev[k][1].record()
    resets += env._buf["any_reset"][0]
  barrier()
  t_wall = time.perf_counter() - t_wall0
  elapsed_ms = sum(a.elapsed_time(b) for a, b in ev)
  mean_resets = resets.item() / args.steps

  # ---- e2e: host actions in, host observations / reward / dones out, copies inside the timed region --------
  ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  host_all = torch.rand((args.steps, nw, nj)) * 2 - 1
  barrier()
  for k in range(args.steps):
    host_actions.copy_(host_all[k])
    flush.zero_()
    ev2[k][0].record()
    obs, rew, term, trunc, _ = env.step(host_actions)      # the H2D copy of the action batch is inside step()
    dev_result[:, : cfg.n_obs_policy] = obs["policy"]
    dev_result[:, cfg.n_obs_policy: cfg.n_obs_policy + cfg.n_obs_critic] = obs["critic"]
    dev_result[:, -3] = rew
    dev_result[:, -2] = term
    dev_result[:, -1] = trunc
    host_result.copy_(dev_result, non_blocking=True)
    ev2[k][1].record()
    ev2[k][1].synchronize()  # the policy consumes the observations before it can produce the next action
  barrier()
  e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
  clk = clocks.stop()

  # ---- roofline pass: the same step enqueued eagerly with CUDA events around every kernel -------------
  roof = None
  if not args.no_roofline_pass:
    from mjlab_b200 import envabi as A
    lib = env._lib
    pm, pd, pe = C.byref(sim._cm), C.byref(sim._cd), C.byref(env._ce)
    times: dict[str, list] = {}

    def timed(name, fn):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      E.check_call(lib, fn(torch.cuda.current_stream().cuda_stream))
      b.record()
      times.setdefault(name, []).append((a, b))

    mask = env._buf["reset_buf"].data_ptr() if args.forward_mode == "masked" else None
    gate = env._buf["any_reset"].data_ptr()
    nroof = min(args.steps, 25)
    barrier()
    for k in range(nroof):
      flush.zero_()
      env._buf["action_in"].copy_(acts[k])
      trk = args.task == "tracking"
      timed("k_env_action", lambda st: (lib.b2_trk_action if trk else lib.b2_env_action)(pm, pd, pe, st))
      for _ in range(DECIMATION):
        timed("k_smooth", lambda st: lib.b2_stage(pm, pd, 0, 1, st))
        timed("k_solve", lambda st: lib.b2_stage(pm, pd, 1, 1, st))
        timed("k_post", lambda st: lib.b2_stage(pm, pd, 2, 1, st))
        if not trk:
          timed("k_env_airtime", lambda st: lib.b2_env_airtime(pm, pd, pe, st))
      timed("k_env_post", lambda st: (lib.b2_trk_post if trk else lib.b2_env_post)(pm, pd, pe, st))
      timed("forward_after_reset", lambda st: lib.b2_forward_masked(pm, pd, mask, gate, st))
      timed("k_env_obs", lambda st: (lib.b2_trk_obs if trk else lib.b2_env_obs)(pm, pd, pe, 1, st))
      timed("k_env_log", lambda st: (lib.b2_trk_log if trk else lib.b2_env_log)(pd, pe, st))
    barrier()
    avg = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in times.items()}
    per_step = {k: avg[k] * len(v) / nroof for k, v in times.items()}
    b_smooth, b_solve = algorithmic_bytes(m)
    b_env = DECIMATION * (b_smooth + b_solve) + 4 * (cfg.n_obs_policy + cfg.n_obs_critic + nj + 3)   # B_env
    peaks = {}
    try:
      peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
      pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = "k_solve" if avg["k_solve"] >= avg["k_smooth"] else "k_smooth"
    dom_bytes = b_solve if dom == "k_solve" else b_smooth
    achieved = dom_bytes * nw / (avg[dom] * 1e-3) / 1e9
    step_ms = elapsed_ms / args.steps
    roof = {
      "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
      "traffic": NCU_TRAFFIC_BYTES.get(dom) if (args.model == "g1_flat" and nw == NUM_ENVS) else None,
      "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/r01z_k_solve_final.txt",
      "peak_source": "measured burst (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)",
      "algorithmic_bytes_per_world": {"k_smooth": b_smooth, "k_solve": b_solve, "env_step": b_env},
      "avg_launch_ms": avg, "ms_per_env_step_by_kernel": per_step,
      "whole_step": {"bytes_per_env_step": b_env, "achieved": b_env * nw / (step_ms * 1e-3) / 1e9,
                     "frac": b_env * nw / (step_ms * 1e-3) / 1e9 / peak},
    }

  # ---- secondary number: forward pass restricted to the reset envs (not the reference's behaviour; reported, not the headline)
  alt = None
  if args.forward_mode == "reference" and not args.no_roofline_pass:
    env2 = make_env("masked")
    env2.reset()
    env2.episode_length_buf = env.episode_length_buf.clone()
    for k in range(warm):
      env2.step(wa[k])
    ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
      flush.zero_()
      ev3[k][0].record()
      env2.step(acts[k])
      ev3[k][1].record()
    barrier()
    alt_ms = sum(a.elapsed_time(b) for a, b in ev3)
    alt = {"forward_mode": "masked", "ms_per_step": alt_ms / args.steps, "value_this_rank": args.steps * nw / (alt_ms * 1e-3)}
    del env2

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
      "warmup": warm, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(args.model, nw, args.forward_mode, args.task), "num_envs_per_gpu": nw, "decimation": DECIMATION,
                 "timestep": 0.005, "nconmax": 35, "njmax": 300, "parallelism": f"replicas x{world} (no physics collective)",
                 "l2": "256 MiB flush between timed steps", "actions": "U(-1,1) redrawn every env step (random policy)",
                 "forward_mode": args.forward_mode, "mean_resets_per_step": mean_resets,
                 "mean_solver_niter": niter, "mean_nefc": nefc, "mean_ncon": ncon},
      "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": nw * nj * 4, "d2h_bytes_per_step": nw * nout * 4},
      "gpu_launches": args.steps * env.launches_per_step,
      "clocks": clk,
      "wall_s": t_wall,
    }
    if roof is not None:
      line["roofline"] = roof
    if alt is not None:
      line["config"]["alt_masked_forward"] = alt
    if world == 1 and not args.no_cpu_baseline and args.task in ("velocity", "rough"):
      cores = os.cpu_count() or 1
      v, secs, desc = cpu_compiled_env_steps_per_s(args.model, cores, target_s=12.0)
      line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": desc}
      v1, secs1 = cpu_env_steps_per_s(args.model, 1, 4, 2)
      line["cpu_baseline"]["numpy_oracle_1core"] = {"value": v1, "sample": f"4 worlds x 2 env steps, oracle/mjstep.py + oracle/envstep.py (the independent checker), {secs1:.1f} s"}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  a = parse()
  if a.impl == "reference":
    run_reference(a)
  else:
    run_b200(a)
