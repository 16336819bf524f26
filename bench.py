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
# DRAM bytes per launch from the committed ncu captures of this workload (profiles/r02_final_k_solve.txt,
# r02b_k_smooth.txt, r02_final_k_post.txt)
NCU_TRAFFIC_BYTES = {"k_solve": 38_247_168 + 627_200, "k_smooth": 14_992_128 + 38_866_688, "k_post": 52_210_944 + 724_992}


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=100)
  ap.add_argument("--warmup", type=int, default=20)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--model", default="g1_flat")
  ap.add_argument("--task", default="velocity", choices=["velocity", "tracking", "rough", "train"],
                  help="velocity (the headline: full env step), rough (BASELINE config 4), tracking (config 5: synthetic clip, domain "
                       "randomisation on), train (config 3: full PPO iteration = 24 env steps + 5 x 4 mini-batch updates with the gradient all-reduce)")
  ap.add_argument("--mode", default="env", choices=["env", "physics"],
                  help="env: ManagerBasedRlEnv.step (measure_env_sps); physics: the decimation x sim.step() loop alone "
                       "(scripts/benchmarks/measure_throughput.py:82-101, BASELINE config 2 with --model go1_flat)")
  ap.add_argument("--prewarm-s", type=float, default=1.5, help="untimed env steps for this many seconds before the warm-up steps "
                       "(clocks and caches reach steady state; declared in config.prewarm_s)")
  ap.add_argument("--matmul-precision", default="highest", choices=["highest", "high"], help="--task train: fp32 (default) or TF32 GEMMs")
  ap.add_argument("--num-envs", type=int, default=None, help="envs per GPU (default 4096; 8192 for --task rough, BASELINE config 4)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-roofline-pass", action="store_true")
  ap.add_argument("--forward-mode", default="reference", choices=["reference", "masked"],
                  help="forward pass after resets: all envs when any env reset (the reference's behaviour) or only the reset envs")
  args = ap.parse_args()
  if args.task == "rough":
    args.model = args.model.split("_")[0] + "_rough"
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


class CpuEnv:
  """The SAME env step on the host cores: compiled all-cores CPU build of the physics (oracle/cport, float64) + the numpy env
  stage (oracle/envstep.py), `nworld` worlds, random actions, staggered episode lengths, resets and the forward pass after
  resets included - the workload of the own arm, measured on the CPU."""

  def __init__(self, name: str, nworld: int, cores: int, forward: bool = True):
    from types import SimpleNamespace as NS

    import mjlab_b200
    from mjlab_b200 import envcfg
    from oracle import envstep as OE
    from oracle.cport import CpuSim
    self.OE, self.NS = OE, NS
    m = mjlab_b200.load_model(name)
    m.opt_timestep, m.opt_iterations, m.opt_ls_iterations = 0.005, 10, 20
    self.m, self.cfg = m, envcfg.velocity_cfg(m, name.split("_")[0])
    self.nworld, self.cores, self.forward = nworld, cores, forward
    self.sim = CpuSim(m, nworld, 35, 300)
    self.sim.qpos[:] = m.key_qpos[0]
    self.rng = np.random.default_rng(7)
    self.s = OE.make_state(self.cfg, nworld)
    self.s.episode_length[:] = self.rng.integers(0, self.cfg.max_episode_length, nworld)   # rsl_rl's init_at_random_ep_len
    self.fields = ("qpos", "qvel", "ctrl", "xpos", "xquat", "subtree_com", "cvel", "site_xpos", "sensordata", "qacc_warmstart", "qacc")
    self.nreset = 0

  def view(self):
    d = self.NS(**{f: getattr(self.sim, f) for f in self.fields})
    d.time = self.sim.time.reshape(-1)
    return d

  def step(self):
    OE, cfg, s, sim = self.OE, self.cfg, self.s, self.sim
    d = self.view()
    OE.apply_action(cfg, s, d, self.rng.uniform(-1, 1, (self.nworld, cfg.nj)))
    for _ in range(DECIMATION):
      sim.step(1, self.cores)
      OE.update_air_time(cfg, s, sim.sensordata.astype(np.float32), d.time)
    OE.post(cfg, self.m, s, d)
    self.nreset += int(s.reset_buf.sum())
    if self.forward and s.reset_buf.any():
      sim.forward(self.cores)       # sim.forward() over all envs when any env was reset (manager_based_rl_env.py:295-298)
    OE.obs_stage(cfg, self.m, s, d)


def cpu_compiled_env_steps_per_s(name: str, cores: int, nworld: int, target_s: float = 12.0, steps: int | None = None, warmup: int = 1):
  """env-steps/s of the CPU arm on `nworld` worlds.  `steps` given: exactly that many timed env steps after `warmup`; else a
  bounded sample of about `target_s` seconds."""
  env = CpuEnv(name, nworld, cores)
  env.step()  # page in the buffers
  t0 = time.perf_counter()
  env.step()
  one = time.perf_counter() - t0
  for _ in range(max(0, warmup - 2)):
    env.step()
  nsteps = int(steps) if steps is not None else int(max(3, min(200, target_s / max(one, 1e-3))))
  env.nreset = 0
  t0 = time.perf_counter()
  for _ in range(nsteps):
    env.step()
  secs = time.perf_counter() - t0
  desc = (f"{nworld} worlds x {nsteps} env steps (each: action -> ctrl, {DECIMATION} physics steps + air-time update, terminations, rewards, "
          f"resets, forward pass after resets, commands, pushes, observations; {env.nreset / nsteps:.1f} resets per step), compiled CPU build "
          f"of the same algorithm (oracle/cport, g++ -O3, float64) + numpy env stage, {cores} threads, {secs:.1f} s")
  return nworld * nsteps / secs, secs, desc, nsteps


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  k, w = args.steps, max(args.warmup, 1)
  value, secs, desc, nsteps = cpu_compiled_env_steps_per_s(args.model, cores, args.num_envs, steps=k, warmup=w)
  line = {
    "impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
    "steps": k, "warmup": w, "ms_per_step": 1e3 * secs / nsteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f64", "data": "synthetic", "same_config": True,
    "config": {"workload": workload_name(args.model, args.num_envs, args.forward_mode), "num_envs_per_gpu": args.num_envs,
               "decimation": DECIMATION, "timestep": 0.005, "nconmax": 35, "njmax": 300,
               "actions": "U(-1,1) redrawn every env step (random policy)", "forward_mode": "reference"},
    "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": desc},
    "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "note": "mujoco / mujoco_warp cannot be installed on this image or on the GPU box (profiles/r02_reference_install_attempt.log): the CPU "
            "arm is a compiled all-cores CPU build of the same algorithm on the same workload (one step = one env step of all "
            f"{args.num_envs} worlds), not MuJoCo C",
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

  physics = args.mode == "physics"
  if physics:
    # scripts/benchmarks/measure_throughput.py:82-101 (`measure_physics_sps`): the decimation x sim.step() loop alone, counted in
    # env steps.  The controls are redrawn every env step through the action path (process_action + ctrl write, SURVEY 8(d)
    # config 2), so the robots keep moving; no managers, no resets.
    lib_, pm_, pd_, pe_ = env._lib, C.byref(sim._cm), C.byref(sim._cd), C.byref(env._ce)
    gph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
      with torch.cuda.graph(gph, stream=side):
        E.check_call(lib_, lib_.b2_env_action(pm_, pd_, pe_, side.cuda_stream))
        E.check_call(lib_, lib_.b2_substeps(pm_, pd_, DECIMATION, side.cuda_stream))
    torch.cuda.current_stream().wait_stream(side)

    def do_step(a):
      env._buf["action_in"].copy_(a, non_blocking=True)
      gph.replay()
  else:
    def do_step(a):
      return env.step(a)

  # declared pre-warm: untimed steps for `prewarm_s` seconds so that clocks, instruction caches and the L2-resident model
  # tables are in steady state before the W warm-up steps (r01: with 5 warm-up steps alone the first timed steps ran 10 % slow)
  t_pre = time.perf_counter()
  npre = 0
  pw = rand_actions(16)
  while time.perf_counter() - t_pre < args.prewarm_s:
    for k in range(16):
      do_step(pw[k])
    torch.cuda.synchronize()
    npre += 16
  warm = max(args.warmup, 3)
  wa = rand_actions(warm)
  for k in range(warm):
    do_step(wa[k])
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
    do_step(acts[k])
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
    if physics:
      do_step(host_actions)                                  # H2D copy of the action batch, then the physics loop
      dev_result.zero_()
      dev_result[:, : m.nq] = sim.data.qpos                  # the state a caller of sim.step() reads back
      dev_result[:, m.nq: m.nq + m.nv] = sim.data.qvel
    else:
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
      for sub in range(DECIMATION):
        if sub == 0 and not physics:   # the first sub-step reuses the forward pass that ended the previous env step
          timed("k_smooth_after_forward", lambda st: lib.b2_stage(pm, pd, 6, 1, st))
        else:
          timed("k_smooth", lambda st: lib.b2_stage(pm, pd, 4, 1, st))
        timed("k_order", lambda st: lib.b2_stage(pm, pd, 5, 1, st))
        timed("k_solve", lambda st: lib.b2_stage(pm, pd, 1, 1, st))
        timed("k_post", lambda st: lib.b2_stage(pm, pd, 2, 1, st))
        if not trk and not physics:
          timed("k_env_airtime", lambda st: lib.b2_env_airtime(pm, pd, pe, st))
      if physics:
        continue
      timed("k_env_post", lambda st: (lib.b2_trk_post if trk else lib.b2_env_post)(pm, pd, pe, st))
      timed("forward_after_reset", lambda st: lib.b2_forward_masked(pm, pd, mask, gate, st))
      timed("k_env_obs", lambda st: (lib.b2_trk_obs if trk else lib.b2_env_obs)(pm, pd, pe, 1, st))
      timed("k_env_log", lambda st: (lib.b2_trk_log if trk else lib.b2_env_log)(pd, pe, st))
    barrier()
    avg = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in times.items()}
    per_step = {k: avg[k] * len(v) / nroof for k, v in times.items()}
    # the graph replay's own time against the sum of its kernels timed one by one (eager launches): what is left is launch
    # gaps inside the graph and the action copy; negative = the graph packs the kernels tighter than eager launches do
    per_step["graph_step_minus_kernel_sum"] = elapsed_ms / args.steps - sum(per_step.values())
    b_smooth, b_solve = algorithmic_bytes(m)
    b_env = DECIMATION * (b_smooth + b_solve) + (0 if physics else 4 * (cfg.n_obs_policy + cfg.n_obs_critic + nj + 3))   # B_env
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
      "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/r02_final_k_{solve,post}.txt, r02b_k_smooth.txt",
      "peak_source": "measured burst (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)",
      "algorithmic_bytes_per_world": {"k_smooth": b_smooth, "k_solve": b_solve, "env_step": b_env},
      "avg_launch_ms": avg, "ms_per_env_step_by_kernel": per_step,
      "whole_step": {"bytes_per_env_step": b_env, "achieved": b_env * nw / (step_ms * 1e-3) / 1e9,
                     "frac": b_env * nw / (step_ms * 1e-3) / 1e9 / peak},
    }

  # ---- secondary number: forward pass restricted to the reset envs (not the reference's behaviour; reported, not the headline)
  alt = None
  if args.forward_mode == "reference" and not args.no_roofline_pass and not physics:
    env2 = make_env("masked")
    env2.reset()
    env2.episode_length_buf = env.episode_length_buf.clone()
    for k in range(npre):           # the same pre-warm as the main arm: the state mix (fallen / standing worlds) sets the solver's load
      env2.step(pw[k % 16])
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
      "config": {"workload": (f"{args.model}: process_action + ctrl write, then {DECIMATION} x sim.step() (measure_physics_sps: physics only, "
                              f"no managers, no resets), {nw} envs/GPU - one CUDA graph" if physics else
                              workload_name(args.model, nw, args.forward_mode, args.task)), "mode": args.mode, "num_envs_per_gpu": nw, "decimation": DECIMATION,
                 "prewarm_s": args.prewarm_s, "prewarm_steps": npre,
                 "timestep": 0.005, "nconmax": 35, "njmax": 300, "parallelism": f"replicas x{world} (no physics collective)",
                 "l2": "256 MiB flush between timed steps", "actions": "U(-1,1) redrawn every env step (random policy)",
                 "forward_mode": args.forward_mode, "mean_resets_per_step": mean_resets,
                 "mean_solver_niter": niter, "mean_nefc": nefc, "mean_ncon": ncon},
      "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": nw * nj * 4, "d2h_bytes_per_step": nw * nout * 4},
      "gpu_launches": args.steps * ((1 + 4 * DECIMATION) if physics else env.launches_per_step),
      "clocks": clk,
      "wall_s": t_wall,
    }
    if roof is not None:
      line["roofline"] = roof
    if alt is not None:
      line["config"]["alt_masked_forward"] = alt
    if world == 1 and not args.no_cpu_baseline and args.task in ("velocity", "rough") and not physics:
      cores = os.cpu_count() or 1
      v, secs, desc, _ = cpu_compiled_env_steps_per_s(args.model, cores, nw, target_s=12.0)
      line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": desc}
      v1, secs1 = cpu_env_steps_per_s(args.model, 1, 4, 2)
      line["cpu_baseline"]["numpy_oracle_1core"] = {"value": v1, "sample": f"4 worlds x 2 env steps, oracle/mjstep.py + oracle/envstep.py (the independent checker), {secs1:.1f} s"}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# --task train: BASELINE config 3, the full PPO iteration (the one path of the reference with a collective)
# ---------------------------------------------------------------------------------------------
def run_train(args):
  """One step = one PPO iteration of the reference's training loop (rsl_rl OnPolicyRunner.learn with config/g1/rl_cfg.py:10-39):
  24 env steps x num_envs envs with policy inference (roll-out), GAE, then 5 epochs x 4 mini-batches of forward / backward /
  gradient all-reduce over the ranks / clip / Adam with the KL-adaptive learning rate.  value = 24 x num_envs x n_gpus / iteration
  time; fp32 GEMMs unless --matmul-precision high."""
  import torch
  import torch.distributed as dist

  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from mjlab_b200.ppo import OnPolicyRunner, PpoCfg
  from mjlab_b200.vecenv import RslRlVecEnvWrapper

  world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = f"cuda:{local}"
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
  nw = args.num_envs
  robot = args.model.split("_")[0]
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot=robot, num_envs=nw, seed=42 + rank, forward_mode=args.forward_mode), dev)
  pcfg = PpoCfg(matmul_precision=args.matmul_precision)
  runner = OnPolicyRunner(RslRlVecEnvWrapper(env), pcfg)
  T = pcfg.num_steps_per_env

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # graph captures (roll-out graph on iteration 1, update graphs after two eager updates) + steady state, then W warm-ups
  runner.learn(4)
  t_pre = time.perf_counter()
  while time.perf_counter() - t_pre < args.prewarm_s:
    runner.learn(1)
    torch.cuda.synchronize()
  warm = max(args.warmup, 3)
  runner.learn(warm)
  barrier()
  clocks = Clocks(local)
  clocks.start()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
  barrier()
  t_wall0 = time.perf_counter()
  for k in range(args.steps):
    ev[3 * k].record()
    ro = runner.collect()
    ev[3 * k + 1].record()
    runner.update(ro)
    ev[3 * k + 2].record()
    runner.iteration += 1
  barrier()
  t_wall = time.perf_counter() - t_wall0
  ro_ms = sum(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(args.steps))
  up_ms = sum(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(args.steps))
  # e2e: the public API (learn), whose iteration reads its metrics back to the host (reward, episode statistics, KL, lr)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  hist = runner.learn(args.steps)
  e1.record()
  barrier()
  e2e_ms = e0.elapsed_time(e1)
  clk = clocks.stop()
  t = torch.tensor([ro_ms + up_ms, e2e_ms, ro_ms, up_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  elapsed_ms, e2e_ms, ro_ms, up_ms = t.tolist()
  total = args.steps * T * nw * world
  if rank == 0:
    nparam = sum(p.numel() for p in runner.policy.parameters())
    nmb = pcfg.num_learning_epochs * pcfg.num_mini_batches
    line = {
      "metric": "env-steps/s", "value": total / (elapsed_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
      "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32" if args.matmul_precision == "highest" else "tf32 GEMMs (physics f32)", "data": "synthetic",
      "config": {"workload": f"{args.model} velocity task, full PPO train iteration (BASELINE config 3): {T} env steps x {nw} envs/GPU with policy "
                             f"inference, GAE, {pcfg.num_learning_epochs} epochs x {pcfg.num_mini_batches} mini-batches (MLP 512-256-128 ELU, obs normalisation, "
                             "adaptive learning rate), gradient all-reduce over the ranks", "task": "train", "num_envs_per_gpu": nw,
                 "decimation": DECIMATION, "timestep": 0.005, "parallelism": f"dp{world}: one world batch per rank, {nmb} all-reduces of {4 * (nparam + 1)} B per iteration",
                 "matmul_precision": args.matmul_precision, "prewarm_s": args.prewarm_s, "l2": "inputs larger than L2: the roll-out storage of one iteration is "
                 f"{4 * T * nw * (env.vcfg.n_obs_policy + env.vcfg.n_obs_critic + 3 * env.vcfg.nj + 5) / 2**20:.0f} MiB",
                 "iteration_split_ms": {"rollout": ro_ms / args.steps, "update": up_ms / args.steps},
                 "update_path": ("per mini-batch: graph (forward + backward) -> NCCL all-reduce -> graph (clip + Adam)" if world > 1 else "one CUDA graph")},
      "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * 11},
      "gpu_launches": args.steps * (T * env.launches_per_step),
      "clocks": clk, "wall_s": t_wall,
      "last_iteration": {k: hist[-1][k] for k in ("mean_step_reward", "kl", "lr", "mean_episode_length")},
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  a = parse()
  if a.impl == "reference":
    run_reference(a)
  elif a.task == "train":
    run_train(a)
  else:
    run_b200(a)
