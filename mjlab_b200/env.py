"""`ManagerBasedRlEnv` for the velocity-tracking tasks on the B200 engine.

Host-side mirror of `mjlab.envs.ManagerBasedRlEnv` (src/mjlab/envs/manager_based_rl_env.py:98-416) specialised to
`Mjlab-Velocity-Flat-Unitree-{G1,Go1}` (tasks/velocity/velocity_env_cfg.py, config/{g1,go1}/env_cfgs.py).  Same public
surface - `step(action) -> (obs, reward, terminated, time_outs, extras)`, `reset()`, `num_envs`, `step_dt`,
`episode_length_buf`, `sim`, `scene["robot"].data`, `action_manager.action`, `command_manager.get_command("twist")`,
`reward_manager`, `termination_manager`, `observation_manager.compute()`, `extras["log"]` - but one env step is ONE CUDA
graph of hand-written kernels behind the C ABI (include/b200env.h):

    b2_env_action -> decimation x (k_smooth, k_solve, b2_env_airtime) -> b2_env_post (terminations, rewards, resets)
    -> conditional forward -> b2_env_obs (commands, push events, observations) -> b2_env_log

No host synchronisation happens inside `step`: the reference's `nonzero()` round trips (reset ids, command resample ids,
interval-event ids) are device-side predicates, and the forward pass after resets is gated by a device counter.
"""

from __future__ import annotations

import ctypes as C
import math
import os
from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import Any, Literal

import numpy as np
import torch

from . import engine as E
from . import envabi as A
from . import envcfg
from .entity import Entity
from .sim import MujocoCfg, Simulation, SimulationCfg

_TORCH = {"f4": torch.float32, "i4": torch.int32, "i8": torch.int64, "u1": torch.uint8}
_TASK_ROBOT = {"Mjlab-Velocity-Flat-Unitree-G1": ("g1", "flat"), "Mjlab-Velocity-Flat-Unitree-Go1": ("go1", "flat"),
               "Mjlab-Velocity-Rough-Unitree-G1": ("g1", "rough"), "Mjlab-Velocity-Rough-Unitree-Go1": ("go1", "rough")}


def _default_sim_cfg() -> SimulationCfg:
  # tasks/velocity/velocity_env_cfg.py:343-351
  return SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20))


@dataclass
class VelocityEnvCfg:
  """What `load_env_cfg("Mjlab-Velocity-Flat-Unitree-<robot>", play=...)` resolves to (tasks/registry.py:48-55)."""

  robot: Literal["g1", "go1"] = "g1"
  # "flat": ground plane (the Flat tasks); "rough": the ROUGH_TERRAINS_CFG box grid with the terrain-level curriculum
  # (velocity_env_cfg.py:296-324, config/g1/env_cfgs.py:20-150; terrain seed fixed to 0)
  terrain: Literal["flat", "rough"] = "flat"
  num_envs: int = 4096
  play: bool = False
  seed: int = 42
  decimation: int = 4
  env_spacing: float = 2.0                      # terrains/terrain_importer.py:44
  foot_friction_range: tuple[float, float] | None = (0.3, 1.2)   # startup DR event (velocity_env_cfg.py:168-179)
  # "reference": sim.forward() over ALL envs whenever any env was reset (manager_based_rl_env.py:293-298);
  # "masked": only the reset envs are re-evaluated (non-reset envs keep the derived quantities of their last step).
  forward_mode: Literal["reference", "masked"] = "reference"
  log_metrics: bool = True
  # command curriculum `commands_vel` (velocity_env_cfg.py:293-305): applied when common_step_counter passes "step"
  velocity_stages: tuple = (
    {"step": 0, "lin_vel_x": (-1.0, 1.0), "ang_vel_z": (-0.5, 0.5)},
    {"step": 5000 * 24, "lin_vel_x": (-1.5, 2.0), "ang_vel_z": (-0.7, 0.7)},
    {"step": 10000 * 24, "lin_vel_x": (-2.0, 3.0)},
  )
  sim: SimulationCfg = field(default_factory=_default_sim_cfg)
  # a reference `ManagerBasedRlEnvCfg` (the live dataclass, or its `envcfg.cfg_to_dict` view) whose reward / termination /
  # observation / event / command / curriculum terms replace the built-in task tables (see `from_manager_cfg`)
  manager_cfg: Any = None

  @classmethod
  def from_manager_cfg(cls, cfg, robot: str, terrain: str = "flat", **kw) -> "VelocityEnvCfg":
    """Consume the reference's `ManagerBasedRlEnvCfg` (envs/manager_based_rl_env.py:59-129, e.g. `load_env_cfg(task)` edited by
    the user: re-weighted / removed reward terms, other noise or command ranges, no push event, ...).  The SCENE is not rebuilt
    from the cfg's `spec_fn`s (that needs MuJoCo's compiler): `robot` / `terrain` name one of the compiled scenes; every term is
    taken from the cfg and must be one the kernels implement (`envcfg.UnsupportedTermError` otherwise)."""
    d = cfg if isinstance(cfg, dict) else envcfg.cfg_to_dict(cfg)
    mj, sc = d["sim"]["mujoco"], d["sim"]
    sim = SimulationCfg(nconmax=sc["nconmax"], njmax=sc["njmax"], ls_parallel=sc.get("ls_parallel", True),
                        contact_sensor_maxmatch=sc.get("contact_sensor_maxmatch", 64),
                        mujoco=MujocoCfg(timestep=mj["timestep"], integrator=mj["integrator"], impratio=mj["impratio"], cone=mj["cone"],
                                         solver=mj["solver"], iterations=mj["iterations"], tolerance=mj["tolerance"],
                                         ls_iterations=mj["ls_iterations"], ls_tolerance=mj["ls_tolerance"], gravity=tuple(mj["gravity"])))
    play = float(d["episode_length_s"]) > 1e6      # which compiled rough scene (5 x 5 play grid or 10 x 20 curriculum grid)
    kw.setdefault("num_envs", int(d["scene"]["num_envs"]))
    kw.setdefault("seed", d.get("seed") if d.get("seed") is not None else 42)
    return cls(robot=robot, terrain=terrain, play=play, decimation=int(d["decimation"]), sim=sim, manager_cfg=d, **kw)

  @classmethod
  def from_task(cls, task: str, play: bool = False, **kw) -> "VelocityEnvCfg":
    if task not in _TASK_ROBOT:
      raise ValueError(f"Unknown task '{task}'. Available: {sorted(_TASK_ROBOT)}")
    robot, terrain = _TASK_ROBOT[task]
    return cls(robot=robot, terrain=terrain, play=play, **kw)


def env_origins_grid(num_envs: int, spacing: float) -> np.ndarray:
  """Grid of env origins on the plane (terrains/terrain_importer.py:246-261)."""
  rows = math.ceil(num_envs / int(math.sqrt(num_envs)))
  cols = math.ceil(num_envs / rows)
  ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
  o = np.zeros((num_envs, 3), np.float32)
  o[:, 0] = -(ii.reshape(-1)[:num_envs] - (rows - 1) / 2) * spacing
  o[:, 1] = (jj.reshape(-1)[:num_envs] - (cols - 1) / 2) * spacing
  return o


class _Scene(dict):
  """`env.scene`: entities and sensors by name (scene/scene.py:97-130).  Sensor views are built on first access from the compiled
  model's sensors (scene/scene.py:196-198 auto-discovers the XML sensors the same way): contact sensors by their cfg name,
  builtin sensors by their full name; `feet_ground_contact` shares the air-time state the fused env kernels advance."""

  def __init__(self, env: "ManagerBasedRlEnv", entities: dict):
    super().__init__(entities)
    self._env = env

  @property
  def env_origins(self) -> torch.Tensor:
    return self._env._buf["env_origins"]

  @property
  def num_envs(self) -> int:
    return self._env.num_envs

  def __missing__(self, name: str):
    from .sensors import SENS_CONTACT, BuiltinSensor, ContactSensor
    sim = self._env.sim
    m = sim.mj_model
    names = list(m.sensor_names)
    if name in names:
      s = BuiltinSensor(name, sim)
    elif any(n.startswith(name + "_") and int(m.sensor_type[i]) == SENS_CONTACT for i, n in enumerate(names)):
      fused = name == "feet_ground_contact"
      s = ContactSensor(name, sim, air_state=self._env._buf["air_state"] if fused else None,
                        air_last_time=self._env._buf["air_last_time"] if fused else None)
    else:
      raise KeyError(f"'{name}' is neither an entity nor a sensor of the scene; sensors: {names}")
    self[name] = s
    return s


class _LogView(Mapping):
  """`extras["log"]`: the reference's per-step log entries computed lazily from a snapshot of the device log sums."""

  def __init__(self, env: "ManagerBasedRlEnv", snap: torch.Tensor):
    self._env, self._snap, self._vals = env, snap, None

  def _resolve(self) -> dict[str, float]:
    if self._vals is None:
      v = self._snap.tolist()
      K, cfg = A.K, self._env.vcfg
      n = max(v[K["B2E_LOG_NRESET"]], 1.0)
      out: dict[str, float] = {}
      for name, k in self._env.reward_manager.slots.items():
        out["Episode_Reward/" + name] = v[K["B2E_LOG_EPISODE_SUM"] + k] / n / cfg.max_episode_length_s
      for name, k in self._env.termination_manager.slots.items():
        out["Episode_Termination/" + name] = v[K["B2E_LOG_TERM_COUNT"] + k]
      out["Metrics/twist/error_vel_xy"] = v[K["B2E_LOG_CMD_ERR_XY"]] / n
      out["Metrics/twist/error_vel_yaw"] = v[K["B2E_LOG_CMD_ERR_YAW"]] / n
      B = float(self._env.num_envs)
      out["Metrics/angular_momentum_mean"] = v[K["B2E_LOG_ANGMOM"]] / B
      out["Metrics/air_time_mean"] = v[K["B2E_LOG_AIR_TIME_SUM"]] / max(v[K["B2E_LOG_AIR_COUNT"]], 1.0)
      land = max(v[K["B2E_LOG_LANDINGS"]], 1.0)
      out["Metrics/peak_height_mean"] = v[K["B2E_LOG_PEAK_SUM"]] / land
      out["Metrics/slip_velocity_mean"] = v[K["B2E_LOG_SLIP_SUM"]] / max(v[K["B2E_LOG_NCONTACT"]], 1.0)
      out["Metrics/landing_force_mean"] = v[K["B2E_LOG_LANDING_FORCE_SUM"]] / land
      out["Episode/num_resets"] = v[K["B2E_LOG_NRESET"]]
      if cfg.terrain_mode:
        out["Curriculum/terrain_levels"] = v[K["B2E_LOG_TERRAIN_LEVEL"]] / B
      self._vals = out
    return self._vals

  def __getitem__(self, k):
    return self._resolve()[k]

  def __iter__(self):
    return iter(self._resolve())

  def __len__(self):
    return len(self._resolve())


class _ActionManager:
  def __init__(self, env):
    self._env = env
    self.active_terms = ["joint_pos"]

  total_action_dim = property(lambda self: self._env.vcfg.nj)
  action_term_dim = property(lambda self: [self._env.vcfg.nj])
  action = property(lambda self: self._env._buf["action"])
  prev_action = property(lambda self: self._env._buf["prev_action"])
  prev_prev_action = property(lambda self: self._env._buf["prev_prev_action"])
  raw_action = property(lambda self: self._env._buf["raw_action"])
  processed_action = property(lambda self: self._env._buf["processed_action"])


class _CommandManager:
  def __init__(self, env):
    self._env = env
    self.active_terms = ["twist"]

  def get_command(self, name: str) -> torch.Tensor:
    if name != "twist":
      raise KeyError(name)
    return self._env._buf["command"]

  time_left = property(lambda self: self._env._buf["cmd_time_left"])
  heading_target = property(lambda self: self._env._buf["heading_target"])
  heading_error = property(lambda self: self._env._buf["heading_error"])
  is_heading_env = property(lambda self: self._env._buf["is_heading"].view(torch.bool))
  is_standing_env = property(lambda self: self._env._buf["is_standing"].view(torch.bool))
  command_counter = property(lambda self: self._env._buf["cmd_counter"])
  metrics = property(lambda self: {"error_vel_xy": self._env._buf["cmd_metrics"][:, 0], "error_vel_yaw": self._env._buf["cmd_metrics"][:, 1]})


class _RewardManager:
  def __init__(self, env):
    self._env = env
    # term name -> kernel slot: the shipped task's 13 / 14 terms always, the generic library (envs/mdp/rewards.py) when weighted
    core = envcfg.REWARD_NAMES.index("self_collisions")
    w = env.vcfg.reward_weights
    staged = {k for k, _ in env.vcfg.extra.get("reward_weight_stages", ())}    # terms a reward_weight curriculum will switch on later
    self.slots = {n: k for k, n in enumerate(envcfg.REWARD_NAMES)
                  if k < core or (k == core and env.vcfg.selfcol_adr >= 0) or (k > core and (w[k] != 0 or k in staged))}
    self.active_terms = list(self.slots)

  _step_reward = property(lambda self: self._env._buf["step_reward"])
  _reward_buf = property(lambda self: self._env._buf["reward"])
  _episode_sums = property(lambda self: {n: self._env._buf["episode_sums"][:, k] for n, k in self.slots.items()})

  def get_active_iterable_terms(self, env_idx: int):
    row = self._env._buf["step_reward"][env_idx].tolist()
    return [(n, [row[k]]) for n, k in self.slots.items()]


class _TerminationManager:
  def __init__(self, env):
    self._env = env
    on = {"time_out": True, "fell_over": True, "illegal_contact": env.vcfg.illegal_count > 0,
          "root_height_below_minimum": env.vcfg.min_root_height is not None, "nan_detection": env.vcfg.nan_termination}
    self.slots = {n: k for k, n in enumerate(envcfg.TERMINATION_NAMES) if on[n]}
    self.active_terms = list(self.slots)

  terminated = property(lambda self: self._env._buf["terminated"].view(torch.bool))
  time_outs = property(lambda self: self._env._buf["time_outs"].view(torch.bool))
  dones = property(lambda self: self._env._buf["reset_buf"].view(torch.bool))

  def get_term(self, name: str) -> torch.Tensor:
    return self._env._buf["term_dones"][:, self.slots[name]].view(torch.bool)


class _ObservationManager:
  def __init__(self, env):
    self._env = env
    nj, nf = env.vcfg.nj, env.vcfg.nfeet
    pol = [("base_lin_vel", 3), ("base_ang_vel", 3), ("projected_gravity", 3), ("joint_pos", nj), ("joint_vel", nj), ("actions", nj), ("command", 3)]
    cri = pol + [("foot_height", nf), ("foot_air_time", nf), ("foot_contact", nf), ("foot_contact_forces", 3 * nf)]
    self._fixed = {"policy": pol, "critic": cri}

  # group / term tables: the kernel's two groups, or - with a post-processing / gathering pipeline (a consumed cfg with its own
  # groups, history, ...) - what the pipeline emits (observation_manager.py `active_terms`, `group_obs_term_dim`, `group_obs_dim`)
  @property
  def active_terms(self) -> dict[str, list[str]]:
    p = self._env._obs_pipeline
    return {g: [t["name"] for t in ts] for g, ts in p.groups.items()} if p is not None else {g: [n for n, _ in ts] for g, ts in self._fixed.items()}

  @property
  def group_obs_term_dim(self) -> dict[str, list[tuple]]:
    p = self._env._obs_pipeline
    return {g: p.term_dims(g) for g in p.groups} if p is not None else {g: [(d,) for _, d in ts] for g, ts in self._fixed.items()}

  @property
  def group_obs_dim(self) -> dict[str, tuple]:
    out = {}
    for g, dims in self.group_obs_term_dim.items():
      out[g] = (sum(int(torch.tensor(d).prod()) for d in dims),) if all(len(d) == 1 for d in dims) else tuple(dims)
    return out

  @property
  def group_obs_concatenate(self) -> dict[str, bool]:
    return {g: True for g in self.active_terms}

  def compute(self, update_history: bool = False) -> dict[str, torch.Tensor]:
    """The buffers of the last `step` / `reset` (the reference returns its cached dict unless update_history is set, :160-161)."""
    return self._env.obs_buf


class ManagerBasedRlEnv:
  """Velocity-task RL env: `num_envs` worlds on one B200, one CUDA graph per env step."""

  is_vector_env = True
  metadata = {"render_modes": [None], "mujoco_version": "b200sim", "warp_version": "none"}

  def __init__(self, cfg: VelocityEnvCfg, device: str = "cuda:0", render_mode: str | None = None, **kwargs: Any):
    from . import load_model
    self.cfg = cfg
    self.device = device
    self.render_mode = render_mode
    self._lib = A.load_library()
    model = load_model(f"{cfg.robot}_flat" if cfg.terrain == "flat" else (f"{cfg.robot}_rough_play" if cfg.play else f"{cfg.robot}_rough"))
    self.sim = Simulation(cfg.num_envs, cfg.sim, model, device)
    dev = torch.device(device)
    self._dev = dev
    self.vcfg = envcfg.velocity_cfg(model, cfg.robot, play=cfg.play)
    if cfg.manager_cfg is not None:
      self.vcfg = envcfg.from_manager_cfg(cfg.manager_cfg, model, cfg.robot)
      ex = self.vcfg.extra
      cfg.foot_friction_range = None          # the cfg's own startup events are applied below, whatever field they randomise
      cfg.velocity_stages = tuple(ex["velocity_stages"])
      post = [t for g in ex["obs_post"].values() for t in g]
      if ex["obs_custom"] or any(t["clip"] or t["scale"] is not None or t["delay_max_lag"] > 0 or t["history_length"] > 0 for t in post):
        self._want_obs_pipeline = ex["obs_post"]
        self._obs_layout = ex["obs_layout"]
    self.vcfg.seed = int(cfg.seed)
    self.vcfg.decimation = cfg.decimation
    self.vcfg.step_dt = cfg.decimation * cfg.sim.mujoco.timestep
    n = cfg.num_envs
    torch.manual_seed(cfg.seed)

    # ---- b2_env: parameters, tables, per-env state ------------------------------------------------
    self._ce = A.B2Env()
    A.fill_params(self._ce, self.vcfg)
    tables = A.table_values(self.vcfg)
    self._buf: dict[str, torch.Tensor] = {}
    for i, (_, attr, kind, dt, shp) in enumerate(A.BUFFERS):
      if kind == "table":
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(tables[attr])), device=dev).to(_TORCH[dt]).contiguous()
      elif kind == "state":
        t = torch.zeros((n,) + tuple(shp(self.vcfg)), dtype=_TORCH[dt], device=dev)
      else:
        t = torch.zeros(tuple(shp(self.vcfg)), dtype=_TORCH[dt], device=dev)
      self._buf[attr] = t
      self._ce.buf[i] = t.data_ptr()
    if self.vcfg.terrain_mode:
      # TerrainImporter._compute_env_origins_curriculum (terrains/terrain_importer.py:224-244): column by env index, a
      # uniformly drawn start row up to max_init_terrain_level
      from . import terrains
      org = self.vcfg.terrain_origins
      rows, cols = org.shape[:2]
      types = terrains.terrain_types(n, cols)
      levels = np.random.default_rng(cfg.seed).integers(0, min(self.vcfg.max_init_terrain_level, rows - 1) + 1, n)
      self._buf["terrain_type"][:] = torch.as_tensor(types, device=dev).to(torch.int32)
      self._buf["terrain_level"][:] = torch.as_tensor(levels, device=dev).to(torch.int32)
      self._buf["env_origins"][:] = torch.as_tensor(org[levels, types], dtype=torch.float32, device=dev)
    else:
      self._buf["env_origins"][:] = torch.as_tensor(env_origins_grid(n, cfg.env_spacing), device=dev)
    if self.vcfg.push_enabled:  # EventManager._prepare_terms (managers/event_manager.py:196-210)
      lo, hi = self.vcfg.push_interval
      self._buf["push_time_left"][:] = torch.rand(n, device=dev) * (hi - lo) + lo
    E.check_call(self._lib, self._lib.b2_env_check(C.byref(self.sim._cm), C.byref(self.sim._cd), C.byref(self._ce)))

    # ---- scene / managers (facades over the same buffers) --------------------------------------------
    from .robots import g1_cfg, go1_cfg
    groups = [a.joint_names_expr for a in {"g1": g1_cfg, "go1": go1_cfg}[cfg.robot]()[0]]   # one actuator object per actuator cfg, as the reference entity
    self.robot = Entity("robot", self.sim, joint_pos_target=self._buf["joint_pos_target"], encoder_bias=self._buf["encoder_bias"], actuator_groups=groups)
    self.scene = _Scene(self, {"robot": self.robot})
    self.action_manager = _ActionManager(self)
    self.command_manager = _CommandManager(self)
    self.reward_manager = _RewardManager(self)
    self.termination_manager = _TerminationManager(self)
    self.observation_manager = _ObservationManager(self)
    self.extras: dict[str, Any] = {"log": {}}
    self.common_step_counter = 0
    self._sim_step_counter = 0
    self.obs_buf = {"policy": self._buf["obs_policy"], "critic": self._buf["obs_critic"]}
    self._raw_obs = {**self.obs_buf, "extra": self._buf["obs_extra"],   # the kernels' term bank (cfg-built groups gather from it)
                     "sensordata": self.sim.data.sensordata}            # builtin_sensor terms outside the bank: slices of the physics kernels' sensordata
    self._obs_pipeline = None
    if getattr(self, "_want_obs_pipeline", None) is not None:
      # clip / scale / delay / history asked for by some observation term: post-processing outside the step graph
      from .obs_pipeline import ObservationPipeline
      nj, nf = self.vcfg.nj, self.vcfg.nfeet
      pol = [("base_lin_vel", 3), ("base_ang_vel", 3), ("projected_gravity", 3), ("joint_pos", nj), ("joint_vel", nj), ("actions", nj), ("command", 3)]
      cri = pol + [("foot_height", nf), ("foot_air_time", nf), ("foot_contact", nf), ("foot_contact_forces", 3 * nf)]
      layout = getattr(self, "_obs_layout", None) or {"policy": pol, "critic": cri}   # a consumed cfg: its own groups, gathered from the bank
      self._obs_pipeline = ObservationPipeline(layout, self._want_obs_pipeline, n, device)
    self.reward_buf = self._buf["reward"]
    self.reset_buf = self._buf["reset_buf"].view(torch.bool)
    self.reset_terminated = self._buf["terminated"].view(torch.bool)
    self.reset_time_outs = self._buf["time_outs"].view(torch.bool)

    # ---- startup events: foot-friction domain randomisation (velocity_env_cfg.py:168-179) ------------
    if cfg.foot_friction_range is not None:
      from .events import randomize_field
      self.sim.expand_model_fields(("geom_friction",))
      names = self._foot_geom_names()
      randomize_field(self, None, "geom_friction", cfg.foot_friction_range, operation="abs", geom_names=names)

    for ev in self.vcfg.extra.get("startup_dr", []):   # a consumed ManagerBasedRlEnvCfg: its startup events (randomize_field / pd gains / effort limits / encoder bias), in order
      from .events import apply_startup_event
      apply_startup_event(self, ev)

    if cfg.manager_cfg is None and cfg.play and cfg.robot == "g1" and cfg.terrain == "flat":  # config/g1/env_cfgs.py:167-173: the FLAT play cfg only (pinned by tests/test_task_cfg.py)
      self.vcfg.cmd_lin_x, self.vcfg.cmd_ang_z = (-1.5, 2.0), (-0.7, 0.7)
      A.fill_params(self._ce, self.vcfg)
    self._next_stage = 0
    st = [s_[0] for _, stages in self.vcfg.extra.get("reward_weight_stages", ()) for s_ in stages]
    self._next_reward_stage_step = min(st) if st else None
    self._graph: torch.cuda.CUDAGraph | None = None
    self.use_cuda_graph = os.environ.get("B200SIM_NO_GRAPH", "0") != "1"  # eager launches for profilers
    self._capture()
    self.sim.add_graph_listener(self._capture)   # sim.expand_model_fields after construction: the step graph follows

  # ---- properties (manager_based_rl_env.py:171-203) --------------------------------------------------
  num_envs = property(lambda self: self.cfg.num_envs)
  physics_dt = property(lambda self: self.cfg.sim.mujoco.timestep)
  step_dt = property(lambda self: self.cfg.sim.mujoco.timestep * self.cfg.decimation)
  max_episode_length_s = property(lambda self: self.vcfg.max_episode_length_s)
  max_episode_length = property(lambda self: self.vcfg.max_episode_length)
  unwrapped = property(lambda self: self)

  @property
  def episode_length_buf(self) -> torch.Tensor:
    return self._buf["episode_length"]

  @episode_length_buf.setter
  def episode_length_buf(self, value: torch.Tensor) -> None:
    self._buf["episode_length"].copy_(value)   # in place: the kernels hold this address (rl/vecenv_wrapper.py:55-61)

  def _foot_geom_names(self) -> list[str]:
    if self.cfg.robot == "g1":
      return [f"{s}_foot{i}_collision" for s in ("left", "right") for i in range(1, 8)]
    return [f"{leg}_foot_collision" for leg in ("FR", "FL", "RR", "RL")]

  # ---- graph ------------------------------------------------------------------------------------------
  def _stream(self) -> int:
    return torch.cuda.current_stream(self._dev).cuda_stream

  def _enqueue_step(self) -> None:
    lib, m, d, e, chk = self._lib, C.byref(self.sim._cm), C.byref(self.sim._cd), C.byref(self._ce), E.check_call
    st = self._stream()
    chk(lib, lib.b2_env_action(m, d, e, st))
    chk(lib, lib.b2_env_substeps(m, d, e, self.cfg.decimation, st))   # decimation x (physics step + air-time update)
    chk(lib, lib.b2_env_post(m, d, e, st))
    mask = self._buf["reset_buf"].data_ptr() if self.cfg.forward_mode == "masked" else None
    chk(lib, lib.b2_forward_masked(m, d, mask, self._buf["any_reset"].data_ptr(), st))
    chk(lib, lib.b2_env_obs(m, d, e, 1, st))
    chk(lib, lib.b2_env_log(d, e, st))

  def _step_guarded(self) -> None:
    """The env step with `NanGuard.watch` around every PHYSICS step, as in the reference, where the env calls `sim.step()`
    (manager_based_rl_env.py:275-280 -> sim.py:267-273).  The fused graph bypasses `Simulation.step()`, and a check after the
    whole env step would come too late: a world that went NaN is usually terminated and reset inside the same step.  Debug
    path (eager launches, one host sync per physics step); only taken when `cfg.sim.nan_guard.enabled`."""
    lib, m, d, e, chk = self._lib, C.byref(self.sim._cm), C.byref(self.sim._cd), C.byref(self._ce), E.check_call
    st = self._stream()
    chk(lib, lib.b2_env_action(m, d, e, st))
    for _ in range(self.cfg.decimation):
      with self.sim.nan_guard.watch(self.sim.data):
        chk(lib, lib.b2_env_substeps(m, d, e, 1, st))
    chk(lib, lib.b2_env_post(m, d, e, st))
    mask = self._buf["reset_buf"].data_ptr() if self.cfg.forward_mode == "masked" else None
    chk(lib, lib.b2_forward_masked(m, d, mask, self._buf["any_reset"].data_ptr(), st))
    chk(lib, lib.b2_env_obs(m, d, e, 1, st))
    chk(lib, lib.b2_env_log(d, e, st))

  def _capture(self) -> None:
    """(Re)capture the env-step graph; needed again whenever a buffer address or a by-value parameter changes."""
    self._graph = None
    if not self.use_cuda_graph:
      return
    torch.cuda.synchronize(self._dev)
    side = torch.cuda.Stream(self._dev)
    g = torch.cuda.CUDAGraph()
    n0 = self._lib.b2_launch_count()
    with torch.cuda.stream(side):
      with torch.cuda.graph(g, stream=side):
        self._enqueue_step()
    self._launches_per_step = int(self._lib.b2_launch_count() - n0)
    torch.cuda.current_stream(self._dev).wait_stream(side)
    self._graph = g
    self._capture_epoch = getattr(self, "_capture_epoch", 0) + 1   # callers that captured _enqueue_step themselves must re-capture

  #: kernels enqueued per env step, counted by the library while the step was captured: action, per chunk decimation x
  #: (smooth, order, solve, post, airtime), env_post + gate, per chunk forward (4), obs, log
  launches_per_step = property(lambda self: getattr(self, "_launches_per_step", 1 + 5 * self.cfg.decimation + 2 + 4 + 1 + 1))

  # ---- API ------------------------------------------------------------------------------------------
  def seed(self, seed: int = -1) -> int:
    if seed == -1:
      seed = int(np.random.randint(0, 10_000))
    self.vcfg.seed = int(seed)
    self._ce.seed = int(seed) & (2**64 - 1)
    torch.manual_seed(seed)
    self._capture()
    return seed

  def reset(self, *, seed: int | None = None, env_ids: torch.Tensor | None = None, options: dict | None = None):
    """manager_based_rl_env.py:254-270: _reset_idx(env_ids), write_data_to_sim, forward, observations."""
    del options
    if seed is not None:
      self.seed(seed)
    lib, m, d, e = self._lib, C.byref(self.sim._cm), C.byref(self.sim._cd), C.byref(self._ce)
    mask = None
    if env_ids is not None:
      mt = torch.zeros(self.num_envs, dtype=torch.uint8, device=self._dev)
      mt[env_ids] = 1
      mask = mt.data_ptr()
    st = self._stream()
    E.check_call(lib, lib.b2_env_reset(m, d, e, mask, st))
    E.check_call(lib, lib.b2_forward(m, d, st))
    E.check_call(lib, lib.b2_env_obs(m, d, e, 0, st))
    # advance the draw counter: the Philox draws are indexed by (seed, env, counter, slot), so without this reset() and the
    # first step() would share their draws (identical observation noise; an env timing out on step 0 would be reset to exactly
    # its initial state), and repeated reset() calls would reproduce the same state
    self._buf["step_counter"].add_(1)
    if self._obs_pipeline is not None:
      self._obs_pipeline.reset(env_ids)
      self.obs_buf = self._obs_pipeline.compute(self._raw_obs, update_history=True)
    self._publish_log()
    return self.obs_buf, self.extras

  def step(self, action: torch.Tensor):
    """One env step for all envs (manager_based_rl_env.py:272-313)."""
    if action.shape[-1] != self.vcfg.nj:
      raise ValueError(f"Invalid action shape, expected: {self.vcfg.nj}, received: {action.shape[-1]}.")
    self._buf["action_in"].copy_(action, non_blocking=True)
    self.sim.flush_memo()
    if self.sim.nan_guard.enabled:
      self._step_guarded()
    elif self._graph is not None:
      self._graph.replay()
    else:
      self._enqueue_step()
    self.common_step_counter += 1
    self._sim_step_counter += self.cfg.decimation
    if self._next_stage < len(self.cfg.velocity_stages) and self.common_step_counter > self.cfg.velocity_stages[self._next_stage]["step"]:
      self._apply_velocity_stages()
    if self._next_reward_stage_step is not None and self.common_step_counter > self._next_reward_stage_step:
      self._apply_reward_weight_stages()
    if self._obs_pipeline is not None:
      # _reset_idx resets the observation manager's buffers of the envs that were reset (manager_based_rl_env.py:392-416) before the
      # observations of this step are pushed; one host sync for the ids - this optional path is outside the fused graph anyway
      ids = self._buf["reset_buf"].nonzero().flatten()
      if ids.numel():
        self._obs_pipeline.reset(ids)
      self.obs_buf = self._obs_pipeline.compute(self._raw_obs, update_history=True)
    self._publish_log()
    return self.obs_buf, self.reward_buf, self.reset_terminated, self.reset_time_outs, self.extras

  def note_steps(self, n: int) -> None:
    """Host-side bookkeeping for `n` env steps that were enqueued through `_enqueue_step` by the caller (a roll-out graph)."""
    self.common_step_counter += n
    self._sim_step_counter += n * self.cfg.decimation
    if self._next_stage < len(self.cfg.velocity_stages) and self.common_step_counter > self.cfg.velocity_stages[self._next_stage]["step"]:
      self._apply_velocity_stages()
    if self._next_reward_stage_step is not None and self.common_step_counter > self._next_reward_stage_step:
      self._apply_reward_weight_stages()
    self._publish_log()

  def _publish_log(self) -> None:
    if self.cfg.log_metrics:
      self.extras["log"] = _LogView(self, self._buf["log"].clone())

  def get_observations(self):
    return self.obs_buf

  def render(self):
    return None

  def close(self) -> None:
    self._graph = None

  def update_reward_weight(self, term: str | int, weight: float) -> None:
    """Reward-weight curriculum hook (tasks/velocity/mdp/curriculums.py:95-107 `reward_weight`): weights are by-value kernel
    parameters, so a change re-captures the step graph (a handful of times in a training run)."""
    k = term if isinstance(term, int) else envcfg.REWARD_NAMES.index(term)
    if float(self.vcfg.reward_weights[k]) != float(weight):
      self.vcfg.reward_weights[k] = float(weight)
      A.fill_params(self._ce, self.vcfg)
      self._capture()

  def _apply_reward_weight_stages(self) -> None:
    """every stage whose step has been passed applies, in order: the last one passed wins (as the reference's loop)"""
    nxt = None
    for slot, stages in self.vcfg.extra.get("reward_weight_stages", ()):
      w = None
      for step, weight in stages:
        if self.common_step_counter > step:
          w = weight
        else:
          nxt = step if nxt is None else min(nxt, step)
      if w is not None:
        self.update_reward_weight(slot, w)
    self._next_reward_stage_step = nxt

  def _apply_velocity_stages(self) -> None:
    """commands_vel (tasks/velocity/mdp/curriculums.py:72-99): every stage whose step has been passed applies, in order."""
    lx = ly = az = None
    while self._next_stage < len(self.cfg.velocity_stages) and self.common_step_counter > self.cfg.velocity_stages[self._next_stage]["step"]:
      st = self.cfg.velocity_stages[self._next_stage]
      lx, ly, az = st.get("lin_vel_x") or lx, st.get("lin_vel_y") or ly, st.get("ang_vel_z") or az
      self._next_stage += 1
    new = (lx or self.vcfg.cmd_lin_x, ly or self.vcfg.cmd_lin_y, az or self.vcfg.cmd_ang_z)
    if tuple(map(tuple, new)) != (tuple(self.vcfg.cmd_lin_x), tuple(self.vcfg.cmd_lin_y), tuple(self.vcfg.cmd_ang_z)):
      self.update_command_ranges(*new)

  def update_command_ranges(self, lin_vel_x=None, lin_vel_y=None, ang_vel_z=None) -> None:
    """Command curriculum hook (tasks/velocity/mdp/curriculums.py `commands_vel`): ranges are by-value kernel
    parameters, so a change re-captures the step graph (it happens twice in a training run)."""
    for name, v in (("cmd_lin_x", lin_vel_x), ("cmd_lin_y", lin_vel_y), ("cmd_ang_z", ang_vel_z)):
      if v is not None:
        setattr(self.vcfg, name, tuple(v))
    A.fill_params(self._ce, self.vcfg)
    self._capture()
