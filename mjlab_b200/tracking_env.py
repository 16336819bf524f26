"""`ManagerBasedRlEnv` for the motion-tracking task (`Mjlab-Tracking-Flat-Unitree-G1`) on the B200 engine.

Mirror of the reference env configured by src/mjlab/tasks/tracking/tracking_env_cfg.py:44-317 and
config/g1/env_cfgs.py:15-100 (BeyondMimic re-implementation): the `MotionCommand` (src/mjlab/tasks/tracking/mdp/commands.py)
plays a reference clip, the policy tracks it.  Same public surface as `mjlab_b200.env.ManagerBasedRlEnv`; one env step is
one CUDA graph of the kernels behind include/b200trk.h:

    b2_trk_action -> decimation x b2_step -> b2_trk_post (terminations, 9 rewards, resets incl. start-frame sampling)
    -> conditional forward -> b2_trk_obs (metrics, frame advance, anchor-relative targets, push, observations) -> b2_trk_log

Startup domain randomisation as in the reference (tracking_env_cfg.py:179-208): torso COM offset (`body_ipos`, add),
encoder bias, foot friction.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Any, Literal

import numpy as np
import torch

from . import engine as E
from . import envcfg
from . import trkabi as A
from .entity import Entity
from .env import _TORCH, env_origins_grid
from .motion import MotionLoader, synthetic_motion
from .sim import MujocoCfg, Simulation, SimulationCfg


def _default_sim_cfg() -> SimulationCfg:
  # tasks/tracking/tracking_env_cfg.py:305-313
  return SimulationCfg(nconmax=35, njmax=250, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20))


@dataclass
class TrackingEnvCfg:
  num_envs: int = 4096
  play: bool = False
  seed: int = 42
  decimation: int = 4
  env_spacing: float = 2.0
  motion_file: str | None = None          # npz in the reference's format; None -> the synthetic clip of SURVEY 8(d)
  synthetic_frames: int = 500
  sampling_mode: Literal["adaptive", "uniform", "start"] | None = None   # None: "adaptive" (train) / "start" (play)
  domain_randomization: bool = True
  forward_mode: Literal["reference", "masked"] = "reference"
  log_metrics: bool = True
  sim: SimulationCfg = field(default_factory=_default_sim_cfg)
  manager_cfg: Any = None   # a reference ManagerBasedRlEnvCfg of the tracking task (or its cfg_to_dict view): see from_manager_cfg
  # False: the registered "Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation" task (tasks/tracking/config/g1/__init__.py:15-22) - the
  # policy group without motion_anchor_pos_b and base_lin_vel (built-in table; a manager_cfg carries its own groups)
  has_state_estimation: bool = True

  @classmethod
  def from_manager_cfg(cls, cfg, **kw) -> "TrackingEnvCfg":
    """The tracking env driven by the reference's own configuration object (`load_env_cfg("Mjlab-Tracking-Flat-Unitree-G1")`,
    edited as usual): reward weights / stds, termination thresholds, MotionCommand bodies and ranges, noise, events, episode
    length are resolved by `envcfg.tracking_from_manager_cfg`; unsupported terms raise `envcfg.UnsupportedTermError`."""
    d = cfg if isinstance(cfg, dict) else envcfg.cfg_to_dict(cfg)
    mj = d["sim"]["mujoco"]
    sim = _default_sim_cfg()
    sim.mujoco.timestep, sim.mujoco.iterations, sim.mujoco.ls_iterations = float(mj["timestep"]), int(mj["iterations"]), int(mj["ls_iterations"])
    sim.mujoco.integrator = mj["integrator"]
    if d["sim"].get("nconmax"):
      sim.nconmax = int(d["sim"]["nconmax"])
    if d["sim"].get("njmax"):
      sim.njmax = int(d["sim"]["njmax"])
    mf = [t for t in d["commands"].values() if t is not None][0].get("motion_file") or None
    kw.setdefault("motion_file", mf)
    return cls(decimation=int(d["decimation"]), sim=sim, manager_cfg=d, **kw)


class TrackingRlEnv:
  is_vector_env = True

  def __init__(self, cfg: TrackingEnvCfg, device: str = "cuda:0", render_mode: str | None = None, vcfg_edit=None, **kwargs: Any):
    """`vcfg_edit(resolved_cfg)` may adjust the resolved task parameters before the buffers are sized (tests, curricula)."""
    from . import load_model
    self.cfg, self.device, self.render_mode = cfg, device, render_mode
    self._lib = A.load_library()
    model = load_model("g1_flat")
    dev = torch.device(device)
    self._dev = dev
    if cfg.motion_file is not None:
      motion = dict(np.load(cfg.motion_file))
    else:
      motion = synthetic_motion(model, cfg.synthetic_frames, device=device)
    self.sim = Simulation(cfg.num_envs, cfg.sim, model, device)
    if cfg.manager_cfg is not None:
      self.vcfg = envcfg.tracking_from_manager_cfg(cfg.manager_cfg, model, motion["joint_pos"].shape[0])
    else:
      self.vcfg = envcfg.tracking_cfg(model, motion["joint_pos"].shape[0], play=cfg.play)
      if not cfg.has_state_estimation:
        envcfg.tracking_without_state_estimation(self.vcfg)
    self.vcfg.seed = int(cfg.seed)
    if cfg.sampling_mode is not None:
      self.vcfg.sampling_mode = cfg.sampling_mode
    self.vcfg.step_dt = cfg.decimation * cfg.sim.mujoco.timestep
    if vcfg_edit is not None:
      vcfg_edit(self.vcfg)
    n = cfg.num_envs
    torch.manual_seed(cfg.seed)
    self.motion = MotionLoader(motion, torch.as_tensor(self.vcfg.body_sel, device=dev), device=device)

    self._ce = A.B2Trk()
    A.fill_params(self._ce, self.vcfg)
    tables = A.table_values(self.vcfg)
    self._buf: dict[str, torch.Tensor] = {}
    for i, (_, attr, kind, dt, shp) in enumerate(A.BUFFERS):
      if kind == "table":
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(tables[attr])), device=dev).to(_TORCH[dt]).contiguous()
      elif kind == "motion":
        t = torch.as_tensor(np.ascontiguousarray(motion[attr]), dtype=torch.float32, device=dev).contiguous()
      elif kind == "state":
        t = torch.zeros((n,) + tuple(shp(self.vcfg)), dtype=_TORCH[dt], device=dev)
      else:
        t = torch.zeros(tuple(shp(self.vcfg)), dtype=_TORCH[dt], device=dev)
      self._buf["m_" + attr if kind == "motion" else attr] = t
      self._ce.buf[i] = t.data_ptr()
    self._buf["rel_quat"][..., 0] = 1.0
    self._buf["env_origins"][:] = torch.as_tensor(env_origins_grid(n, cfg.env_spacing), device=dev)
    if self.vcfg.push_enabled:
      lo, hi = self.vcfg.push_interval
      self._buf["push_time_left"][:] = torch.rand(n, device=dev) * (hi - lo) + lo
    E.check_call(self._lib, self._lib.b2_trk_check(C.byref(self.sim._cm), C.byref(self.sim._cd), C.byref(self._ce)))

    from .robots import g1_cfg
    self.robot = Entity("robot", self.sim, joint_pos_target=self._buf["joint_pos_target"], encoder_bias=self._buf["encoder_bias"],
                        actuator_groups=[a.joint_names_expr for a in g1_cfg()[0]])
    self.scene = {"robot": self.robot}
    self.extras: dict[str, Any] = {"log": {}}
    self.common_step_counter = 0
    self.obs_buf = {"policy": self._buf["obs_policy"], "critic": self._buf["obs_critic"]}
    self._raw_obs = dict(self.obs_buf)     # the kernels' term bank; a consumed cfg with its own groups gathers from it
    self._obs_pipeline = None
    ex = getattr(self.vcfg, "extra", None) or {}
    if ex.get("obs_custom"):
      from .obs_pipeline import ObservationPipeline
      self._obs_pipeline = ObservationPipeline(ex["obs_layout"], ex["obs_post"], n, device)
    self.reward_buf = self._buf["reward"]
    self.reset_buf = self._buf["reset_buf"].view(torch.bool)
    self.reset_terminated = self._buf["terminated"].view(torch.bool)
    self.reset_time_outs = self._buf["time_outs"].view(torch.bool)

    if cfg.manager_cfg is not None:   # the cfg's own startup events, in order (any field events.randomize_field knows)
      from .events import apply_startup_event
      for ev in self.vcfg.extra["startup_dr"]:
        apply_startup_event(self, ev)
      br = self.vcfg.extra["encoder_bias_range"]
      if br is not None:   # randomize_encoder_bias, events.py:709-744
        self._buf["encoder_bias"][:] = torch.rand(n, self.vcfg.nj, device=dev) * (br[1] - br[0]) + br[0]
    elif cfg.domain_randomization:   # startup events (tracking_env_cfg.py:179-208)
      from .events import randomize_field
      self.sim.expand_model_fields(("body_ipos", "geom_friction"))
      randomize_field(self, None, "body_ipos", {0: (-0.025, 0.025), 1: (-0.05, 0.05), 2: (-0.05, 0.05)}, operation="add", body_names=("torso_link",))
      randomize_field(self, None, "geom_friction", (0.3, 1.2), operation="abs", geom_names=r"^(left|right)_foot[1-7]_collision$")
      self._buf["encoder_bias"][:] = torch.rand(n, self.vcfg.nj, device=dev) * 0.02 - 0.01   # randomize_encoder_bias, events.py:709-744

    self._graph: torch.cuda.CUDAGraph | None = None
    self.use_cuda_graph = os.environ.get("B200SIM_NO_GRAPH", "0") != "1"
    self._capture()
    self.sim.add_graph_listener(self._capture)   # sim.expand_model_fields after construction: the step graph follows

  num_envs = property(lambda self: self.cfg.num_envs)
  physics_dt = property(lambda self: self.cfg.sim.mujoco.timestep)
  step_dt = property(lambda self: self.cfg.sim.mujoco.timestep * self.cfg.decimation)
  max_episode_length_s = property(lambda self: self.vcfg.max_episode_length_s)
  max_episode_length = property(lambda self: self.vcfg.max_episode_length)
  unwrapped = property(lambda self: self)
  launches_per_step = property(lambda self: getattr(self, "_launches_per_step", 1 + 4 * self.cfg.decimation + 2 + 4 + 1 + 1))

  @property
  def episode_length_buf(self) -> torch.Tensor:
    return self._buf["episode_length"]

  @episode_length_buf.setter
  def episode_length_buf(self, value: torch.Tensor) -> None:
    self._buf["episode_length"].copy_(value)

  # minimal manager facades (same attribute names as the reference's managers)
  class _AM:
    def __init__(self, env):
      self._e = env
    total_action_dim = property(lambda self: self._e.vcfg.nj)
    action = property(lambda self: self._e._buf["action"])
    prev_action = property(lambda self: self._e._buf["prev_action"])

  class _OM:
    def __init__(self, env):
      self._e = env

    def compute(self, update_history: bool = False):
      return self._e.obs_buf

    @property
    def active_terms(self):
      p = self._e._obs_pipeline
      if p is not None:
        return {g: [t["name"] for t in ts] for g, ts in p.groups.items()}
      pol = ["command", "motion_anchor_pos_b", "motion_anchor_ori_b", "base_lin_vel", "base_ang_vel", "joint_pos", "joint_vel", "actions"]
      return {"policy": pol, "critic": pol[:3] + ["body_pos", "body_ori"] + pol[3:]}

    @property
    def group_obs_dim(self):
      return {g: tuple(v.shape[1:]) for g, v in self._e.obs_buf.items()}

  class _CM:
    def __init__(self, env):
      self._e = env
    time_steps = property(lambda self: self._e._buf["time_steps"])
    body_pos_relative_w = property(lambda self: self._e._buf["rel_pos"])
    body_quat_relative_w = property(lambda self: self._e._buf["rel_quat"])
    bin_failed_count = property(lambda self: self._e._buf["bin_failed_count"])
    metrics = property(lambda self: {n: self._e._buf["metrics"][:, k] for k, n in enumerate(
      ("error_anchor_pos", "error_anchor_rot", "error_anchor_lin_vel", "error_anchor_ang_vel", "error_body_pos", "error_body_rot",
       "error_joint_pos", "error_joint_vel"))})

    def get_command(self, name: str) -> torch.Tensor:
      e = self._e
      t = e._buf["time_steps"]
      return torch.cat([e.motion.joint_pos[t], e.motion.joint_vel[t]], dim=1)

  class _TM:
    active_terms = list(envcfg.TRACKING_TERMINATION_NAMES)

    def __init__(self, env):
      self._e = env
    terminated = property(lambda self: self._e._buf["terminated"].view(torch.bool))
    time_outs = property(lambda self: self._e._buf["time_outs"].view(torch.bool))

    def get_term(self, name: str) -> torch.Tensor:
      return self._e._buf["term_dones"][:, self.active_terms.index(name)].view(torch.bool)

  class _RM:
    active_terms = list(envcfg.TRACKING_REWARD_NAMES)

    def __init__(self, env):
      self._e = env
    _step_reward = property(lambda self: self._e._buf["step_reward"])
    _episode_sums = property(lambda self: {n: self._e._buf["episode_sums"][:, k] for k, n in enumerate(self.active_terms)})

  action_manager = property(lambda self: TrackingRlEnv._AM(self))
  observation_manager = property(lambda self: TrackingRlEnv._OM(self))
  command_manager = property(lambda self: TrackingRlEnv._CM(self))
  termination_manager = property(lambda self: TrackingRlEnv._TM(self))
  reward_manager = property(lambda self: TrackingRlEnv._RM(self))

  def _stream(self) -> int:
    return torch.cuda.current_stream(self._dev).cuda_stream

  def _enqueue_step(self) -> None:
    lib, m, d, e, chk = self._lib, C.byref(self.sim._cm), C.byref(self.sim._cd), C.byref(self._ce), E.check_call
    st = self._stream()
    chk(lib, lib.b2_trk_action(m, d, e, st))
    chk(lib, lib.b2_substeps_after_forward(m, d, self.cfg.decimation, st))
    chk(lib, lib.b2_trk_post(m, d, e, st))
    mask = self._buf["reset_buf"].data_ptr() if self.cfg.forward_mode == "masked" else None
    chk(lib, lib.b2_forward_masked(m, d, mask, self._buf["any_reset"].data_ptr(), st))
    chk(lib, lib.b2_trk_obs(m, d, e, 1, st))
    chk(lib, lib.b2_trk_log(d, e, st))

  def _capture(self) -> None:
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

  def seed(self, seed: int = -1) -> int:
    if seed == -1:
      seed = int(np.random.randint(0, 10_000))
    self.vcfg.seed = int(seed)
    self._ce.seed = int(seed) & (2**64 - 1)
    torch.manual_seed(seed)
    self._capture()
    return seed

  def reset(self, *, seed: int | None = None, env_ids: torch.Tensor | None = None, options: dict | None = None):
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
    E.check_call(lib, lib.b2_trk_reset(m, d, e, mask, st))
    E.check_call(lib, lib.b2_forward(m, d, st))
    E.check_call(lib, lib.b2_trk_obs(m, d, e, 0, st))
    self._buf["step_counter"].add_(1)   # reset() consumes one draw-counter value (see env.ManagerBasedRlEnv.reset)
    if self._obs_pipeline is not None:
      self._obs_pipeline.reset(env_ids)
      self.obs_buf = self._obs_pipeline.compute(self._raw_obs, update_history=True)
    self._publish_log()
    return self.obs_buf, self.extras

  def step(self, action: torch.Tensor):
    if action.shape[-1] != self.vcfg.nj:
      raise ValueError(f"Invalid action shape, expected: {self.vcfg.nj}, received: {action.shape[-1]}.")
    self._buf["action_in"].copy_(action, non_blocking=True)
    self.sim.flush_memo()
    if self.sim.nan_guard.enabled:
      raise NotImplementedError("NanGuard on the tracking env: enable it on the velocity env (per-physics-step checks, "
                                "env.ManagerBasedRlEnv._step_guarded) or step `sim` directly")
    if self._graph is not None:
      self._graph.replay()
    else:
      self._enqueue_step()
    self.common_step_counter += 1
    if self._obs_pipeline is not None:   # as env.ManagerBasedRlEnv.step: buffers of reset envs are cleared before this step's push
      ids = self._buf["reset_buf"].nonzero().flatten()
      if ids.numel():
        self._obs_pipeline.reset(ids)
      self.obs_buf = self._obs_pipeline.compute(self._raw_obs, update_history=True)
    self._publish_log()
    return self.obs_buf, self.reward_buf, self.reset_terminated, self.reset_time_outs, self.extras

  def note_steps(self, n: int) -> None:
    self.common_step_counter += n
    self._publish_log()

  def _publish_log(self) -> None:
    if self.cfg.log_metrics:
      self.extras["log"] = _TrkLog(self, self._buf["log"].clone())

  def get_observations(self):
    return self.obs_buf

  def close(self) -> None:
    self._graph = None


class _TrkLog(dict):
  """extras["log"]: resolved lazily from a snapshot of the device log sums (one small D2H on first access)."""

  def __init__(self, env: TrackingRlEnv, snap: torch.Tensor):
    super().__init__()
    self._env, self._snap, self._done = env, snap, False

  def _resolve(self):
    if not self._done:
      v, K, cfg = self._snap.tolist(), A.K, self._env.vcfg
      n = max(v[K["B2T_LOG_NRESET"]], 1.0)
      for k, name in enumerate(envcfg.TRACKING_REWARD_NAMES):
        super().__setitem__("Episode_Reward/" + name, v[K["B2T_LOG_EPISODE_SUM"] + k] / n / cfg.max_episode_length_s)
      for k, name in enumerate(envcfg.TRACKING_TERMINATION_NAMES):
        super().__setitem__("Episode_Termination/" + name, v[K["B2T_LOG_TERM_COUNT"] + k])
      for k, name in enumerate(("error_anchor_pos", "error_anchor_rot", "error_anchor_lin_vel", "error_anchor_ang_vel", "error_body_pos",
                                "error_body_rot", "error_joint_pos", "error_joint_vel")):
        super().__setitem__("Metrics/motion/" + name, v[K["B2T_LOG_METRIC"] + k] / n)
      super().__setitem__("Episode/num_resets", v[K["B2T_LOG_NRESET"]])
      self._done = True

  def __getitem__(self, k):
    self._resolve()
    return super().__getitem__(k)

  def __iter__(self):
    self._resolve()
    return super().__iter__()

  def __len__(self):
    self._resolve()
    return super().__len__()

  def keys(self):
    self._resolve()
    return super().keys()

  def items(self):
    self._resolve()
    return super().items()
