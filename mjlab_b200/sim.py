"""`Simulation`: host-side mirror of `mjlab.sim.Simulation` on the B200 engine.

Same names, argument meaning and error behaviour as the reference class
(src/mjlab/sim/sim.py:100-286) and its tensor bridge (src/mjlab/sim/sim_data.py):

  * `Simulation(num_envs, cfg, model, device)`, `.step()`, `.forward()`,
    `.reset(env_ids)`, `.create_graph()`, `.expand_model_fields(fields)`,
    `.get_default_field(name)`, `.model` / `.data` bridges, `.mj_model`
  * model tensors carry a leading world axis: shared fields are (1, ...) storage
    expanded to nworld with stride 0 (sim_data.py:23-31) until
    `expand_model_fields` makes them per-world; that reallocates, so the CUDA
    graphs are re-captured (sim.py:226-240)
  * bridges are read-only attribute bags with stable device addresses
    (sim_data.py:217-223; tests/test_sim_data.py:56-75)

Differences, on purpose: there is no CPU device (the CUDA
extension is mandatory), and everything runs on torch's current stream, so the
Warp-stream redirection of the reference is not needed.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Literal

import numpy as np
import torch

from . import engine as E
from . import mjcf
from .nan_guard import NanGuard, NanGuardCfg

_INTEGRATOR_MAP = {"euler": mjcf.INT_EULER, "implicitfast": mjcf.INT_IMPLICITFAST}
_CONE_MAP = {"pyramidal": mjcf.CONE_PYRAMIDAL, "elliptic": mjcf.CONE_ELLIPTIC}
_SOLVER_MAP = {"newton": mjcf.SOL_NEWTON, "cg": mjcf.SOL_CG, "pgs": mjcf.SOL_PGS}


@dataclass
class MujocoCfg:
  """Physics options (reference: sim.py:43-79)."""

  timestep: float = 0.002
  integrator: Literal["euler", "implicitfast"] = "implicitfast"
  impratio: float = 1.0
  cone: Literal["pyramidal", "elliptic"] = "pyramidal"
  jacobian: Literal["auto", "dense", "sparse"] = "auto"
  solver: Literal["newton", "cg", "pgs"] = "newton"
  iterations: int = 100
  tolerance: float = 1e-8
  ls_iterations: int = 50
  ls_tolerance: float = 0.01
  ccd_iterations: int = 50
  gravity: tuple[float, float, float] = (0, 0, -9.81)

  def apply(self, model: mjcf.ModelDesc) -> None:
    model.opt_cone = _CONE_MAP[self.cone]
    model.opt_integrator = _INTEGRATOR_MAP[self.integrator]
    model.opt_solver = _SOLVER_MAP[self.solver]
    model.opt_timestep = float(self.timestep)
    model.opt_impratio = float(self.impratio)
    model.opt_gravity = np.array(self.gravity, float)
    model.opt_iterations = int(self.iterations)
    model.opt_tolerance = float(self.tolerance)
    model.opt_ls_iterations = int(self.ls_iterations)
    model.opt_ls_tolerance = float(self.ls_tolerance)
    model.opt_ccd_iterations = int(self.ccd_iterations)   # piped like the reference (sim.py:66-79); no CCD in the supported subset


@dataclass(kw_only=True)
class SimulationCfg:
  """Reference: sim.py:82-97."""

  nconmax: int | None = None
  njmax: int | None = None
  ls_parallel: bool = True  # accepted for API parity; the engine always runs the exact iterative line search
  export_efc_J: bool = False  # also write the dense constraint Jacobian to HBM every step (off on the hot path)
  contact_sensor_maxmatch: int = 64
  mujoco: MujocoCfg = field(default_factory=MujocoCfg)
  nan_guard: NanGuardCfg = field(default_factory=NanGuardCfg)   # sim.py:97; off by default


class WarpBridge:
  """Read-only attribute bag of torch tensors: the reference's `WarpBridge` (sim/sim_data.py:177-240) without Warp underneath.
  Attributes are the engine's own device buffers as `torch.Tensor`s (stable addresses for the process lifetime); assigning an
  attribute raises, in-place writes (`bridge.qpos[:] = x`, index assignment) go straight to the buffer the kernels read."""

  def __init__(self, fields: dict[str, Any], on_access: Any = None) -> None:
    object.__setattr__(self, "_fields", fields)
    object.__setattr__(self, "_on_access", on_access)

  def __getattr__(self, name: str) -> Any:
    try:
      cb = object.__getattribute__(self, "_on_access")
      if cb is not None:
        cb()
      return object.__getattribute__(self, "_fields")[name]
    except KeyError:
      raise AttributeError(name) from None

  def __setattr__(self, name: str, value: Any) -> None:
    raise AttributeError(
      f"Cannot set attribute '{name}' on WarpBridge. This wrapper is read-only to preserve memory addresses "
      f"for CUDA graphs. Use in-place operations instead: obj.{name}[:] = value"
    )

  def __contains__(self, name: str) -> bool:
    return name in self._fields

  def field_names(self) -> list[str]:
    return list(self._fields)

  def clear_cache(self) -> None:  # API parity; tensors are not cached wrappers here
    pass


TensorBridge = WarpBridge   # round-1 name


class _Opt:
  """`model.opt`: tensors for per-world options, plain attributes for the rest."""

  def __init__(self, sim: "Simulation") -> None:
    self._sim = sim
    self.ls_parallel = sim.cfg.ls_parallel
    self.contact_sensor_maxmatch = sim.cfg.contact_sensor_maxmatch

  @property
  def timestep(self) -> torch.Tensor:
    self._sim.invalidate_memo()   # the tensor may be edited in place (see Simulation.model)
    return self._sim._model_views["opt_timestep"].reshape(-1)

  @property
  def gravity(self) -> torch.Tensor:
    self._sim.invalidate_memo()   # the tensor may be edited in place (see Simulation.model)
    return self._sim._model_views["opt_gravity"]

  @property
  def tolerance(self) -> torch.Tensor:
    self._sim.invalidate_memo()   # the tensor may be edited in place (see Simulation.model)
    return self._sim._model_views["opt_tolerance"].reshape(-1)

  @property
  def ls_tolerance(self) -> torch.Tensor:
    self._sim.invalidate_memo()   # the tensor may be edited in place (see Simulation.model)
    return self._sim._model_views["opt_ls_tolerance"].reshape(-1)

  @property
  def impratio(self) -> torch.Tensor:
    self._sim.invalidate_memo()   # the tensor may be edited in place (see Simulation.model)
    return self._sim._model_views["opt_impratio"].reshape(-1)

  @property
  def integrator(self) -> int:
    return self._sim._cm.integrator

  @property
  def cone(self) -> int:
    return self._sim._cm.cone

  @property
  def solver(self) -> int:
    return self._sim._cm.solver

  @property
  def iterations(self) -> int:
    return self._sim._cm.iterations

  @property
  def ls_iterations(self) -> int:
    return self._sim._cm.ls_iterations


def _natural_shape(m: mjcf.ModelDesc, attr: str) -> tuple[int, ...]:
  v = np.atleast_1d(np.asarray(getattr(m, attr)))
  return tuple(v.shape)


class Simulation:
  """World-batched physics on one B200 (one process per GPU)."""

  def __init__(self, num_envs: int, cfg: SimulationCfg, model, device: str):
    """`model`: a `mjcf.ModelDesc`, or a compiled `mujoco.MjModel` (any object with MjModel's attribute surface), which is
    converted in-process (`ModelDesc.from_mjmodel`, the put_model half of sim/sim.py:139)."""
    if not isinstance(model, mjcf.ModelDesc):
      model = mjcf.ModelDesc.from_mjmodel(model)
    if not str(device).startswith("cuda"):
      raise RuntimeError("mjlab_b200.Simulation requires a CUDA device (there is no CPU fallback)")
    if not torch.cuda.is_available():
      raise RuntimeError("CUDA is not available: the B200 engine cannot run")
    self.cfg = cfg
    self.device = device
    self.num_envs = num_envs
    self._default_model_fields: dict[str, torch.Tensor] = {}
    self._lib = E.load_library()

    self._mj_model = model
    cfg.mujoco.apply(model)
    E.check_supported(model)
    self._tables = E.derived_int_tables(model)

    self.nconmax = int(cfg.nconmax) if cfg.nconmax is not None else max(16, 4 * model.ngeom // 3)
    self.njmax = int(cfg.njmax) if cfg.njmax is not None else max(32, 4 * self.nconmax + 2 * model.njnt + E.count_friction_dofs(model))

    dev = torch.device(device)
    self._dev = dev
    # ---- model upload (put_model) ---------------------------------------------------
    self._cm = E.B2Model()
    for k, v in E.model_scalars(model, self._tables).items():
      setattr(self._cm, k, v)
    self._model_store: dict[str, torch.Tensor] = {}   # attr -> storage tensor (1|nworld, *natural)
    self._model_views: dict[str, torch.Tensor] = {}   # attr -> (nworld, *natural) view
    self._mf_index = {attr: i for i, (_, attr, _) in enumerate(E.MODEL_FLOAT_FIELDS)}
    self._int_store: list[torch.Tensor] = []
    for i, (_, attr, _rl) in enumerate(E.MODEL_FLOAT_FIELDS):
      flat = E.model_float_array(model, attr).astype(np.float32)
      shape = _natural_shape(model, attr)
      t = torch.from_numpy(flat.copy()).to(dev).reshape((1,) + shape) if flat.size else torch.zeros((1,) + shape, device=dev)
      self._bind_model_field(attr, t)
    for i, (_, attr) in enumerate(E.MODEL_INT_FIELDS):
      a = E.model_int_array(model, attr, self._tables)
      t = torch.from_numpy(a.copy()).to(dev)
      self._int_store.append(t)
      self._cm.mi[i] = t.data_ptr()
    # every other MjModel array stays reachable through sim.model.<name> (read-only to the engine)
    for name, v in model.fields().items():
      if name in self._model_views or name in ("qM0",):
        continue
      if v.dtype.kind == "f":
        t = torch.from_numpy(np.asarray(v, np.float32).copy()).to(dev)
      else:
        t = torch.from_numpy(np.asarray(v, np.int32).copy()).to(dev)
      self._model_store[name] = t.unsqueeze(0)
      self._model_views[name] = t.unsqueeze(0).expand((num_envs,) + tuple(t.shape))
    self._opt = _Opt(self)

    # ---- data allocation (put_data) ----------------------------------------------------
    self._cd = E.B2Data()
    self._cd.nworld, self._cd.nconmax, self._cd.njmax = num_envs, self.nconmax, self.njmax
    self._cd.flags = 1 if cfg.export_efc_J else 0
    self._data_fields: dict[str, Any] = {"nworld": num_envs}
    for i, (_, attr, shp) in enumerate(E.DATA_FLOAT_FIELDS):
      shape = (num_envs,) + tuple(shp(model, self.nconmax, self.njmax))
      t = torch.zeros(shape, dtype=torch.float32, device=dev)
      if t.numel() == 0:
        pad = torch.zeros(max(num_envs, 1), dtype=torch.float32, device=dev)
        self._int_store.append(pad)
        self._cd.df[i] = pad.data_ptr()
      else:
        self._cd.df[i] = t.data_ptr()
      self._data_fields[attr] = t[..., : model.nv] if attr == "qM" else t
    for i, (_, attr, shp) in enumerate(E.DATA_INT_FIELDS):
      shape = (num_envs,) + tuple(shp(model, self.nconmax, self.njmax))
      t = torch.zeros(shape, dtype=torch.int32, device=dev)
      self._cd.di[i] = t.data_ptr()
      self._data_fields[attr] = t
    self._data_fields["qpos"][:] = torch.from_numpy(np.asarray(model.qpos0, np.float32)).to(dev)
    if model.nmocap:
      ids = [i for i in range(model.nbody) if model.body_mocapid[i] >= 0]
      self._data_fields["mocap_pos"][:] = torch.from_numpy(np.asarray(model.body_pos[ids], np.float32)).to(dev)
      self._data_fields["mocap_quat"][:] = torch.from_numpy(np.asarray(model.body_quat[ids], np.float32)).to(dev)
    self._reset_mask = torch.zeros(num_envs, dtype=torch.bool, device=dev)

    E.check_call(self._lib, self._lib.b2_check(C.byref(self._cm), C.byref(self._cd)))
    # static geoms (terrain): their world pose is written once; the step kernels only update moving geoms
    E.check_call(self._lib, self._lib.b2_init_static(C.byref(self._cm), C.byref(self._cd), self._stream()))

    self._memo_dirty = True
    self._model_bridge = TensorBridge(_ModelFields(self), on_access=self.invalidate_memo)
    self._data_bridge = TensorBridge(self._data_fields)
    self.nan_guard = NanGuard(cfg.nan_guard, num_envs, model)   # sim.py:162

    self.use_cuda_graph = True
    self.step_graph = None
    self.forward_graph = None
    self.reset_graph = None
    self.forward()  # mj_forward at construction, like put_data after mj_forward (sim.py:134-149)
    self.create_graph()

  # -- internals ----------------------------------------------------------------------
  def _bind_model_field(self, attr: str, storage: torch.Tensor) -> None:
    """storage: (1|nworld, *natural) contiguous float32 on the device."""
    i = self._mf_index[attr]
    self._model_store[attr] = storage
    rows = storage.shape[0]
    row_len = storage[0].numel()
    self._cm.mf[i] = storage.data_ptr() if storage.numel() else 0
    self._cm.mf_stride[i] = row_len if rows > 1 else 0
    if storage.numel() == 0:
      pad = torch.zeros(1, dtype=torch.float32, device=self._dev)
      self._int_store.append(pad)
      self._cm.mf[i] = pad.data_ptr()
    self._model_views[attr] = storage if rows == self.num_envs and rows > 1 else storage.expand((self.num_envs,) + tuple(storage.shape[1:]))

  def _stream(self) -> int:
    return torch.cuda.current_stream(self._dev).cuda_stream

  def _launch_step(self) -> None:
    E.check_call(self._lib, self._lib.b2_step(C.byref(self._cm), C.byref(self._cd), self._stream()))

  def _launch_forward(self) -> None:
    E.check_call(self._lib, self._lib.b2_forward(C.byref(self._cm), C.byref(self._cd), self._stream()))

  def _launch_reset(self) -> None:
    E.check_call(self._lib, self._lib.b2_reset(C.byref(self._cm), C.byref(self._cd), self._reset_mask.data_ptr(), self._stream()))

  # -- graphs -------------------------------------------------------------------------
  def create_graph(self) -> None:
    """Capture CUDA graphs for step / forward / masked reset (sim.py:164-191).

    Must be re-run whenever a model or data buffer is replaced; called by
    `__init__` and `expand_model_fields`.
    """
    self.step_graph = self.forward_graph = self.reset_graph = None
    if not self.use_cuda_graph:
      self._notify_graph_listeners()
      return
    torch.cuda.synchronize(self._dev)
    side = torch.cuda.Stream(self._dev)
    with torch.cuda.stream(side):
      # warm-up launch outside capture (sets function attributes, loads modules)
      graphs = []
      for fn in (self._launch_step, self._launch_forward, self._launch_reset):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
          fn()
        graphs.append(g)
    torch.cuda.current_stream(self._dev).wait_stream(side)
    self.step_graph, self.forward_graph, self.reset_graph = graphs
    self._notify_graph_listeners()

  def _notify_graph_listeners(self) -> None:
    # graphs captured by others over this simulation's buffers (the env step graph) hold the same addresses / by-value
    # parameters and must be re-captured with ours
    for cb in list(getattr(self, "_graph_listeners", ())):
      cb()

  def add_graph_listener(self, cb) -> None:
    """`cb()` runs after every `create_graph()` (i.e. after `expand_model_fields` replaced a model buffer)."""
    if not hasattr(self, "_graph_listeners"):
      self._graph_listeners = []
    self._graph_listeners.append(cb)

  # -- properties -----------------------------------------------------------------------
  @property
  def mj_model(self) -> mjcf.ModelDesc:
    return self._mj_model

  @property
  def data(self) -> WarpBridge:
    return self._data_bridge

  @property
  def model(self) -> WarpBridge:
    # Handing out a model tensor (any attribute access on this bridge) may be followed by an in-place edit the engine cannot
    # see; the env step's reuse of the forward pass (b2_substeps_after_forward) compares qpos / qvel only, so the bridge
    # calls invalidate_memo() on every access.
    return self._model_bridge

  def invalidate_memo(self) -> None:
    """Forget the position / velocity stage a forward pass left behind (needed after MODEL fields were edited through a tensor
    obtained earlier than the previous env step; `sim.model.<field>` accesses do it by themselves)."""
    self._memo_dirty = True

  def flush_memo(self) -> None:
    """Called by the envs before they replay their step graph."""
    if self._memo_dirty:
      self._data_fields["smooth_memo"].zero_()
      self._memo_dirty = False

  # the reference exposes the raw mjwarp structs next to the torch bridges (sim.py:193-207); here they are the same objects
  @property
  def wp_data(self) -> WarpBridge:
    return self._data_bridge

  @property
  def wp_model(self) -> WarpBridge:
    return self._model_bridge

  @property
  def default_model_fields(self) -> dict[str, torch.Tensor]:
    return self._default_model_fields

  # -- methods --------------------------------------------------------------------------
  def expand_model_fields(self, fields: tuple[str, ...]) -> None:
    """Give the listed model fields one row per world (sim.py:226-240)."""
    if not fields:
      return
    invalid = [f for f in fields if not hasattr(self._mj_model, f)]
    if invalid:
      raise ValueError(f"Fields not found in model: {invalid}")
    dead = [f for f in fields if f in E.UNCONSUMED_PHYSICS_FIELDS]
    if dead:
      raise ValueError(f"Fields {dead} are not consumed by the engine; randomising them would have no effect")
    if self.num_envs == 1:
      return
    for f in fields:
      src = self._model_store[f]
      if src.shape[0] == self.num_envs:
        continue
      dst = torch.empty((self.num_envs,) + tuple(src.shape[1:]), dtype=src.dtype, device=self._dev)
      if src.dtype == torch.float32 and src.numel():
        E.check_call(self._lib, self._lib.b2_repeat_rows(dst.data_ptr(), src.data_ptr(), src[0].numel(), self.num_envs, self._stream()))
      else:
        dst.copy_(src.expand_as(dst))
      if f in self._mf_index:
        self._bind_model_field(f, dst)
      else:
        self._model_store[f] = dst
        self._model_views[f] = dst
    if "dof_frictionloss" in fields and self._cm.nfriction < self._mj_model.nv:
      # per-world friction loss: any dof may get a row in some world, so the solver is planned for nv friction-loss rows
      # (and the friction-aware kernel instantiation is launched even if the compiled model had frictionloss = 0)
      self._cm.nfriction = int(self._mj_model.nv)
      E.check_call(self._lib, self._lib.b2_check(C.byref(self._cm), C.byref(self._cd)))
    self._model_bridge.clear_cache()
    self.create_graph()

  def get_default_field(self, name: str) -> torch.Tensor:
    """Compiled-model value of a field, cached (sim.py:242-258)."""
    if name not in self._default_model_fields:
      if not hasattr(self._mj_model, name):
        raise ValueError(f"Field '{name}' not found in model")
      model_field = getattr(self.model, name)
      self._default_model_fields[name] = torch.as_tensor(
        np.asarray(getattr(self._mj_model, name)), dtype=model_field.dtype, device=self.device
      ).clone()
    return self._default_model_fields[name]

  def forward(self) -> None:
    if self.use_cuda_graph and self.forward_graph is not None:
      self.forward_graph.replay()
    else:
      self._launch_forward()

  def step(self) -> None:
    with self.nan_guard.watch(self.data):      # sim.py:267-273; a no-op unless cfg.nan_guard.enabled
      if self.use_cuda_graph and self.step_graph is not None:
        self.step_graph.replay()
      else:
        self._launch_step()

  def reset(self, env_ids: torch.Tensor | None = None) -> None:
    if env_ids is None:
      self._reset_mask.fill_(True)
    else:
      self._reset_mask.fill_(False)
      self._reset_mask[env_ids] = True
    if self.use_cuda_graph and self.reset_graph is not None:
      self.reset_graph.replay()
    else:
      self._launch_reset()


class _ModelFields(dict):
  """Lazy dict view used by the model bridge: resolves MjModel names and `opt`."""

  def __init__(self, sim: Simulation) -> None:
    super().__init__()
    self._sim = sim

  def __getitem__(self, name: str) -> Any:
    if name == "opt":
      return self._sim._opt
    if name == "nworld":
      return self._sim.num_envs
    v = self._sim._model_views.get(name)
    if v is not None:
      return v
    if hasattr(self._sim._mj_model, name):
      return getattr(self._sim._mj_model, name)
    raise KeyError(name)

  def __contains__(self, name: object) -> bool:
    try:
      self[name]  # type: ignore[index]
      return True
    except KeyError:
      return False

  def __iter__(self):
    return iter(self._sim._model_views)
