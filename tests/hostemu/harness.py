"""TEST-ONLY harness around the single-lane host build of the device code.

Builds tests/hostemu/libhostemu_{f32,f64}.so with g++ from the same .cuh
sources the CUDA library compiles, and drives it with numpy buffers.  It lets
the kernel logic be compared with the oracle without a GPU.  Never imported by
the product package.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from mjlab_b200 import engine as E
from mjlab_b200 import envabi as A
from mjlab_b200 import trkabi as TA

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]


def build(real: str = "float") -> Path:
  tag = "f32" if real == "float" else "f64"
  out = HERE / f"libhostemu_{tag}.so"
  srcs = [HERE / "hostemu.cpp"] + sorted((ROOT / "mjlab_b200" / "csrc").glob("*.cuh")) + [ROOT / "include" / "b200sim.h"]
  if out.exists() and all(out.stat().st_mtime >= s.stat().st_mtime for s in srcs):
    return out
  cmd = ["g++", "-O2", "-std=c++17", "-DB2_HOSTEMU", f"-DB2_REAL={real}", "-shared", "-fPIC", "-ffp-contract=off",
         "-Wno-unused-variable", "-o", str(out), str(HERE / "hostemu.cpp")]
  subprocess.run(cmd, check=True)
  return out


class HostSim:
  """Numpy-backed model/data driven through the host-emulated kernels."""

  def __init__(self, m, nworld: int, nconmax: int = 64, njmax: int = 300, real: str = "float"):
    self.m = m
    self.dtype = np.float32 if real == "float" else np.float64
    self.lib = C.CDLL(str(build(real)))
    assert self.lib.b2h_sizeof_real() == np.dtype(self.dtype).itemsize
    assert self.lib.b2h_sizeof_model() == C.sizeof(E.B2Model)
    assert self.lib.b2h_sizeof_data() == C.sizeof(E.B2Data)
    E.check_supported(m)
    self.tables = E.derived_int_tables(m)
    self.cm = E.B2Model()
    for k, v in E.model_scalars(m, self.tables).items():
      setattr(self.cm, k, v)
    self._keep = []
    self.model_arrays = {}
    for i, (_, attr, _rl) in enumerate(E.MODEL_FLOAT_FIELDS):
      a = np.ascontiguousarray(E.model_float_array(m, attr).astype(self.dtype))
      if a.size == 0:
        a = np.zeros(1, self.dtype)
      self.model_arrays[attr] = a
      self.cm.mf[i] = a.ctypes.data
      self.cm.mf_stride[i] = 0
    for i, (_, attr) in enumerate(E.MODEL_INT_FIELDS):
      a = E.model_int_array(m, attr, self.tables)
      self._keep.append(a)
      self.cm.mi[i] = a.ctypes.data
    self.cd = E.B2Data()
    self.cd.nworld, self.cd.nconmax, self.cd.njmax = nworld, nconmax, njmax
    self.cd.flags = 1  # export dense efc_J for comparison with the oracle
    self.nworld = nworld
    for i, (_, attr, shp) in enumerate(E.DATA_FLOAT_FIELDS):
      a = np.zeros((nworld,) + tuple(shp(m, nconmax, njmax)), self.dtype)
      if a.size == 0:
        a = np.zeros((nworld, 1), self.dtype)
      setattr(self, attr, a)
      self.cd.df[i] = a.ctypes.data
    for i, (_, attr, shp) in enumerate(E.DATA_INT_FIELDS):
      a = np.zeros((nworld,) + tuple(shp(m, nconmax, njmax)), np.int32)
      setattr(self, attr, a)
      self.cd.di[i] = a.ctypes.data
    self.qpos[:] = m.qpos0
    E.init_static_numpy(m, self.geom_xpos, self.geom_xmat)

  def expand(self, attr: str) -> np.ndarray:
    """Per-world copy of a model float field (expand_model_fields)."""
    idx = [a for _, a, _ in E.MODEL_FLOAT_FIELDS].index(attr)
    row = self.model_arrays[attr]
    big = np.ascontiguousarray(np.tile(row, (self.nworld, 1)))
    self.model_arrays[attr] = big
    self.cm.mf[idx] = big.ctypes.data
    self.cm.mf_stride[idx] = row.size
    return big

  def step(self, after_forward: bool = False):
    """after_forward: let the position / velocity stage reuse what the preceding pass left (b2_substeps_after_forward)."""
    self.lib.b2h_set_memo_consume(int(after_forward))
    self.lib.b2h_run(C.byref(self.cm), C.byref(self.cd), 1, 0, self.nworld)
    self.lib.b2h_set_memo_consume(0)

  def memo_hits(self) -> int:
    self.lib.b2h_memo_hit_count.restype = C.c_longlong
    return int(self.lib.b2h_memo_hit_count())

  def forward(self):
    self.lib.b2h_run(C.byref(self.cm), C.byref(self.cd), 0, 0, self.nworld)


_NP = {"f4": None, "i4": np.int32, "i8": np.int64, "u1": np.uint8}


class HostEnv:
  """The env-layer device code (b2_env.cuh) host-emulated on top of a HostSim."""

  def __init__(self, sim: HostSim, cfg, inject=False):
    self.sim, self.cfg = sim, cfg
    assert sim.lib.b2h_sizeof_env() == C.sizeof(A.B2Env)
    self.ce = A.B2Env()
    A.fill_params(self.ce, cfg)
    tables = A.table_values(cfg)
    self.bufs = {}
    for i, (_, attr, kind, dt, shp) in enumerate(A.BUFFERS):
      dtype = sim.dtype if dt == "f4" else _NP[dt]
      if kind == "table":
        a = np.ascontiguousarray(np.asarray(tables[attr]).astype(dtype))
      elif kind == "state":
        a = np.zeros((sim.nworld,) + tuple(shp(cfg)), dtype)
      else:
        a = np.zeros(tuple(shp(cfg)), dtype)
      self.bufs[attr] = a
      self.ce.buf[i] = a.ctypes.data
    self.inject = None
    if inject:
      self.inject = np.zeros((sim.nworld, A.NDRAW), sim.dtype)
      self.ce.buf[A.K["B2E_B_INJECT_U"]] = self.inject.ctypes.data

  def __getattr__(self, name):
    try:
      return self.__dict__["bufs"][name]
    except KeyError:
      raise AttributeError(name) from None

  def stage(self, k: int) -> None:
    rc = self.sim.lib.b2h_env_stage(C.byref(self.sim.cm), C.byref(self.sim.cd), C.byref(self.ce), k)
    assert rc == 0

  def forward_masked(self, mask=None) -> None:
    ptr = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data
    self.sim.lib.b2h_forward_masked(C.byref(self.sim.cm), C.byref(self.sim.cd), C.c_void_p(ptr))


class HostTrk:
  """The tracking-env device code (b2_trk.cuh) host-emulated on top of a HostSim."""

  def __init__(self, sim: HostSim, cfg, motion: dict, inject=False):
    self.sim, self.cfg = sim, cfg
    assert sim.lib.b2h_sizeof_trk() == C.sizeof(TA.B2Trk)
    self.ce = TA.B2Trk()
    TA.fill_params(self.ce, cfg)
    tables = TA.table_values(cfg)
    self.bufs = {}
    for i, (_, attr, kind, dt, shp) in enumerate(TA.BUFFERS):
      dtype = sim.dtype if dt == "f4" else _NP[dt]
      if kind == "table":
        a = np.ascontiguousarray(np.asarray(tables[attr]).astype(dtype))
      elif kind == "motion":
        a = np.ascontiguousarray(np.asarray(motion[attr]).astype(dtype))
      elif kind == "state":
        a = np.zeros((sim.nworld,) + tuple(shp(cfg)), dtype)
      else:
        a = np.zeros(tuple(shp(cfg)), dtype)
      self.bufs["m_" + attr if kind == "motion" else attr] = a
      self.ce.buf[i] = a.ctypes.data
    self.bufs["rel_quat"][..., 0] = 1
    self.inject = None
    if inject:
      self.inject = np.zeros((sim.nworld, TA.NDRAW), sim.dtype)
      self.ce.buf[TA.K["B2T_B_INJECT_U"]] = self.inject.ctypes.data

  def __getattr__(self, name):
    try:
      return self.__dict__["bufs"][name]
    except KeyError:
      raise AttributeError(name) from None

  def refresh(self):
    TA.fill_params(self.ce, self.cfg)

  def stage(self, k: int) -> None:
    rc = self.sim.lib.b2h_trk_stage(C.byref(self.sim.cm), C.byref(self.sim.cd), C.byref(self.ce), k)
    assert rc == 0
