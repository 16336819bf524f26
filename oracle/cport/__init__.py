"""Driver of the multi-threaded CPU build of the physics step (oracle/cport/cport.cpp): TEST / MEASUREMENT infrastructure.

Used by bench.py's `cpu_baseline` leg and `--impl reference` arm to time a compiled, all-cores CPU implementation of the
same algorithm (float64) on the box's host cores, and by tests to cross-check it against the numpy oracle.  Never imported
by the product package.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from mjlab_b200 import engine as E

HERE = Path(__file__).resolve().parent
LIB = HERE.parent / "_build" / "libcport.so"


def build() -> Path:
  subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
  return LIB


class CpuSim:
  """nworld worlds in numpy buffers (float64), stepped by the C++ host build with `nthread` threads."""

  def __init__(self, m, nworld: int, nconmax: int = 35, njmax: int = 300):
    self.lib = C.CDLL(str(build()))
    assert self.lib.cport_sizeof_real() == 8
    E.check_supported(m)
    self.m, self.nworld = m, nworld
    self.tables = E.derived_int_tables(m)
    self.cm = E.B2Model()
    for k, v in E.model_scalars(m, self.tables).items():
      setattr(self.cm, k, v)
    self._keep = []
    for i, (_, attr, _rl) in enumerate(E.MODEL_FLOAT_FIELDS):
      a = np.ascontiguousarray(E.model_float_array(m, attr).astype(np.float64))
      if a.size == 0:
        a = np.zeros(1)
      self._keep.append(a)
      self.cm.mf[i] = a.ctypes.data
      self.cm.mf_stride[i] = 0
    for i, (_, attr) in enumerate(E.MODEL_INT_FIELDS):
      a = E.model_int_array(m, attr, self.tables)
      self._keep.append(a)
      self.cm.mi[i] = a.ctypes.data
    self.cd = E.B2Data()
    self.cd.nworld, self.cd.nconmax, self.cd.njmax, self.cd.flags = nworld, nconmax, njmax, 0
    for i, (_, attr, shp) in enumerate(E.DATA_FLOAT_FIELDS):
      a = np.zeros((nworld,) + tuple(shp(m, nconmax, njmax)))
      if a.size == 0:
        a = np.zeros((nworld, 1))
      setattr(self, attr, a)
      self.cd.df[i] = a.ctypes.data
    for i, (_, attr, shp) in enumerate(E.DATA_INT_FIELDS):
      a = np.zeros((nworld,) + tuple(shp(m, nconmax, njmax)), np.int32)
      setattr(self, attr, a)
      self.cd.di[i] = a.ctypes.data
    self.qpos[:] = m.qpos0
    E.init_static_numpy(m, self.geom_xpos, self.geom_xmat)

  def step(self, nstep: int = 1, nthread: int = 1) -> None:
    self.lib.cport_steps(C.byref(self.cm), C.byref(self.cd), int(nstep), int(nthread))

  def forward(self, nthread: int = 1) -> None:
    self.lib.cport_forward(C.byref(self.cm), C.byref(self.cd), int(nthread))
