"""A stand-in for a compiled `mujoco.MjModel`, for the boundary tests of `Simulation(num_envs, cfg, mujoco.MjModel, device)`
(sim/sim.py:122-149) in an image without the `mujoco` wheel.  It exposes MuJoCo's OBJECT surface - nested `opt` / `stat`
structs, `model.body(i).name` style accessors, `exclude_signature`, arrays under MjModel's names and shapes - built from a
compiled scene of this repo, and NOTHING of `ModelDesc`'s own conventions (no flat `opt_*`, no name lists, no derived
collision tables), so `ModelDesc.from_mjmodel` has to do the whole conversion.  Recorded from a model, not from MuJoCo: it
tests the converter's logic and the attribute surface, not MuJoCo's compiler."""
from types import SimpleNamespace as NS

import numpy as np

from mjlab_b200.mjcf import ModelDesc


class FakeMjModel:
  def __init__(self, m: ModelDesc):
    for k in ModelDesc._MJ_INTS:
      setattr(self, k, int(getattr(m, k)))
    for k in ModelDesc._MJ_ARRAYS:
      setattr(self, k, np.array(getattr(m, k)))
    self.opt = NS(timestep=float(m.opt_timestep), gravity=np.array(m.opt_gravity, float), integrator=int(m.opt_integrator),
                  cone=int(m.opt_cone), solver=int(m.opt_solver), iterations=int(m.opt_iterations),
                  ls_iterations=int(m.opt_ls_iterations), tolerance=float(m.opt_tolerance), ls_tolerance=float(m.opt_ls_tolerance),
                  impratio=float(m.opt_impratio), disableflags=int(m.opt_disableflags), ccd_iterations=50)
    self.stat = NS(meaninertia=float(m.stat_meaninertia))
    ex = np.asarray(m.exclude_pairs, np.int64).reshape(-1, 2)
    self.exclude_signature = (ex[:, 0] << 16) + ex[:, 1]
    self._names = {k: list(getattr(m, f"{k}_names")) for k in ("body", "jnt", "geom", "site", "actuator", "sensor")}

  def _acc(self, kind, i):
    return NS(id=int(i), name=self._names[kind][int(i)])

  def body(self, i): return self._acc("body", i)          # noqa: E704
  def joint(self, i): return self._acc("jnt", i)          # noqa: E704
  def geom(self, i): return self._acc("geom", i)          # noqa: E704
  def site(self, i): return self._acc("site", i)          # noqa: E704
  def actuator(self, i): return self._acc("actuator", i)  # noqa: E704
  def sensor(self, i): return self._acc("sensor", i)      # noqa: E704
