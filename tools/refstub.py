"""Import the reference's pure-torch Python layers in a container without MuJoCo / Warp.

`install()` registers an import hook that serves placeholder modules for the native / GUI dependencies
(mujoco, mujoco_warp, warp, trimesh, viser, ...) the reference imports at module scope, and puts
/root/reference/src on sys.path.  Only used by tools/make_mdp_golden.py to GENERATE golden vectors from the
reference's own functions (EntityData, reward / observation / termination terms, contact-sensor air time,
velocity command); nothing at run time imports it and the GPU box never sees /root/reference.
"""
import importlib.abc
import importlib.machinery
import sys
import types
from unittest.mock import MagicMock

STUB_ROOTS = {"mujoco", "mujoco_warp", "warp", "trimesh", "viser", "mediapy", "tensordict", "rsl_rl", "onnx", "wandb",
              "tyro", "torchrunx", "prettytable", "glfw", "OpenGL", "PIL", "imageio", "moviepy", "onnxruntime"}


class _Stub(types.ModuleType):
  def __getattr__(self, name):
    if name == "__version__":
      return "0.0.0-stub"
    if name.startswith("__"):
      raise AttributeError(name)
    m = MagicMock(name=f"{self.__name__}.{name}")
    setattr(self, name, m)
    return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
  def find_spec(self, fullname, path, target=None):
    if fullname.split(".")[0] in STUB_ROOTS:
      return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
    return None

  def create_module(self, spec):
    m = _Stub(spec.name)
    m.__path__ = []
    return m

  def exec_module(self, module):
    pass


def install(ref_src="/root/reference/src"):
  real = set()
  for r in list(STUB_ROOTS):
    try:
      __import__(r)
      real.add(r)
    except Exception:
      pass
  for r in real:
    STUB_ROOTS.discard(r)
  sys.meta_path.insert(0, _Finder())
  if ref_src not in sys.path:
    sys.path.insert(0, ref_src)
