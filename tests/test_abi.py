"""The C-ABI library loads and exports every symbol include/b200sim.h declares; the Python field tables mirror
the header enums; the product refuses to run without the CUDA library (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

from mjlab_b200 import engine as E

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "b200sim.h").read_text()


def enum_names(tag):
  body = re.search(r"enum %s \{(.*?)\};" % tag, HEADER, re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  return [n.split("=")[0].strip() for n in body.split(",") if n.strip()]


@pytest.mark.parametrize("tag,prefix,table", [
  ("b2_mf", "B2_MF_", E.MODEL_FLOAT_FIELDS), ("b2_mi", "B2_MI_", E.MODEL_INT_FIELDS),
  ("b2_df", "B2_DF_", E.DATA_FLOAT_FIELDS), ("b2_di", "B2_DI_", E.DATA_INT_FIELDS)])
def test_field_tables_mirror_header(tag, prefix, table):
  names = enum_names(tag)
  assert names[-1] == prefix + "COUNT"
  assert [prefix + t[0] for t in table] == names[:-1]


def declared_functions():
  return re.findall(r"^\s*(?:int|long long|const char\*)\s+(b2_\w+)\s*\(", HEADER, re.M)


def test_library_exports_declared_symbols():
  __import__("__graft_entry__").build()
  lib = ctypes.CDLL(str(E.LIB_PATH))
  fns = declared_functions()
  assert {"b2_step", "b2_forward", "b2_reset", "b2_repeat_rows", "b2_check", "b2_stage", "b2_last_error"} <= set(fns)
  for fn in fns:
    assert hasattr(lib, fn), fn
  assert lib.b2_abi_version() == 2
  assert lib.b2_sizeof_model() == ctypes.sizeof(E.B2Model)
  assert lib.b2_sizeof_data() == ctypes.sizeof(E.B2Data)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
  monkeypatch.setattr(E, "_lib", None)
  monkeypatch.setenv("B200SIM_LIB", str(tmp_path / "nope.so"))
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    E.load_library()


def test_simulation_requires_cuda():
  import mjlab_b200
  from mjlab_b200.sim import Simulation, SimulationCfg
  m = mjlab_b200.load_model("go1_flat")
  with pytest.raises(RuntimeError):
    Simulation(2, SimulationCfg(), m, "cpu")


def test_product_does_not_import_oracle():
  """The oracle is test infrastructure: nothing under mjlab_b200/ may import or execute it."""
  for p in (ROOT / "mjlab_b200").rglob("*.py"):
    src = p.read_text()
    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), p
    assert "hostemu" not in src, p


ENV_HEADER = (ROOT / "include" / "b200env.h").read_text()


def test_env_header_symbols_exported_and_struct_matches():
  """include/b200env.h: every declared entry point is exported, and the ctypes mirror (mjlab_b200/envabi.py, which reads
  the enum values from the header itself) has the size the library reports."""
  from mjlab_b200 import envabi as A
  __import__("__graft_entry__").build()
  lib = ctypes.CDLL(str(E.LIB_PATH))
  fns = re.findall(r"^\s*int\s+(b2_\w+)\s*\(", ENV_HEADER, re.M)
  assert {"b2_env_action", "b2_env_airtime", "b2_env_post", "b2_env_obs", "b2_env_log", "b2_env_reset", "b2_forward_masked",
          "b2_actuator_pd", "b2_delay_step", "b2_actuator_mlp", "b2_env_check", "b2_sizeof_env"} <= set(fns)
  for fn in fns:
    assert hasattr(lib, fn), fn
  lib.b2_sizeof_env.restype = ctypes.c_int
  assert lib.b2_sizeof_env() == ctypes.sizeof(A.B2Env)
  # the buffer table follows the header enum order, and the reward / termination name lists match the enum sizes
  from mjlab_b200 import envcfg
  assert len(envcfg.REWARD_NAMES) == A.NREW and len(envcfg.TERMINATION_NAMES) == A.NTERM
  assert A.K["B2E_DRAW_NOISE"] + 160 == A.K["B2E_DRAW_TERRAIN"] and A.K["B2E_DRAW_TERRAIN"] + 3 == A.NDRAW


def test_env_requires_cuda():
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  with pytest.raises(RuntimeError):
    ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=2), "cpu")
  from mjlab_b200.actuators import IdealPdActuator
  with pytest.raises(RuntimeError):
    IdealPdActuator(2, 3, 1.0, 0.1, device="cpu")


def test_solver_slot_map_is_a_permutation():
  """k_order scatters the worlds (heaviest-first rank) onto solver slots; for EVERY batch size - odd, tiny, larger than
  the resident grid - the map must be a permutation of [0, n), else a world is solved twice and another not at all."""
  import ctypes
  lib = E.load_library()
  lib.b2_debug_solver_slot.argtypes = [ctypes.c_int, ctypes.c_int]
  lib.b2_debug_solver_slot.restype = ctypes.c_int
  for n in list(range(1, 130)) + [255, 256, 257, 2367, 2368, 2369, 2370, 4095, 4096, 4097, 8191, 8192]:
    slots = sorted(lib.b2_debug_solver_slot(p, n) for p in range(n))
    assert slots == list(range(n)), n
  # first wave: the c-th heaviest world shares CTA c (slots 2c, 2c+1) with a light one
  n = 4096
  assert [lib.b2_debug_solver_slot(p, n) for p in (0, 1, 2)] == [0, 2, 4]
  assert lib.b2_debug_solver_slot(n - 1, n) % 2 == 1 and lib.b2_debug_solver_slot(n - 1184, n) == 1
