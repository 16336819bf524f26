"""Helpers shared by the env-layer tests: golden snapshots -> batched data namespaces."""
from pathlib import Path
from types import SimpleNamespace as NS

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
FIELDS = ["qpos", "qvel", "xpos", "xquat", "xipos", "subtree_com", "cvel", "site_xpos", "site_xmat", "sensordata", "qacc", "actuator_force", "time"]


def load_golden(robot):
  return np.load(GOLDEN / f"mdp_velocity_{robot}.npz")


def snapshot(g, t, m, dtype=np.float32):
  d = NS(**{f: np.array(g[f"data_{f}"][t], dtype) for f in FIELDS})
  B = d.qpos.shape[0]
  d.ctrl = np.zeros((B, m.nu), dtype)
  d.qacc_warmstart = np.zeros((B, m.nv), dtype)
  return d


def init_state_from_golden(s, g):
  s.command[:] = g["command0"]
  s.heading_target[:] = g["heading_target"]
  s.is_heading[:] = g["is_heading"]
  s.is_standing[:] = g["is_standing"]
  s.episode_length[:] = g["episode_len0"]


def snapshot_from(name):
  """path of a golden fixture; snapshot_from.load(g, t, m) -> data namespace of env step t."""
  return GOLDEN / f"{name}.npz"


def _load(g, t, m, dtype=np.float32):
  d = NS(**{f: np.array(g[f"data_{f}"][t], dtype) for f in FIELDS if f"data_{f}" in g.files})
  B = d.qpos.shape[0]
  d.ctrl = np.zeros((B, m.nu), dtype)
  d.qacc_warmstart = np.zeros((B, m.nv), dtype)
  d.qacc = np.zeros((B, m.nv), dtype)
  d.time = np.zeros(B, dtype)
  return d


snapshot_from.load = _load
