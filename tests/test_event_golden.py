"""Reset events, interval push events, the velocity command and model-field randomisation against golden vectors produced by
THE REFERENCE'S OWN CODE with its random draws recorded (tools/make_event_golden.py): `reset_root_state_uniform`,
`reset_joints_by_offset`, `push_by_setting_velocity` under the real `EventManager`, `UniformVelocityCommand` under
`CommandTerm.reset / compute`, `randomize_field`.  The recorded unit uniforms are placed into the draw slots of the env oracle
(the kernels are checked against the oracle with injected draws in tests/test_env_hostemu.py and tests/test_gpu_env.py), so
both sides consume the same uniforms and every value is compared, not a distribution."""
from pathlib import Path
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import mjlab_b200
from mjlab_b200 import envabi as A
from mjlab_b200 import envcfg
from oracle import envstep as OE
from tests.envutil import load_golden, snapshot

GOLDEN = Path(__file__).resolve().parent / "golden"
K = A.K
KEYS6 = ("x", "y", "z", "roll", "pitch", "yaw")
TOL = dict(rtol=3e-6, atol=3e-6)


def _cfg(robot, e):
  m = mjlab_b200.load_model(f"{robot}_flat")
  cfg = envcfg.velocity_cfg(m, robot)
  c = e["cmd_cfg"]
  cfg.cmd_resample, cfg.cmd_lin_x, cfg.cmd_lin_y, cfg.cmd_ang_z, cfg.cmd_heading = (c[0], c[1]), (c[2], c[3]), (c[4], c[5]), (c[6], c[7]), (c[8], c[9])
  cfg.rel_standing, cfg.rel_heading, cfg.heading_stiffness = float(c[10]), float(c[11]), float(c[12])
  cfg.reset_pose_range = {k: tuple(e["reset_pose_range"][i]) for i, k in enumerate(KEYS6)}
  cfg.reset_velocity_range = {k: tuple(e["reset_vel_range"][i]) for i, k in enumerate(KEYS6)}
  cfg.push_velocity_range = {k: tuple(e["step_push_range"][i]) for i, k in enumerate(KEYS6)}
  cfg.reset_joint_pos, cfg.reset_joint_vel = tuple(e["reset_jp_range"]), tuple(e["reset_jv_range"])
  cfg.push_interval = tuple(e["push_interval"])
  cfg.enable_corruption = False
  return m, cfg


@pytest.mark.parametrize("robot", ["g1", "go1"])
def test_reset_events_and_command_reset_match_reference(robot):
  e = np.load(GOLDEN / f"events_{robot}.npz")
  m, cfg = _cfg(robot, e)
  g = load_golden(robot)
  d = snapshot(g, 3, m)
  B = d.qpos.shape[0]
  ids = e["reset_ids"]
  s = OE.make_state(cfg, B, e["origins"])
  s.cmd_metrics[:, 0] = 0.5
  before_q, before_v = d.qpos.copy(), d.qvel.copy()
  U = np.zeros((B, A.NDRAW), np.float32)
  U[ids, K["B2E_DRAW_RESET_POSE"]: K["B2E_DRAW_RESET_POSE"] + 6] = e["u_reset_pose"]
  U[ids, K["B2E_DRAW_RESET_VEL"]: K["B2E_DRAW_RESET_VEL"] + 6] = e["u_reset_vel"]
  U[ids, K["B2E_DRAW_RESET_JOINT_POS"]: K["B2E_DRAW_RESET_JOINT_POS"] + cfg.nj] = e["u_reset_jp"]
  U[ids, K["B2E_DRAW_RESET_JOINT_VEL"]: K["B2E_DRAW_RESET_JOINT_VEL"] + cfg.nj] = e["u_reset_jv"]
  U[ids, K["B2E_DRAW_RESET_CMD"]: K["B2E_DRAW_RESET_CMD"] + 7] = e["u_reset_cmd"][:, :7]
  OE.reset_idx(cfg, m, s, d, ids, U)
  np.testing.assert_allclose(d.qpos, e["ref_reset_qpos"], **TOL)
  np.testing.assert_allclose(d.qvel, e["ref_reset_qvel"], **TOL)
  rest = np.setdiff1d(np.arange(B), ids)
  np.testing.assert_array_equal(d.qpos[rest], before_q[rest])      # other envs untouched
  np.testing.assert_array_equal(d.qvel[rest], before_v[rest])
  assert np.abs(d.qvel[ids, 3:6]).max() > 0.05                     # world -> body angular velocity was exercised
  np.testing.assert_allclose(s.command, e["ref_reset_command"], **TOL)
  np.testing.assert_allclose(s.heading_target, e["ref_reset_heading_target"], **TOL)
  np.testing.assert_array_equal(s.is_heading, e["ref_reset_is_heading"])
  np.testing.assert_array_equal(s.is_standing, e["ref_reset_is_standing"])
  np.testing.assert_allclose(s.cmd_time_left, e["ref_reset_time_left"], **TOL)
  np.testing.assert_array_equal(s.cmd_counter, e["ref_reset_counter"])
  # metrics of the finished episodes are logged, then cleared (command_manager.py:40-51)
  np.testing.assert_allclose(s.log_env[ids, K["B2E_LOG_CMD_ERR_XY"]].mean(), e["ref_reset_metric_mean"], rtol=1e-6)
  assert np.all(s.cmd_metrics[ids] == 0) and np.all(s.cmd_metrics[rest, 0] == 0.5)


@pytest.mark.parametrize("robot", ["g1", "go1"])
def test_command_compute_and_interval_push_match_reference(robot):
  e = np.load(GOLDEN / f"events_{robot}.npz")
  m, cfg = _cfg(robot, e)
  g = load_golden(robot)
  B = g["data_qpos"].shape[1]
  s = OE.make_state(cfg, B, e["origins"])
  s.command[:], s.heading_target[:] = e["step_command0"], e["step_heading_target0"]
  s.is_heading[:], s.is_standing[:], s.cmd_time_left[:] = e["step_is_heading0"], e["step_is_standing0"], e["step_time_left0"]
  s.cmd_counter[:] = 1
  s.push_time_left[:] = e["step_push_time_left0"]
  t0, T = int(e["step_t0"]), e["step_command"].shape[0]
  npush = nres = 0
  for k in range(T):
    d = snapshot(g, t0 + k, m)
    U = np.zeros((B, A.NDRAW), np.float32)
    U[:, K["B2E_DRAW_CMD"]: K["B2E_DRAW_CMD"] + 7] = e["step_u_cmd"][k][:, :7]
    U[:, K["B2E_DRAW_PUSH"]] = e["step_u_push_t"][k]
    U[:, K["B2E_DRAW_PUSH"] + 1: K["B2E_DRAW_PUSH"] + 7] = e["step_u_push_v"][k]
    cnt = s.cmd_counter.copy()
    qv = d.qvel.copy()
    OE.obs_stage(cfg, m, s, d, u=U)
    np.testing.assert_array_equal(s.cmd_counter - cnt, e["step_cmd_ids"][k].astype(np.int32))    # the same envs resample ...
    np.testing.assert_array_equal(np.any(d.qvel != qv, axis=1), e["step_push_ids"][k])            # ... and the same envs are pushed
    np.testing.assert_allclose(s.command, e["step_command"][k], **TOL)
    np.testing.assert_allclose(s.heading_target, e["step_heading_target"][k], **TOL)
    np.testing.assert_array_equal(s.is_heading, e["step_is_heading"][k])
    np.testing.assert_array_equal(s.is_standing, e["step_is_standing"][k])
    np.testing.assert_allclose(s.cmd_time_left, e["step_time_left"][k], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(s.cmd_counter, e["step_counter"][k])
    np.testing.assert_allclose(s.cmd_metrics, e["step_metrics"][k], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s.push_time_left, e["step_push_time_left"][k], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d.qvel, e["step_qvel"][k], **TOL)
    npush += int(e["step_push_ids"][k].sum()); nres += int(e["step_cmd_ids"][k].sum())
  assert npush >= 8 and nres >= 8


class _Replay:
  def __init__(self, draws):
    self.draws, self.i = draws, 0

  def __call__(self, distribution, lo, hi, shape, device):
    import math
    raw = torch.as_tensor(self.draws[self.i])
    self.i += 1
    if distribution != "gaussian":
      assert tuple(raw.shape) == tuple(shape), (raw.shape, shape)      # same call sequence and shapes as the reference
    if distribution == "uniform":
      return raw * (hi - lo) + lo
    if distribution == "log_uniform":
      return torch.exp(raw * (math.log(hi) - math.log(lo)) + math.log(lo))
    return torch.as_tensor(self.draws[self.i - 1])      # gaussian: the reference's torch.normal result (ONE shared sample)


def test_randomize_field_matches_reference(monkeypatch):
  """events.randomize_field on the same stand-in model and the same draws as the reference's randomize_field: abs / scale / add,
  tuple and per-axis ranges, default axes, explicit axes, env subsets, and no accumulation on repeated scale calls."""
  from mjlab_b200 import events as EV
  r = np.load(GOLDEN / "randomize_field.npz")
  B = 5
  base = {k: torch.as_tensor(r[f"base_{k}"]) for k in ("geom_friction", "body_mass", "body_ipos", "dof_damping", "dof_armature")}
  model = NS(**{k: v[None].repeat(B, *([1] * v.dim())).clone() for k, v in base.items()})
  idx = NS(geom_ids=torch.as_tensor(r["idx_geom_ids"]).long(), body_ids=torch.as_tensor(r["idx_body_ids"]).long(),
           joint_v_adr=torch.as_tensor(r["idx_joint_v_adr"]).long(), joint_q_adr=torch.zeros(5, dtype=torch.long),
           joint_ids=torch.arange(5), site_ids=torch.zeros(1, dtype=torch.long))
  env = NS(num_envs=B, device="cpu", scene={"robot": NS(indexing=idx)}, sim=NS(model=model, get_default_field=lambda f: base[f]))
  calls = [
    ("geom_friction", (0.3, 1.2), "uniform", "abs", dict(geom_ids=[0, 2]), None, torch.tensor([0, 1, 3])),
    ("geom_friction", {0: (0.5, 1.5), 1: (0.9, 1.1)}, "uniform", "scale", dict(geom_ids=slice(None)), None, None),
    ("body_mass", (0.8, 1.25), "log_uniform", "scale", dict(body_ids=[0, 2]), None, None),
    ("body_mass", (0.8, 1.25), "log_uniform", "scale", dict(body_ids=[0, 2]), None, None),
    ("body_ipos", {0: (-0.05, 0.05), 2: (-0.02, 0.02)}, "uniform", "add", dict(body_ids=[1]), None, torch.tensor([4, 2])),
    ("dof_damping", (0.0, 0.3), "uniform", "abs", dict(joint_ids=[0, 1, 4]), None, None),
    ("dof_armature", (1.0, 0.1), "gaussian", "scale", dict(joint_ids=slice(None)), None, torch.tensor([1])),
    ("body_ipos", (-0.01, 0.01), "uniform", "add", dict(body_ids=slice(None)), [1], None),
  ]
  assert int(r["ncalls"]) == len(calls)
  for i, (field, ranges, dist, op, sel, axes, env_ids) in enumerate(calls):
    rp = _Replay([r[f"call{i}_draw{j}"] for j in range(int(r[f"call{i}_ndraws"]))])
    monkeypatch.setattr(EV, "_sample", rp)
    acfg = NS(name="robot", geom_ids=slice(None), body_ids=slice(None), joint_ids=slice(None), site_ids=slice(None))
    for k, v in sel.items():
      setattr(acfg, k, v)
    EV.randomize_field(env, env_ids, field, ranges, dist, op, axes=axes, asset_cfg=acfg)
    assert rp.i == len(rp.draws), (i, field)
    np.testing.assert_allclose(getattr(model, field).numpy(), r[f"call{i}_after"], rtol=2e-6, atol=1e-7, err_msg=f"call {i} {field}")
  # repeated scale calls start from the defaults (tests/test_domain_randomization.py:182-240): bounded by the range, not its square
  ratio = (model.body_mass / base["body_mass"]).numpy()
  assert ratio.min() >= 0.8 - 1e-6 and ratio.max() <= 1.25 + 1e-6
