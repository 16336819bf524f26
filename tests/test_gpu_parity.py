"""GPU parity: CUDA path (through the C ABI via Simulation) vs the float64 oracle.

Bit-exact: contact counts, contact geom-pair lists, efc row counts.
Tolerance (stated per assertion): fp32 device arithmetic vs fp64 oracle, north_star's 1e-4 relative per step on qpos, qvel,
qacc, efc_force and sensordata.  Measured on these states (float32 host build of the kernel source vs float64, round 2): qvel
1e-6 after 5 steps, efc_force 2e-5, qacc 5e-6, sensordata 8e-6, efc_aref 5e-6, efc_D 1e-4 (G1) - the bounds below leave 4-10 x room
for the CUDA build's different summation order.  `relerr` normalises by the field's max-norm per world: components near zero have
no relative error of their own.  The population statistic (128 worlds, N steps, worst world) is profiles/r02_divergence.txt.
"""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def make_sim(m, nworld, nconmax=35, njmax=300, integrator="implicitfast", export_efc_J=True):
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  cfg = SimulationCfg(nconmax=nconmax, njmax=njmax, export_efc_J=export_efc_J,
                      mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20, integrator=integrator))
  return Simulation(nworld, cfg, m, "cuda:0")


def load_state(sim, qpos, qvel, ctrl, xfrc):
  f = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda:0")
  sim.data.qpos[:] = f(qpos)
  sim.data.qvel[:] = f(qvel)
  sim.data.ctrl[:] = f(ctrl)
  sim.data.xfrc_applied[:] = f(xfrc)
  sim.data.qacc_warmstart[:] = 0
  sim.data.time[:] = 0


def relerr(a, b):
  a, b = np.asarray(a, float), np.asarray(b, float)
  return np.abs(a - b).max() / max(np.abs(b).max(), 1e-6)


@pytest.mark.parametrize("name", ["go1_flat", "g1_flat"])
def test_forward_stagewise(name):
  import mjlab_b200
  m = mjlab_b200.load_model(name)
  nw = 8
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=1, xfrc_world=3)
  sim = make_sim(m, nw, nconmax=64)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  sim.forward()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 64, 300, integrate_last=False)
  D = sim.data
  for w, d in enumerate(ds):
    assert int(D.ncon[w]) == len(d.contacts)
    assert D.contact_geom[w, : len(d.contacts)].cpu().tolist() == [list(p) for p in util.contact_pairs(d)]
    assert int(D.nefc[w]) == d.nefc
    for fld, tol in [("xpos", 1e-6), ("xquat", 1e-6), ("geom_xpos", 1e-6), ("subtree_com", 1e-6), ("cdof", 2e-6),
                     ("qM", 1e-5), ("cvel", 1e-5), ("qfrc_bias", 1e-4), ("qfrc_actuator", 1e-5), ("qfrc_smooth", 1e-4),
                     ("qacc_smooth", 1e-4)]:
      assert relerr(getattr(D, fld)[w].cpu().numpy(), getattr(d, fld)) < tol, (w, fld)
    n = d.nefc
    if n:
      assert relerr(D.efc_J[w, :n].cpu().numpy(), d.efc_J) < 1e-5
      assert relerr(D.efc_D[w, :n].cpu().numpy(), d.efc_D) < 5e-4
      assert relerr(D.efc_aref[w, :n].cpu().numpy(), d.efc_aref) < 1e-4
      assert relerr(D.efc_force[w, :n].cpu().numpy(), d.efc_force) < 2e-4
    assert relerr(D.qacc[w].cpu().numpy(), d.qacc) < 1e-4
    assert relerr(D.sensordata[w].cpu().numpy(), d.sensordata) < 1e-4


@pytest.mark.parametrize("name,integrator", [("go1_flat", "implicitfast"), ("g1_flat", "implicitfast"), ("go1_flat", "euler")])
def test_step_trajectory(name, integrator):
  import mjlab_b200
  m = mjlab_b200.load_model(name)
  m.opt_integrator = {"euler": 0, "implicitfast": 3}[integrator]
  nw, nstep = 6, 5
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=2)
  sim = make_sim(m, nw, integrator=integrator)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  for _ in range(nstep):
    sim.step()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, nstep, 35, 300)
  worst_q = worst_v = 0.0
  for w, d in enumerate(ds):
    assert int(sim.data.ncon[w]) == len(d.contacts)
    worst_q = max(worst_q, relerr(sim.data.qpos[w].cpu().numpy(), d.qpos))
    worst_v = max(worst_v, relerr(sim.data.qvel[w].cpu().numpy(), d.qvel))
    assert abs(float(sim.data.time[w]) - d.time) < 1e-6
  # 1e-4 relative per step (north_star), accumulated over nstep steps
  assert worst_q < 1e-4 * nstep, worst_q
  assert worst_v < 1e-4 * nstep, worst_v


def test_reset_and_graph_replay(go1):
  nw = 16
  sim = make_sim(go1, nw)
  qpos, qvel, ctrl, xfrc = util.random_states(go1, nw, seed=3)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  for _ in range(4):
    sim.step()
  ids = torch.tensor([1, 5, 7], device="cuda:0")
  before = sim.data.qpos.clone()
  sim.reset(ids)
  torch.cuda.synchronize()
  q0 = torch.as_tensor(go1.qpos0, dtype=torch.float32, device="cuda:0")
  for w in range(nw):
    if w in (1, 5, 7):
      assert torch.equal(sim.data.qpos[w], q0)
      assert float(sim.data.qvel[w].abs().max()) == 0.0
      assert float(sim.data.qacc_warmstart[w].abs().max()) == 0.0
      assert float(sim.data.time[w]) == 0.0
    else:
      assert torch.equal(sim.data.qpos[w], before[w])
  sim.reset()
  torch.cuda.synchronize()
  assert torch.equal(sim.data.qpos, q0.expand(nw, -1))


def test_determinism(g1):
  nw = 32
  outs = []
  for _ in range(2):
    sim = make_sim(g1, nw)
    qpos, qvel, ctrl, xfrc = util.random_states(g1, nw, seed=4)
    load_state(sim, qpos, qvel, ctrl, xfrc)
    for _ in range(8):
      sim.step()
    torch.cuda.synchronize()
    outs.append((sim.data.qpos.clone(), sim.data.qvel.clone(), sim.data.sensordata.clone()))
  for a, b in zip(outs[0], outs[1]):
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["go1_flat", "g1_flat"])
def test_gpu_matches_committed_golden(name):
  """CUDA path against tests/golden (made by tools/make_golden.py); no oracle needed at run time."""
  import mjlab_b200
  g = np.load(f"tests/golden/physics_{name}.npz")
  m = mjlab_b200.load_model(name)
  nw = g["qpos0"].shape[0]
  sim = make_sim(m, nw, nconmax=int(g["nconmax"]), njmax=int(g["njmax"]))
  load_state(sim, g["qpos0"], g["qvel0"], g["ctrl"], g["xfrc"])
  for _ in range(int(g["nstep"])):
    sim.step()
  torch.cuda.synchronize()
  # bit-exact integer results
  assert sim.data.ncon.cpu().tolist() == g["ncon"].tolist()
  assert sim.data.nefc.cpu().tolist() == g["nefc"].tolist()
  for w in range(nw):
    n = int(g["ncon"][w])
    assert sim.data.contact_geom[w, :n].cpu().tolist() == g["contact_geom"][w, :n].tolist()
  # fp32 vs fp64: 1e-4 relative per step on qpos / qvel (north_star), nstep steps
  nstep = int(g["nstep"])
  assert relerr(sim.data.qpos.cpu().numpy(), g["qpos"]) < 1e-4 * nstep
  assert relerr(sim.data.qvel.cpu().numpy(), g["qvel"]) < 1e-4 * nstep
  assert relerr(sim.data.xpos.cpu().numpy(), g["xpos"]) < 1e-4 * nstep


def test_full_size_properties(g1):
  """BASELINE size (4096 worlds): size-independent properties instead of an oracle run.
  Identical worlds give identical results (batch invariance), quaternions stay normalised, no NaNs,
  contact counts bounded, a reset world equals a freshly constructed one."""
  nw = 4096
  sim = make_sim(g1, nw, export_efc_J=False)
  qpos, qvel, ctrl, xfrc = util.random_states(g1, 8, seed=7)
  reps = nw // 8
  load_state(sim, np.tile(qpos, (reps, 1)), np.tile(qvel, (reps, 1)), np.tile(ctrl, (reps, 1)), np.tile(xfrc, (reps, 1, 1)))
  for _ in range(10):
    sim.step()
  torch.cuda.synchronize()
  q = sim.data.qpos.view(reps, 8, -1)
  assert torch.equal(q, q[:1].expand_as(q))
  s = sim.data.sensordata.view(reps, 8, -1)
  assert torch.equal(s, s[:1].expand_as(s))
  assert torch.isfinite(sim.data.qpos).all() and torch.isfinite(sim.data.qvel).all()
  assert torch.allclose(sim.data.qpos[:, 3:7].norm(dim=1), torch.ones(nw, device="cuda:0"), atol=1e-5)
  assert int(sim.data.ncon.max()) <= 35 and int(sim.data.nefc.max()) <= 300
  # small case == same worlds inside the big batch
  small = make_sim(g1, 8, export_efc_J=False)
  load_state(small, qpos, qvel, ctrl, xfrc)
  for _ in range(10):
    small.step()
  torch.cuda.synchronize()
  assert torch.equal(small.data.qpos, sim.data.qpos[:8])


def test_trajectory_divergence_statistic(go1):
  """N-step divergence of the f32 CUDA step from the f64 CPU build (tools/divergence.py reports the full table):
  the median per-world qpos error stays below 1e-5 over 100 steps and contact counts stay identical for most worlds."""
  import torch
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  from oracle.cport import CpuSim
  from tests.util import random_states
  m, nworld = go1, 64
  sim = Simulation(nworld, SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20)), m, "cuda:0")
  qpos, qvel, ctrl, _ = random_states(m, nworld, seed=123, vel=0.2)
  cpu = CpuSim(m, nworld, 35, 300)
  cpu.qpos[:], cpu.qvel[:], cpu.ctrl[:] = qpos, qvel, ctrl
  f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda:0")  # noqa: E731
  sim.data.qpos[:], sim.data.qvel[:], sim.data.ctrl[:] = f32(qpos), f32(qvel), f32(ctrl)
  same = np.ones(nworld, bool)
  for n in range(100):
    sim.step()
    cpu.step(1, 8)
    same &= sim.data.ncon.cpu().numpy().reshape(-1) == cpu.ncon.reshape(-1)
    if n == 0:
      e1 = np.abs(sim.data.qpos.cpu().numpy() - cpu.qpos).max(1)
      assert np.median(e1) < 1e-6 and e1.max() < 1e-4   # ~1e-4 relative per step (north_star), in practice ~1e-7
  err = np.abs(sim.data.qpos.cpu().numpy() - cpu.qpos).max(1)
  assert np.median(err) < 1e-5 and same.mean() > 0.8


def test_solver_arena_contention_does_not_change_results():
  """k_solve's two warps per CTA share one row arena (two-ended allocator; a world waits when its block does not fit next
  to its CTA-mate's).  The smallest arena that still holds the heaviest world of the batch makes most pairs wait; results
  must be bit-identical to the default arena's."""
  import os
  import mjlab_b200
  m = mjlab_b200.load_model("g1_flat")
  nw, nstep = 256, 6
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=9, drop=0.015)

  def run(arena):
    if arena is None:
      os.environ.pop("B200SIM_ARENA", None)
    else:
      os.environ["B200SIM_ARENA"] = str(arena)
    sim = make_sim(m, nw, export_efc_J=False)
    sim.use_cuda_graph = False
    load_state(sim, qpos, qvel, ctrl, xfrc)
    dropped = 0
    for _ in range(nstep):
      sim.step()
      dropped += int(sim.data.nefc_dropped.sum())
    torch.cuda.synchronize()
    D = sim.data
    return dropped, (D.qpos.clone(), D.qvel.clone(), D.efc_force.clone(), D.solver_niter.clone(), D.nefc.clone())

  old = os.environ.get("B200SIM_ARENA")
  try:
    d0, ref = run(None)
    assert d0 == 0
    small = None
    for arena in (800, 1000, 1200, 1600, 2000, 2400):       # floats; the default is ~3400
      d1, out = run(arena)
      if d1 == 0:
        small = arena
        break
  finally:
    if old is None:
      os.environ.pop("B200SIM_ARENA", None)
    else:
      os.environ["B200SIM_ARENA"] = old
  assert small is not None, "no arena below the default held the batch without dropping rows"   # (smaller ones truncate rows, and must not hang)
  assert int(ref[4].max()) > 30
  for a, b in zip(ref, out):
    assert torch.equal(a, b)


@pytest.mark.parametrize("nw", [1, 7, 33])
def test_odd_world_counts(nw):
  """Batch sizes that are odd / smaller than a CTA pair: every world is solved exactly once (scheduling map of k_order)."""
  import mjlab_b200
  m = mjlab_b200.load_model("go1_flat")
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=4)
  sim = make_sim(m, nw)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  for _ in range(3):
    sim.step()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 3, 35, 300)
  for w, d in enumerate(ds):
    assert int(sim.data.ncon[w]) == len(d.contacts) and int(sim.data.nefc[w]) == d.nefc
    assert relerr(sim.data.qpos[w].cpu().numpy(), d.qpos) < 3e-4
    assert relerr(sim.data.qvel[w].cpu().numpy(), d.qvel) < 5e-3


def test_nan_guard_dumps_on_device_nan(tmp_path):
  """`SimulationCfg.nan_guard.enabled`: a NaN that enters one world's state is detected after the step, that world's
  history is dumped, the other worlds carry on (NaN / Inf propagate, they are detected, not prevented: sim.py:267-273)."""
  import mjlab_b200
  from mjlab_b200.nan_guard import NanGuardCfg
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  m = mjlab_b200.load_model("go1_flat")
  cfg = SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20),
                      nan_guard=NanGuardCfg(enabled=True, buffer_size=4, output_dir=str(tmp_path), max_envs_to_dump=3))
  sim = Simulation(8, cfg, m, "cuda:0")
  qpos, qvel, ctrl, xfrc = util.random_states(m, 8, seed=2)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  for _ in range(3):
    sim.step()
  assert not list(tmp_path.iterdir())
  sim.data.qvel[5, 7] = float("nan")
  sim.step()
  dump = np.load(tmp_path / "nan_dump_latest.npz", allow_pickle=True)
  meta = dump["_metadata"].item()
  assert meta["nan_env_ids"] == [5] and meta["buffer_size"] == 4 and meta["detection_step"] == 4
  assert np.isnan(dump["states_step_000003"][0]).any() and np.isfinite(dump["states_step_000002"]).all()
  ok = [w for w in range(8) if w != 5]
  assert torch.isfinite(sim.data.qpos[ok]).all() and torch.isfinite(sim.data.qvel[ok]).all()
